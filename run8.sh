mkdir -p gpurun_out
python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench8_ref.json 2> gpurun_out/bench8_ref.err; echo "rc=$?"
python benchmarks/bench_configs.py > gpurun_out/configs8.jsonl 2> gpurun_out/configs8.err; echo "rc=$?"
tail -c 600 gpurun_out/bench8.json
