mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t16.log 2>&1; echo "rc=$?" >> gpurun_out/t16.log
tail -25 gpurun_out/t16.log
python benchmarks/bench_configs.py > gpurun_out/configs16.jsonl 2> gpurun_out/configs16.err; echo "rc=$?"
