mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t17.log 2>&1; echo "rc=$?" >> gpurun_out/t17.log
tail -25 gpurun_out/t17.log
python benchmarks/bench_configs.py > gpurun_out/configs17.jsonl 2> gpurun_out/configs17.err; echo "rc=$?"
