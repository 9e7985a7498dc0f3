mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t15.log 2>&1; echo "rc=$?" >> gpurun_out/t15.log
tail -3 gpurun_out/t15.log
python benchmarks/debug/dbg1.py 2>&1 | tail -9
python benchmarks/bench_configs.py > gpurun_out/configs15.jsonl 2> gpurun_out/configs15.err; echo "rc=$?"
