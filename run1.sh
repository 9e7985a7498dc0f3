set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block_path.py -x -q -m gpu > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
i=0
for cfg in "ASG_SORT_POINTS=0" "ASG_SORT_POINTS=1000000" "ASG_EVAL_KERNEL=stream"; do
  i=$((i+1))
  timeout 300 env $cfg python bench.py --profile --points 2e7 --steps 2 --warmup 1 > gpurun_out/b1_$i.json 2> gpurun_out/b1_$i.err
done
tail -5 gpurun_out/t1.log
for f in gpurun_out/b1_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"; done
