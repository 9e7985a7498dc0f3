mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_block_path.py -x -q -m gpu > gpurun_out/t9.log 2>&1; echo "rc=$?" >> gpurun_out/t9.log
tail -4 gpurun_out/t9.log
for cfg in "ASG_BLOCK_MAXS=6" "ASG_BLOCK_MAXS=4" "ASG_BLOCK_MAXS=3"; do
  timeout 300 env $cfg python bench.py --profile --points 2e7 --steps 3 --warmup 1 > gpurun_out/b9.json 2> gpurun_out/b9.err
  python -c "
import json
d=json.load(open('gpurun_out/b9.json')); print('$cfg', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
done
