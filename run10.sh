mkdir -p gpurun_out
for cfg in "ASG_BLOCK_MAXS=7" "ASG_BLOCK_MAXS=6"; do
  timeout 300 env $cfg python bench.py --profile --points 2e7 --steps 3 --warmup 1 > gpurun_out/b10.json 2> gpurun_out/b10.err
  python -c "
import json
d=json.load(open('gpurun_out/b10.json')); print('$cfg', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
done
