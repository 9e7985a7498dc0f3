mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t5.log 2>&1; echo "rc=$?" >> gpurun_out/t5.log
tail -4 gpurun_out/t5.log
for cfg in "ASG_BLOCK_R=2" ; do
  timeout 300 env $cfg python bench.py --profile --points 2e7 --steps 3 --warmup 1 > gpurun_out/b5.json 2> gpurun_out/b5.err
  python -c "
import json
d=json.load(open('gpurun_out/b5.json')); print('$cfg', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
done
