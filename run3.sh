mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_block_path.py -x -q -m gpu > gpurun_out/t2.log 2>&1; echo "rc=$?" >> gpurun_out/t2.log
tail -4 gpurun_out/t2.log
for nt in 512 448 384; do
  ASG_BLOCK_THREADS=$nt timeout 300 python bench.py --profile --points 2e7 --steps 3 --warmup 1 > gpurun_out/b3_$nt.json 2> gpurun_out/b3_$nt.err
  python -c "
import json
d=json.load(open('gpurun_out/b3_$nt.json')); print($nt, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
done
