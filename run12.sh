mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t12.log 2>&1; echo "rc=$?" >> gpurun_out/t12.log
tail -3 gpurun_out/t12.log
timeout 300 python bench.py --profile --points 2e7 --steps 3 --warmup 1 > gpurun_out/b12.json 2> gpurun_out/b12.err
python -c "
import json
d=json.load(open('gpurun_out/b12.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
