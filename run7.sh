mkdir -p gpurun_out
python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/plain7.log 2>&1 || { tail -5 gpurun_out/plain7.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_blk.csv python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu7a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_block -s 1 -c 1 -o gpurun_out/prof_blk2 -f python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu7b.log 2>&1
tail -2 gpurun_out/ncu7b.log | cut -c1-200
python bench.py --points 1e8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python -c "
import json
d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms'], d['roofline']['point_order_ms'])"
