mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches14.csv python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu14a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_block -s 1 -c 1 -o gpurun_out/prof14 -f python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu14b.log 2>&1
python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "rc=$?"
python benchmarks/bench_configs.py > gpurun_out/configs14.jsonl 2> gpurun_out/configs14.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench14.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
