mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench13_n$N.json 2> gpurun_out/bench13_n$N.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench13_n$N.json')); print($N, d['value'], d['ms_per_step'], d['e2e'], d['gather_ms'])"
