mkdir -p gpurun_out
python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/plain11.log 2>&1 || { tail -5 gpurun_out/plain11.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_eval_block -s 1 -c 1 -o gpurun_out/prof_blk3 -f python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu11.log 2>&1
tail -1 gpurun_out/ncu11.log | cut -c1-100
