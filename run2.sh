mkdir -p gpurun_out
python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/plain_blk.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_eval_block -s 1 -c 1 -o gpurun_out/prof_blk -f python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu_blk.log 2>&1
tail -3 gpurun_out/ncu_blk.log
ASG_SORT_POINTS=0 ncu --set full --clock-control none --import-source on -k regex:k_eval_block -s 1 -c 1 -o gpurun_out/prof_blk_nosort -f python bench.py --profile --points 8e6 --steps 2 --warmup 1 > gpurun_out/ncu_blk_ns.log 2>&1
tail -3 gpurun_out/ncu_blk_ns.log
