O=gpurun_out/r2b; mkdir -p $O
timeout 500 python bench.py --steps 20 --warmup 5 --op-table $O/optable_squeezesegv2_kitti_64x2048_b32.json > $O/bench_r2.json 2> $O/bench_r2.err; tail -2 $O/bench_r2.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-eval > $O/ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:conv_head_kernel --launch-skip 1 -c 1 -o $O/head python tools/ncu_forward.py > $O/ncu_head.log 2>&1
ncu -i $O/head.ncu-rep --page raw --csv > $O/head_raw.csv 2>/dev/null; rm -f $O/head.ncu-rep
ls -la $O
