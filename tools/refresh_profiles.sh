#!/bin/bash
# Regenerates the measured artefacts under profiles/ (run on a B200 box through gpurun; outputs land in gpurun_out/r1/).
set -u
O=gpurun_out/r1; mkdir -p $O
timeout 400 python bench.py --steps 20 --warmup 5 --op-table $O/optable_squeezesegv2_r1.json > $O/bench_r1.json 2> $O/bench_r1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_r1.json 2> $O/bench_reference_r1.err
: > $O/bench_other_workloads_r1.jsonl
for w in darknet21_kitti_64x2048_b32 darknet53_kitti_64x2048_b16 squeezesegv2_nuscenes_32x1024_b32; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --op-table $O/optable_${w%%_*}_${w#*_}.json 2>/dev/null | tail -1 >> $O/bench_other_workloads_r1.jsonl
done
for w in projection_kitti_64x2048_b64 darknet53_projection_64x2048_b64; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/bench_other_workloads_r1.jsonl
done
# launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launches_r1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
# full capture of the three heaviest kernels of one forward (conv14+head, fire13 expand, first CAM)
timeout 400 ncu --set full --clock-control none -k regex:'conv_tc_kernel|cam_kernel' --launch-skip 35 -c 38 --csv --page raw \
  --log-file $O/ncu_full_r1_forward.csv python tools/ncu_forward.py > $O/ncu_full.log 2>&1
ls -la $O
