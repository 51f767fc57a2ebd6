#!/bin/bash
# Regenerates the measured artefacts under profiles/ (run on a B200 box through gpurun; outputs land in gpurun_out/r2/,
# copy them to profiles/ afterwards - tools/ncu_summarize.py turns the .ncu-rep into the committed CSV / JSON).
set -u
O=gpurun_out/r2; mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2.log 2>&1; tail -2 $O/pytest_gpu_r2.log
cp gpurun_out/parity_report.jsonl $O/parity_report_r2.jsonl
timeout 500 python bench.py --steps 20 --warmup 5 --op-table $O/optable_squeezesegv2_kitti_64x2048_b32.json > $O/bench_r2.json 2> $O/bench_r2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_r2.json 2> $O/bench_reference_r2.err
# launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_r2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-eval > $O/ncu_launches.log 2>&1
# full capture of every kernel of one plain forward (second forward of tools/ncu_forward.py: 36 launches)
timeout 600 ncu --set full --clock-control none -k regex:'conv_tc|conv_head|cam|pool_conv|squeeze_upconv|net_input' \
  --launch-skip 36 -c 36 -o $O/ncu_full_r2_forward python tools/ncu_forward.py > $O/ncu_full.log 2>&1
# the report is > 64 MiB (gpurun's limit for what travels back): reduce it here, keep only the summaries
python tools/ncu_summarize.py $O/ncu_full_r2_forward.ncu-rep $O/optable_squeezesegv2_kitti_64x2048_b32.json \
  $O/ncu_full_r2_forward_summary.csv $O/ncu_traffic_r2.json 32 && rm -f $O/ncu_full_r2_forward.ncu-rep
# projection kernels (config 4 shape) 
timeout 300 ncu --set full --clock-control none -k regex:'project_' --launch-skip 2 -c 2 -o $O/ncu_full_r2_projection \
  python tools/projection_run.py 64 2 > $O/ncu_proj.log 2>&1
ncu -i $O/ncu_full_r2_projection.ncu-rep --page raw --csv > $O/ncu_full_r2_projection_raw.csv 2>/dev/null
# compute-sanitizer on the atomic kernels (scatter / resolve / confusion) and the validation kernel
timeout 300 compute-sanitizer --tool memcheck python tools/projection_run.py 4 1 > $O/sanitizer_memcheck_projection_confusion_r2.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python tools/projection_run.py 4 1 > $O/sanitizer_racecheck_projection_confusion_r2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_elementwise.py -q -k "validation or test_step" > $O/sanitizer_memcheck_validation_r2.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_elementwise.py -q -k "validation or test_step" > $O/sanitizer_racecheck_validation_r2.log 2>&1
# (tools/umma_bench.cu, tools/tma_bench.cu: microbenchmarks, built by hand with nvcc - their outputs are committed as they are)
ls -la $O
