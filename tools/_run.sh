mkdir -p gpurun_out/s17
for v in 3 1; do
timeout 300 python bench.py --workload darknet21_kitti_64x2048_b32 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-eval --opt tc_rtma=$v --op-table gpurun_out/s17/optable_dk$v.json > gpurun_out/s17/bench_dk$v.json 2>gpurun_out/s17/bench_dk$v.err; tail -3 gpurun_out/s17/bench_dk$v.err
python -c "
import json; d=json.load(open('gpurun_out/s17/bench_dk$v.json')); print('dk21 rtma=$v', round(d['value']), round(d['ms_per_step'],4), d.get('clocks'))
t=json.load(open('gpurun_out/s17/optable_dk$v.json'))
for o in t['ops']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3), o['bound'])
"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval > gpurun_out/s17/bench.json 2>gpurun_out/s17/bench.err; python -c "
import json; d=json.load(open('gpurun_out/s17/bench.json')); print('ssv2', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])"
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s17/pytest.log 2>&1; tail -3 gpurun_out/s17/pytest.log
