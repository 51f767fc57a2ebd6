mkdir -p gpurun_out/s19
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --op-table gpurun_out/s19/optable.json > gpurun_out/s19/bench.json 2>gpurun_out/s19/bench.err; tail -3 gpurun_out/s19/bench.err
python -c "
import json; d=json.load(open('gpurun_out/s19/bench.json')); print('ssv2', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s19/optable.json'))
for o in t['ops']:
  if 'x20_' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
for v in 1 0; do
timeout 300 python bench.py --workload darknet21_kitti_64x2048_b32 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-eval --opt pair_s2=$v --op-table gpurun_out/s19/optable_dk$v.json > gpurun_out/s19/bench_dk$v.json 2>gpurun_out/s19/bench_dk$v.err; tail -3 gpurun_out/s19/bench_dk$v.err
python -c "
import json; d=json.load(open('gpurun_out/s19/bench_dk$v.json')); print('dk21 pair_s2=$v', round(d['value']), round(d['ms_per_step'],4), d.get('clocks'))
t=json.load(open('gpurun_out/s19/optable_dk$v.json'))
for o in t['ops']:
  if 's2_' in o['op'] and 'deconv' not in o['op'] or 'x20_' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3), o['bound'])
"
done
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s19/pytest.log 2>&1; tail -3 gpurun_out/s19/pytest.log
