mkdir -p gpurun_out/s16
for v in 3 1; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --opt tc_rtma=$v --op-table gpurun_out/s16/optable$v.json > gpurun_out/s16/bench$v.json 2>gpurun_out/s16/bench$v.err; tail -3 gpurun_out/s16/bench$v.err
python -c "
import json; d=json.load(open('gpurun_out/s16/bench$v.json')); print('rtma=$v', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s16/optable$v.json'))
for o in t['ops']:
  if 'conv' in o['op'] and 'fused' not in o['op'] and 'pool' not in o['op'] and '+' not in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
done
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s16/pytest.log 2>&1; tail -3 gpurun_out/s16/pytest.log
