mkdir -p gpurun_out/s35
timeout 200 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --op-table gpurun_out/s35/optable.json > gpurun_out/s35/bench.json 2>gpurun_out/s35/bench.err; tail -3 gpurun_out/s35/bench.err
python -c "
import json; d=json.load(open('gpurun_out/s35/bench.json')); print('ssv2', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s35/optable.json'))
for o in t['ops']:
  if 'x20_' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s35/pytest.log 2>&1; tail -4 gpurun_out/s35/pytest.log
