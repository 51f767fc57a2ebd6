mkdir -p gpurun_out/s30
for v in 1 0; do
timeout 200 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --opt tc_head_vstream=$v --op-table gpurun_out/s30/optable$v.json > gpurun_out/s30/bench$v.json 2>gpurun_out/s30/bench$v.err; tail -3 gpurun_out/s30/bench$v.err
python -c "
import json; d=json.load(open('gpurun_out/s30/bench$v.json')); print('ssv2 vstream=$v', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s30/optable$v.json'))
for o in t['ops']:
  if 'x20_' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
done
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s30/pytest.log 2>&1; tail -5 gpurun_out/s30/pytest.log
