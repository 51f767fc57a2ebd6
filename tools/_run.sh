mkdir -p gpurun_out/s26
for sfx in "" _c4; do
PCLS_LIB_SUFFIX=$sfx timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --op-table gpurun_out/s26/optable$sfx.json > gpurun_out/s26/bench$sfx.json 2>gpurun_out/s26/bench$sfx.err; tail -3 gpurun_out/s26/bench$sfx.err
python -c "
import json; d=json.load(open('gpurun_out/s26/bench$sfx.json')); print('ssv2 lib=$sfx', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s26/optable$sfx.json'))
for o in t['ops']:
  if 'cam' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
done
