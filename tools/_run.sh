mkdir -p gpurun_out/s27
for sfx in "" _th; do
PCLS_LIB_SUFFIX=$sfx timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --op-table gpurun_out/s27/optable$sfx.json > gpurun_out/s27/bench$sfx.json 2>gpurun_out/s27/bench$sfx.err; tail -3 gpurun_out/s27/bench$sfx.err
python -c "
import json; d=json.load(open('gpurun_out/s27/bench$sfx.json')); print('ssv2 lib=$sfx', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s27/optable$sfx.json'))
for o in t['ops']:
  if 'cam' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
done
rm -f gpurun_out/parity_report.jsonl
(PCLS_LIB_SUFFIX=_th PCLS_NVCC_FLAGS="-DPCLS_CAM_TANH=1" timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/s27/pytest_th.log 2>&1; tail -15 gpurun_out/s27/pytest_th.log
cp gpurun_out/parity_report.jsonl gpurun_out/s27/parity_th.jsonl
