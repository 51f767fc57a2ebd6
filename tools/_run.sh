mkdir -p gpurun_out/s20
for v in 0 3; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --opt cam_px=$v --op-table gpurun_out/s20/optable$v.json > gpurun_out/s20/bench$v.json 2>gpurun_out/s20/bench$v.err; tail -3 gpurun_out/s20/bench$v.err
python -c "
import json; d=json.load(open('gpurun_out/s20/bench$v.json')); print('ssv2 cam_px=$v', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])
t=json.load(open('gpurun_out/s20/optable$v.json'))
for o in t['ops']:
  if 'cam' in o['op']: print('  ', o['op'], round(o['ms'],4), round(o['frac'],3))
"
done
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s20/pytest.log 2>&1; tail -3 gpurun_out/s20/pytest.log
timeout 300 python bench.py --workload darknet21_kitti_64x2048_b32 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-eval > gpurun_out/s20/bench_dk.json 2>gpurun_out/s20/bench_dk.err; tail -3 gpurun_out/s20/bench_dk.err
python -c "
import json; d=json.load(open('gpurun_out/s20/bench_dk.json')); print('dk21', round(d['value']), round(d['ms_per_step'],4), d.get('clocks'))"
timeout 300 python bench.py --workload squeezesegv2_nuscenes_32x1024_b32 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --no-eval > gpurun_out/s20/bench_nu.json 2>gpurun_out/s20/bench_nu.err; tail -3 gpurun_out/s20/bench_nu.err
python -c "
import json; d=json.load(open('gpurun_out/s20/bench_nu.json')); print('nuscenes', round(d['value']), round(d['ms_per_step'],4), d.get('p50_latency_ms'))"
