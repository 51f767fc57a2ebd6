mkdir -p gpurun_out/n8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/n8/bench_n8.json 2> gpurun_out/n8/bench_n8.err; tail -3 gpurun_out/n8/bench_n8.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/n8/bench_n8.json') if l.startswith('{')][-1]); print('N=8', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d.get('eval'))"
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q) > gpurun_out/n8/pytest_multi.log 2>&1; tail -3 gpurun_out/n8/pytest_multi.log
