mkdir -p gpurun_out/s36
for v in 1 0; do
timeout 120 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-eval --opt tc_qgroup=$v --op-table gpurun_out/s36/optable$v.json > gpurun_out/s36/bench$v.json 2>gpurun_out/s36/bench$v.err; tail -3 gpurun_out/s36/bench$v.err
python -c "
import json; d=json.load(open('gpurun_out/s36/bench$v.json')); print('ssv2 qgroup=$v', round(d['value']), round(d['ms_per_step'],4), d['p50_latency_ms'])"
done
(timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/s36/pytest.log 2>&1; tail -4 gpurun_out/s36/pytest.log
