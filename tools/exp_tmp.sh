python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['final_result'])
for k in d['kernels'][:9]: print('    ', k['entry'], k['calls_per_step'], k['ms_per_step'])
PY
timeout 200 python tools/blstm_bench.py 2>&1 | grep "NB=auto"
