python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x 2>&1 | tail -8
python -m pytest tests/test_abi.py -q 2>&1 | tail -2
for v in 1 0; do
AMSS_TRAIN_V_FP32=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; tail -2 gpurun_out/bench_v$v.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().splitlines()[-1])
print('V fp32=$v:', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['final_result'])
for k in d['kernels'][:9]: print('    ', k['entry'], k['calls_per_step'], k['ms_per_step'])
PY
done
