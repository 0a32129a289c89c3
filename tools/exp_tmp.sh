python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "dataset_normalize" -s 2>&1 | grep -E "dataset_normalize step|assert|passed|failed"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('blstm_tc_util_pct'))
for k in d['kernels'][:8]: print(k['entry'], k['ms_per_step'])
print(d.get('parity'))
PY
timeout 100 python tools/blstm_profile.py 16 128 2>&1 | tee gpurun_out/blstm_step_profile.txt | grep "bwd step"
