python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "head_backward or bf16" 2>&1 | tail -5
for mb in 0 48 24; do
AMSS_HEAD_BWD_L2_MB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_l2_$mb.json 2> gpurun_out/bench_l2_$mb.err; tail -2 gpurun_out/bench_l2_$mb.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_l2_$mb.json').read().strip().splitlines()[-1])
print('L2 MB $mb:', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
for k in d['kernels'][:8]: print('    ', k['entry'], k['calls_per_step'], k['ms_per_step'])
PY
done
