python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x -k "stft or istft or finetun or FineTune or inference" 2>&1 | tail -3
for v in 0 1; do
if [ $v = 1 ]; then export AMSS_ISTFT_PER_BLOCK=1; fi
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_$v.json 2> gpurun_out/bench_cfg5_$v.err; tail -1 gpurun_out/bench_cfg5_$v.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg5_$v.json').read().strip().splitlines()[-1])
print('per-block=$v:', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
for k in d['kernels'][:6]: print('    ', k['entry'], k['calls_per_step'], k['ms_per_step'], k.get('achieved'), k.get('unit'))
PY
done
