python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for c in 4 2; do
timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; tail -2 gpurun_out/bench_cfg$c.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg$c.json').read().strip().splitlines()[-1])
print('cfg $c:', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
for k in d['kernels'][:7]: print('    ', k['entry'], k['calls_per_step'], k['ms_per_step'], k.get('achieved'), k.get('unit'))
print('   roofline', d['roofline']['kernel'][:50], d['roofline']['bound'], d['roofline']['achieved'], d['roofline']['frac'])
PY
done
