#!/bin/bash
# Final evidence pass: full GPU suite, smoke, five bench configurations, sustained config 2, k-means fit timings + tile profile.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in 2 1 3 4 5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg $c rc=$?"
done
timeout 600 python bench.py --seconds 8 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_sustained_8s.json 2> gpurun_out/bench_cfg2_sustained.err; echo "sustained rc=$?"
for a in "16 3" "64 3" "32 2"; do python tools/profile_kmeans.py $a; done 2>&1 | tee gpurun_out/kmeans_fit_times.txt
python tools/kmeans_tile_profile.py 8 3 > gpurun_out/kmeans_tile_profile.txt 2>&1
for c in 2 1 3 4 5; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg$c.json').read().strip().splitlines()[-1])
print('cfg $c', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['roofline'].get('frac'), d.get('blstm_tc_util_pct'))
PY
done
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg2_sustained_8s.json').read().strip().splitlines()[-1])
print('sustained', round(d['value'],1), d['clocks'])
PY
