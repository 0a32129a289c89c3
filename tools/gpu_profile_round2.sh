#!/bin/bash
# Evidence pass after the sparse-kernel / GEMM-epilogue work (run through gpurun): all five bench configurations, launch lists of
# configs 2 and 4, one ncu --set full capture of the sparse kernels.
mkdir -p gpurun_out
for c in 2 1 3 4 5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg $c rc=$?"
done
timeout 600 python bench.py --seconds 8 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_sustained_8s.json 2> gpurun_out/bench_cfg2_sustained.err; echo "sustained rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-cuda-graph > gpurun_out/launch_bench.log 2>&1; echo "launch list cfg2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-cuda-graph > gpurun_out/launch_bench4.log 2>&1; echo "launch list cfg4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:synthesis_fwd_fs|sparse_filter_grad_fs|synthesis_bwd_vals_fs" -c 4 -o gpurun_out/prof_sparse python tools/bench_sparse.py 32 > gpurun_out/ncu_sparse.log 2>&1; echo "ncu sparse rc=$?"
python tools/bench_sparse.py 32 > gpurun_out/sparse_kernel_times.txt 2>&1
for c in 2 1 3 4 5; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg$c.json').read().strip().splitlines()[-1])
print('cfg $c', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d.get('final_result'))
PY
done
