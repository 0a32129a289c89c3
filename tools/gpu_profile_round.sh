#!/bin/bash
# Evidence pass of a round (run through gpurun): bench lines, ncu launch list of the bench command, one ncu --set full capture of the
# recurrence kernels, the schedule trace.  A number printed by a run under ncu is never a bench value.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --batch 256 --no-cpu > gpurun_out/bench_cfg2_B256.json 2> gpurun_out/bench_cfg2_B256.err; echo "bench B256 rc=$?"
timeout 600 python bench.py --seconds 8 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_sustained_8s.json 2> gpurun_out/bench_cfg2_sustained.err; echo "sustained rc=$?"
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-cuda-graph > gpurun_out/launch_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blstm_rec -s 0 -c 2 -o gpurun_out/prof_blstm_rec python tools/profile_once.py 128 > gpurun_out/ncu_blstm.log 2>&1; echo "ncu full rc=$?"
timeout 200 python tools/blstm_sched.py 128 > gpurun_out/blstm_sched.txt 2>&1
timeout 200 python tools/blstm_bench.py > gpurun_out/blstm_layer_times.txt 2>&1
for f in bench_cfg2 bench_cfg2_B256 bench_cfg2_sustained_8s; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d.get('blstm_tc_util_pct'), d['clocks'])
PY
done
