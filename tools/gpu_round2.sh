#!/bin/bash
# BLSTM recurrence variants + the step profile + a short bench A/B (NB = 16 two-CTA variant vs NB = 32).
mkdir -p gpurun_out
python tools/blstm_bench.py > gpurun_out/blstm_bench.txt 2>&1; cat gpurun_out/blstm_bench.txt
python tools/blstm_profile.py > gpurun_out/blstm_step_profile.txt 2>&1; tail -20 gpurun_out/blstm_step_profile.txt
for nb in 16 32; do
  AMSS_BLSTM_NB=$nb python bench.py --config 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_nb$nb.json 2> gpurun_out/bench_cfg2_nb$nb.err
  echo "nb=$nb rc=$?"; tail -2 gpurun_out/bench_cfg2_nb$nb.err; cut -c1-400 gpurun_out/bench_cfg2_nb$nb.json
done
