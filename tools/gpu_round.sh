#!/bin/bash
# One gpurun call: GPU test suite, BLSTM recurrence variants, the five bench configs (JSON lines under gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python tools/blstm_bench.py > gpurun_out/blstm_bench.txt 2>&1; cat gpurun_out/blstm_bench.txt
timeout 300 python tools/blstm_profile.py > gpurun_out/blstm_step_profile.txt 2>&1; tail -12 gpurun_out/blstm_step_profile.txt
for nb in 16 32; do
  AMSS_BLSTM_NB=$nb timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_nb$nb.json 2> gpurun_out/bench_cfg2_nb$nb.err
  echo "nb=$nb rc=$?"; tail -2 gpurun_out/bench_cfg2_nb$nb.err; cut -c1-300 gpurun_out/bench_cfg2_nb$nb.json
done
for c in 2 1 4 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c rc=$?"; tail -3 gpurun_out/bench_cfg$c.err; cut -c1-600 gpurun_out/bench_cfg$c.json
done
