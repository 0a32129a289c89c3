#!/bin/bash
# One gpurun call: GPU test suite, BLSTM recurrence variants, the five bench configs, a sustained run, ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python tools/blstm_bench.py > gpurun_out/blstm_bench.txt 2>&1; cat gpurun_out/blstm_bench.txt
for c in 2 1 4 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c rc=$?"; tail -3 gpurun_out/bench_cfg$c.err; cut -c1-400 gpurun_out/bench_cfg$c.json
done
timeout 600 python bench.py --config 2 --seconds 8 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_sustained.json 2> gpurun_out/bench_cfg2_sustained.err
echo "sustained rc=$?"; cut -c1-300 gpurun_out/bench_cfg2_sustained.json
timeout 600 python bench.py --config 2 --batch 256 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2_b256.json 2> gpurun_out/bench_cfg2_b256.err
echo "b256 rc=$?"; cut -c1-300 gpurun_out/bench_cfg2_b256.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --config 2 --steps 2 --warmup 3 --no-cpu --no-cuda-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
