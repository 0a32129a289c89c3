#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -15
python tools/profile_kmeans.py 16 3 > gpurun_out/kmeans_time.txt 2>&1; python tools/profile_kmeans.py 64 3 >> gpurun_out/kmeans_time.txt 2>&1; python tools/profile_kmeans.py 32 2 >> gpurun_out/kmeans_time.txt 2>&1; cat gpurun_out/kmeans_time.txt
for c in 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c rc=$?"; tail -3 gpurun_out/bench_cfg$c.err; cut -c1-300 gpurun_out/bench_cfg$c.json
done
