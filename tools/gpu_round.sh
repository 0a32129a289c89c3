#!/bin/bash
# One gpurun call: GPU test suite + the five bench configs (JSON lines under gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for c in 2 1 4 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c rc=$?"; tail -3 gpurun_out/bench_cfg$c.err; cut -c1-700 gpurun_out/bench_cfg$c.json
done
