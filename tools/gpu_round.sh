#!/bin/bash
# k-means on the tensor cores: parity tests, timing (TC vs SIMT), ncu; configs 3 and 5; BLSTM ncu capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "kmeans or inference or enhance or golden" > gpurun_out/pytest_kmeans.log 2>&1; echo "pytest kmeans rc=$?"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_kmeans.log | tail -15
python tools/profile_kmeans.py 16 3 > gpurun_out/kmeans_time.txt 2>&1; AMSS_KMEANS_SIMT=1 python tools/profile_kmeans.py 16 3 >> gpurun_out/kmeans_time.txt 2>&1
python tools/profile_kmeans.py 64 3 >> gpurun_out/kmeans_time.txt 2>&1; python tools/profile_kmeans.py 32 2 >> gpurun_out/kmeans_time.txt 2>&1; cat gpurun_out/kmeans_time.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -15
for c in 3 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c rc=$?"; tail -3 gpurun_out/bench_cfg$c.err; cut -c1-300 gpurun_out/bench_cfg$c.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_pass -s 4 -c 2 -o gpurun_out/prof_kmeans -f python tools/profile_kmeans.py 8 3 > gpurun_out/ncu_kmeans.log 2>&1
echo "ncu kmeans rc=$?"; tail -2 gpurun_out/ncu_kmeans.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blstm_rec -c 2 -o gpurun_out/prof_blstm -f python tools/profile_once.py 128 > gpurun_out/ncu_blstm.log 2>&1
echo "ncu blstm rc=$?"; tail -2 gpurun_out/ncu_blstm.log | cut -c1-200
