#!/bin/bash
# k-means fit timings + its parity tests + configs 3 and 5
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "kmeans or separate or infer" 2>&1 | tail -3
for a in "16 3" "64 3" "32 2"; do python tools/profile_kmeans.py $a; done 2>&1 | tee gpurun_out/kmeans_fit_times.txt
for c in 5 3; do
  python bench.py --config $c --no-cpu > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg$c rc=$?"; cut -c1-260 gpurun_out/bench_cfg$c.json
done
