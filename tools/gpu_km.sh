#!/bin/bash
# k-means fit timings + tile profile + the full GPU suite + configs 5, 3 and the default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for a in "16 3" "64 3" "32 2"; do python tools/profile_kmeans.py $a; done 2>&1 | tee gpurun_out/kmeans_fit_times.txt
python tools/kmeans_tile_profile.py 8 3 2>&1 | tee gpurun_out/kmeans_tile_profile.txt
for c in 5 3; do
  python bench.py --config $c > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; echo "cfg$c rc=$?"; cut -c1-260 gpurun_out/bench_cfg$c.json
done
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-260 gpurun_out/bench_cfg2.json
