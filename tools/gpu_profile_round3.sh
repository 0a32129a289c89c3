#!/bin/bash
# Evidence pass after the k-means series (run through gpurun): ncu --set full capture of the tensor-core k-means pass, launch
# list of config 5, fit timings under both group schedules.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:kmeans_pass_tc" -s 22 -c 2 -o gpurun_out/prof_kmeans3 python tools/profile_kmeans.py 64 3 > gpurun_out/ncu_km3.log 2>&1; echo "ncu kmeans rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --config 5 --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_bench5.log 2>&1; echo "launch list cfg5 rc=$?"
for mb in 96 700; do echo "AMSS_KMEANS_GROUP_MB=$mb"; AMSS_KMEANS_GROUP_MB=$mb python tools/profile_kmeans.py 64 3; done 2>&1 | tee gpurun_out/kmeans_group_ab.txt
