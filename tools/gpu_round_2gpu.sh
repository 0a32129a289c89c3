#!/bin/bash
# 2 GPUs: the bucketed gradient all-reduce captured into the step graph (default) vs one flat all-reduce after backward.
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_cfg2_2gpu.json 2> gpurun_out/bench_cfg2_2gpu.err
echo "2gpu rc=$?"; tail -5 gpurun_out/bench_cfg2_2gpu.err | cut -c1-300; cut -c1-400 gpurun_out/bench_cfg2_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --flat-allreduce > gpurun_out/bench_cfg2_2gpu_flat.json 2> gpurun_out/bench_cfg2_2gpu_flat.err
echo "2gpu flat rc=$?"; tail -3 gpurun_out/bench_cfg2_2gpu_flat.err | cut -c1-300; cut -c1-400 gpurun_out/bench_cfg2_2gpu_flat.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_cfg2_1gpu_same_box.json 2> /dev/null; cut -c1-300 gpurun_out/bench_cfg2_1gpu_same_box.json
