#!/bin/bash
# 2 GPUs of one box: config 2 (training, bucketed all-reduce) and config 5 (inference replicas) under torchrun, then 1 GPU on the same box.
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
for c in 2 5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$c bench.py --gpus 2 --config $c --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_cfg${c}_2gpu.json 2> gpurun_out/bench_cfg${c}_2gpu.err
  echo "cfg$c 2gpu rc=$?"; tail -3 gpurun_out/bench_cfg${c}_2gpu.err | cut -c1-300; cut -c1-330 gpurun_out/bench_cfg${c}_2gpu.json
  python bench.py --gpus 1 --config $c --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_cfg${c}_1gpu_same_box.json 2> /dev/null; cut -c1-330 gpurun_out/bench_cfg${c}_1gpu_same_box.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-200
