#!/bin/bash
# 8 GPUs of one box: config 2 under torchrun (the driver's scaling launch), then config 5.
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
for c in 2 5; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$c bench.py --gpus 8 --config $c --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_cfg${c}_8gpu.json 2> gpurun_out/bench_cfg${c}_8gpu.err
  echo "cfg$c 8gpu rc=$?"; tail -2 gpurun_out/bench_cfg${c}_8gpu.err | cut -c1-300; cut -c1-330 gpurun_out/bench_cfg${c}_8gpu.json
done
