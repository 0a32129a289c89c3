#!/bin/bash
# 8 GPUs of one box under torchrun (the driver's scaling launch): config 5 end to end with and without the NUMA binding of
# the ranks, then config 2.
mkdir -p gpurun_out
export TORCH_NCCL_ASYNC_ERROR_HANDLING=0
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo.txt; lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/topo.txt
run() {  # name config extra-env
  env $3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --config $2 --steps 30 --warmup 5 --no-cpu > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
print("$1", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("host_numa_node"))
PY
}
run bench_cfg5_8gpu_nobind 5 AMSS_NO_NUMA_BIND=1 29621
run bench_cfg5_8gpu 5 AMSS_X=0 29622
run bench_cfg2_8gpu 2 AMSS_X=0 29623
