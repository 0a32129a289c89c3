#!/bin/bash
# BLSTM recurrence check: parity test, layer timing, step profile (run through gpurun)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "blstm" -s 2>&1 | grep -E "blstm tc|passed|failed|Error|error" | tee gpurun_out/blstm_test.txt
timeout 200 python tools/blstm_bench.py 2>&1 | tee gpurun_out/blstm_layer_times.txt
timeout 200 python tools/blstm_profile.py 2>&1 | tee gpurun_out/blstm_step_profile.txt
