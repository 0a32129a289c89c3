timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -x -k "blstm" -s 2>&1 | grep -E "blstm tc|passed|failed|Error|error"
timeout 100 python tools/blstm_profile.py 16 128 2>&1 | grep -v "bwd step"
timeout 200 python tools/blstm_bench.py 2>&1 | grep "NB=auto"
timeout 200 python tools/blstm_sched.py 128 2>&1 | grep -v "  cluster"
