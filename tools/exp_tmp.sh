for i in 1 2; do python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "dataset_normalize" 2>&1 | grep -E "assert 0|passed|failed"; done
echo "--- old library"
cp adaptive-multispeaker-separation_b200/libamss_b200.so /tmp/new.so; cp tools/libamss_old.so.keep adaptive-multispeaker-separation_b200/libamss_b200.so
for i in 1 2; do python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "dataset_normalize" 2>&1 | grep -E "assert 0|passed|failed"; done
cp /tmp/new.so adaptive-multispeaker-separation_b200/libamss_b200.so
