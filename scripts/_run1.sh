timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -2
bash scripts/gpu_ab.sh fold fold_f32:QK_REFRESH_FIRST=32 fold_f24:QK_REFRESH_FIRST=24
