timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_gpu_index.py -m gpu -x -q 2>&1 | tail -2
bash scripts/gpu_ab.sh i128+ncu i256+ncu:QK_SEED_SAMPLE=256 i512:QK_SEED_SAMPLE=512
