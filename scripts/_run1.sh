timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_ab.sh m128+ncu m256+ncu:QK_LIB_PATH=$PWD/quake_b200/lib/libquake_b200_m256.so 2>&1 | grep "AB\|refine\|sum"
