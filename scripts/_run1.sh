timeout 900 python -m pytest tests/test_gpu_index.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
QK_APS_TRACE=1 timeout 600 python scripts/aps_probe.py 10000000 2>&1 | grep -v "pseudo active 1024 -> 1024 ([0-9.]* ms)$" | tail -22
