timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python scripts/run_configs.py c3 --tag r02a 2>&1 | tail -2
