timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -4
