timeout 400 python -m pytest tests/test_gpu_knobs.py -m gpu -x -q 2>&1 | tail -4
