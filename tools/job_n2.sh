timeout 600 python -m pytest tests/test_gpu_async.py -m gpu -q -rs -x 2>&1 | tail -15
