timeout 1200 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -6
