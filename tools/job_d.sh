timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bvh or soup or builder or coincident or synthetic" 2>&1 | tail -2
B200RT_BUILD_TIMING=1 timeout 120 python tools/build_once.py --reps 3 2>&1 | grep "b200rt build\|build 2" | tail -2
