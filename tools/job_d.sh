timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optix_parity.py -m gpu -q -x -k "bvh or soup or builder or coincident or synthetic or whitted" 2>&1 | tail -2
B200RT_BUILD_TIMING=1 timeout 120 python tools/build_once.py --reps 3 2>&1 | grep "b200rt build\|build 2" | tail -2
python bench.py --workload whitted_duck --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('whitted', '%.1f Mrays/s ms=%.4f' % (d['value'], d['ms_per_step']))"
