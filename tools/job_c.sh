timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
B200RT_BUILD_TIMING=2 timeout 120 python tools/build_once.py 2>&1 | grep "b200rt build\|build 1" | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], 'build', d['config']['bvh_build_ms'], 'nodes/seg', r['nodes_per_segment'], 'tris/seg', r['tris_per_segment'], 'nodes', d['config']['bvh8_nodes'])"
