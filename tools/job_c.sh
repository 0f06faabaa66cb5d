python -m pytest tests -m gpu -q -x -k "face_culling" 2>&1 | tail -2
B200RT_BUILD_TIMING=2 python tools/build_once.py 2>&1 | grep "b200rt build" | tail -2
for b in 16 18; do B200RT_MORTON_BITS=$b python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('bits $b', d['value'], 'build', d['config']['bvh_build_ms'], 'nodes/seg', r['nodes_per_segment'], 'tris/seg', r['tris_per_segment'], 'passes', r['build']['radix_passes'])"; done
