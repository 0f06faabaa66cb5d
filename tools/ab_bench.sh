#!/bin/bash
# A/B of library builds on the bench workload: the in-tree library, then every gpurun_variants/*.so (tools/ab_variants.sh also times the
# small scenes).  Prints Mrays/s, ms per step, trace share, accel build ms.
for lib in optix_raytracer_b200/libb200rt.so gpurun_variants/*.so; do
  [ -f "$lib" ] || continue
  B200RT_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')]
d=json.loads(lines[-1]) if lines else None
print('$lib', 'FAILED' if d is None else '%.1f Mrays/s  %.2f ms/step  e2e %.1f  trace share %.3f  nodes/seg %.2f  accel build %.2f ms  setup %.1f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['trace_share_of_step'], d['roofline']['nodes_per_segment'], d['config']['bvh_build_ms'], d['config']['scene_setup_cold_ms']))"
done
