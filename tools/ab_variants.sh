#!/bin/bash
# A/B runs of kernel variants built into gpurun_variants/*.so (see optix_raytracer_b200/csrc/Makefile: OUT=, BUILD=, EXTRA=).
#   tools/ab_variants.sh [bench args...]      prints: variant  Mrays/s  ms/step  nodes/seg  tris/seg
for lib in optix_raytracer_b200/libb200rt.so gpurun_variants/*.so; do
  [ -f "$lib" ] || continue
  B200RT_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')]
d=json.loads(lines[-1]) if lines else None
print('$lib', 'FAILED' if d is None else '%.1f Mrays/s  %.2f ms/step  nodes/seg %.2f  tris/seg %.2f  trace share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['nodes_per_segment'], d['roofline']['tris_per_segment'], d['roofline']['trace_share_of_step']))"
done
