#!/bin/bash
# A/B of library builds (in-tree + gpurun_variants/*.so) over the bench workloads.  usage: tools/ab_all.sh "synthetic cornell duck_raycast" [steps]
wls=${1:-"synthetic cornell duck_raycast whitted_duck"}; steps=${2:-4}
for lib in optix_raytracer_b200/libb200rt.so gpurun_variants/*.so; do
  [ -f "$lib" ] || continue
  for w in $wls; do
    B200RT_LIB_PATH=$PWD/$lib timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')]
d=json.loads(lines[-1]) if lines else None
r=(d or {}).get('roofline') or {}
print('%-44s %-13s' % ('$lib', '$w'), 'FAILED' if d is None else '%9.1f Mrays/s  %8.3f ms/step  e2e %9.1f  nodes/seg %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('nodes_per_segment')))"
  done
done
