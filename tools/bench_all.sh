#!/bin/bash
# Both bench arms on every BASELINE workload (C1-C5), one JSON line each -> gpurun_out/<tag>_<workload>_<impl>.json
# usage: tools/bench_all.sh <tag> [steps] [warmup]
tag=${1:-r02}; steps=${2:-6}; warm=${3:-3}
mkdir -p gpurun_out
for w in synthetic cornell duck_raycast whitted_duck playground; do
  for impl in reference b200rt; do
    extra="--no-cpu-baseline"; [ "$w" = synthetic ] && [ "$impl" = b200rt ] && extra=""
    python bench.py --workload $w --impl $impl --steps $steps --warmup $warm $extra > gpurun_out/${tag}_${w}_${impl}.json 2> gpurun_out/${tag}_${w}_${impl}.err
    echo "$w $impl rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${w}_${impl}.json').read().strip().splitlines()[-1])
    print(d.get('reference_class','-'), 'value=%.1f e2e=%.1f ms=%.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e:
    print('unparsed', e)
PY
)"
  done
done
