#!/bin/bash
# A/B of (library build, environment) pairs on the bench workload.  usage: tools/ab_env.sh "lib.so VAR=val ..." ...
for spec in "$@"; do
  set -- $spec; lib=$1; shift
  env "$@" B200RT_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')]
d=json.loads(lines[-1]) if lines else None
print('$spec', 'FAILED' if d is None else '%.1f Mrays/s  %.2f ms/step  trace share %.3f  avg trace launch %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['trace_share_of_step'], d['roofline']['avg_launch_ms']))"
done
