#!/bin/bash
# compute-sanitizer over the smoke launch (Cornell 64x64x4 through the whole wavefront loop: INIT, TRACE, SHADE, RESOLVE, the loop graph)
# plus a small build + ray-buffer query + whitted + playground pass: memcheck (global / local / shared out-of-bounds, misaligned),
# racecheck (shared-memory hazards: the cooperative triangle rounds and the queue-append scratch are the shared-write sites),
# synccheck (barrier misuse), initcheck (reads of uninitialised device memory).  GPU box only.
#   tools/sanitize.sh [outdir]        -> <outdir>/sanitize_<tool>.log, summary on stdout
out=${1:-gpurun_out}; mkdir -p $out
cat > /tmp/sanitize_job.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
from optix_raytracer_b200 import host
from tests import common
g.smoke()
ctx = host.Context(0)
rng = np.random.default_rng(1)
tris = (rng.random((3000, 1, 3), dtype=np.float32) * 10 + (rng.random((3000, 3, 3), dtype=np.float32) - 0.5)).astype(np.float32)
for hier in ("lbvh", "ploc"):
    import os; os.environ["B200RT_HIERARCHY"] = hier
    acc = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
    rays = ctx.to_device(common.random_rays(rng, 20000, [0, 0, 0], [10, 10, 10]))
    ctx.trace_closest(acc, rays); ctx.trace_any(acc, rays)
os.environ.pop("B200RT_HIERARCHY")
pt = host.PathTracer(ctx, 48, 40, 4, multigpu=(0, 1)); pt.sample_groups = 2
pt.launch_subframe(0); pt.launch_subframe(1)
for mode in (0, 1, 2):
    sc = common.duck_scene() if mode == 0 else common.duck_alpha_scene(mode)
    rc = host.Raycaster(ctx, sc); rc.buffer_rays(96); rc.launch(); rc.close()
    mv = host.MeshViewer(ctx, sc, 96, 64); mv.launch_subframe(0); mv.launch_subframe(1); mv.close()
pg = host.Playground(ctx, 64, 48, spf=2, rows=8); pg.launch_frame(dirty=True); pg.launch_frame()
torch.cuda.synchronize()
print("sanitize job done")
PY
rc=0
for tool in memcheck racecheck synccheck initcheck; do
  compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_job.py > $out/sanitize_$tool.log 2>&1
  echo "$tool: exit $? — $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize job done' $out/sanitize_$tool.log | tr '\n' ' ')"
done
