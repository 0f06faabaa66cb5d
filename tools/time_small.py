#!/usr/bin/env python3
"""b200rt-only timings of the small-scene launches (Duck raycast, whitted opaque / MASK / BLEND, Cornell) for A/B runs of kernel
variants (B200RT_LIB_PATH=...).  Prints one line of ms per case."""
import pathlib, sys
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host
from tests import common


def cuda_ms(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ctx = host.Context(0)
out = {}
rc = host.Raycaster(ctx, common.duck_scene()); rc.buffer_rays(1040)
out["raycast"] = cuda_ms(lambda: rc.launch(want_ext=False))
rc2 = host.Raycaster(ctx, common.duck_alpha_scene(1)); rc2.buffer_rays(1040)
out["raycast_mask"] = cuda_ms(lambda: rc2.launch(want_ext=False))
for name, sc in (("whitted", common.duck_scene()), ("whitted_mask", common.duck_alpha_scene(1)), ("whitted_blend", common.duck_alpha_scene(2))):
    mv = host.MeshViewer(ctx, sc, 1920, 1080)
    out[name] = cuda_ms(lambda: mv.launch_subframe(5))
    mv.close()
for name, kw in (("pg_ref", {}), ("pg_close", {"eye": (0.5, 0.7, -1.4), "up": (0.0, 1.0, 0.000073), "lookat": (0.0, 0.1, 0.0), "fov": 50.0})):
    cam = host.playground_camera(aperture=0.0, **kw)
    pg = host.Playground(ctx, 1920, 1080, spf=8, rows=132, camera=cam)
    out[name] = cuda_ms(lambda: pg.launch_frame(dirty=True), reps=3, warm=1)
    del pg
pt = host.PathTracer(ctx, 768, 768, 16)
pt.sample_groups = 4
out["cornell_sg4"] = cuda_ms(lambda: pt.launch_subframe(3))
print("  ".join(f"{k} {v:.3f}" for k, v in out.items()))
