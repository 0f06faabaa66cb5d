#!/usr/bin/env python3
"""Build the synthetic scene's acceleration structure a few times (for ncu launch lists of the build and for host / device timing).
    python tools/build_once.py [--triangles 50000000] [--reps 2]"""
import argparse, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from optix_raytracer_b200 import host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--triangles", type=int, default=50_000_000)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = host.Context(0, log_level=4 if False else 0)
verts, mats = host.synthetic_mesh(ctx, a.triangles, 0)
bi = ctx.triangle_input(verts, sbt_index=mats, num_sbt=4, vertex_stride=16)
torch.cuda.synchronize()
for k in range(a.reps):
    t0 = time.perf_counter()
    times = ctx.time_accel_build([bi], reps=1, warm=0)
    print(f"build {k}: device {times[0]:.3f} ms, host wall incl. sync {(time.perf_counter() - t0) * 1e3:.3f} ms", flush=True)
