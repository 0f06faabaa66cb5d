#!/usr/bin/env python3
"""Minimal driver for ncu: build the synthetic scene and run a few path-tracing steps (no timing passes).
    python tools/profile_step.py --triangles 50000000 --width 1920 --height 1080 --spl 4 --steps 2"""
import argparse, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from optix_raytracer_b200 import host

ap = argparse.ArgumentParser()
ap.add_argument("--triangles", type=int, default=50_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spl", type=int, default=4)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--cornell", action="store_true")
ap.add_argument("--sample-groups", type=int, default=4)
a = ap.parse_args()
ctx = host.Context(0)
if a.cornell:
    pt = host.PathTracer(ctx, a.width, a.height, a.spl)
else:
    verts, mats = host.synthetic_mesh(ctx, a.triangles, 0)
    pt = host.PathTracer(ctx, a.width, a.height, a.spl, vertices=verts, mat_indices=mats, multigpu=(0, 1))
pt.sample_groups = a.sample_groups
torch.cuda.synchronize()
for i in range(a.steps):
    t = time.perf_counter()
    st = pt.launch_subframe(i, collect_stats=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    segs = st.radiance_segments + st.shadow_segments
    print(f"step {i}: {segs} segments, {st.iterations} iterations, {dt*1e3:.1f} ms, {segs/dt/1e6:.1f} Mrays/s")
