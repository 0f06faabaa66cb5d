#!/usr/bin/env python3
"""Where the end-to-end time of a small-scene step goes (whitted_duck / duck_raycast / playground): per step, CUDA events behind the
launch and behind the device->host copy of the frame, and the host time of the two calls.  GPU box only.
    python tools/e2e_probe.py --workload whitted_duck --impl b200rt|reference"""
import argparse, sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from optix_raytracer_b200 import host

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="whitted_duck")
ap.add_argument("--impl", default="b200rt")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--no-flush", action="store_true")
a0 = ap.parse_args()
a = bench.parse(["--workload", a0.workload, "--impl", a0.impl, "--steps", str(a0.steps)])
torch.cuda.set_device(0)
if a0.impl == "reference":
    from oracle.optix_ref import backend as ob
    ctx = ob.OptixContext(0)
else:
    ctx = host.Context(0)
job = bench.make_job(a, ctx, 0, 1, host)
frames = job.frame if isinstance(job.frame, tuple) else (job.frame,)
h_frames = [torch.empty(f.shape, dtype=f.dtype).pin_memory() for f in frames]
print("frames:", [(tuple(f.shape), f.dtype, f.is_contiguous()) for f in frames])
flush = None if a0.no_flush else torch.empty(int(torch.cuda.get_device_properties(0).L2_cache_size * 1.5) // 4, dtype=torch.float32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
for sub in range(5):
    job.step(sub)
torch.cuda.synchronize()
rows, host_launch, host_copy = [], [], []
for sub in range(5, 5 + a0.steps):
    if flush is not None:
        flush.fill_(1.0)
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    t0 = time.perf_counter()
    job.step(sub)
    t1 = time.perf_counter()
    e1.record()
    for h, f in zip(h_frames, frames):
        h.copy_(f, non_blocking=True)
    t2 = time.perf_counter()
    e2.record()
    rows.append((e0, e1, e2)); host_launch.append(t1 - t0); host_copy.append(t2 - t1)
torch.cuda.synchronize()
l = np.array([e0.elapsed_time(e1) for e0, e1, e2 in rows]); c = np.array([e1.elapsed_time(e2) for e0, e1, e2 in rows])
print(f"{a0.workload} {a0.impl}: launch {np.median(l):.4f} ms (min {l.min():.4f}), copy {np.median(c):.4f} ms, host launch call {1e3 * np.median(host_launch):.4f} ms, "
      f"host copy call {1e3 * np.median(host_copy):.4f} ms")
