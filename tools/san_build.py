import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from optix_raytracer_b200 import host
os.environ["B200RT_HIERARCHY"] = "lbvh"
ctx = host.Context(0)
rng = np.random.default_rng(1)
for n in (1, 2, 3, 31, 32, 33, 64, 100, 3000):
    tris = (rng.random((n, 1, 3), dtype=np.float32) * 10 + (rng.random((n, 3, 3), dtype=np.float32) - 0.5)).astype(np.float32)
    acc = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
    torch.cuda.synchronize()
    print("built", n, acc.info().num_nodes, flush=True)
