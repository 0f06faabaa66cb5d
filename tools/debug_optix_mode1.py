import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from optix_raytracer_b200 import host
from oracle.optix_ref import backend as ob
bctx, octx = host.Context(0), ob.OptixContext(0)

def cmp(label, w, h, spl, mg, verts=None, mats=None, subs=2):
    b = host.PathTracer(bctx, w, h, spl, vertices=verts, mat_indices=mats, multigpu=mg)
    o = host.PathTracer(octx, w, h, spl, vertices=verts, mat_indices=mats, multigpu=mg)
    for s in range(subs):
        b.launch_subframe(s); o.launch_subframe(s)
    torch.cuda.synchronize()
    ab = b.accum.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64)
    ao = o.accum.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64)
    d = np.abs(ab - ao).max(axis=1)
    print(f"{label}: mean b200rt {ab.mean():.6f} optix {ao.mean():.6f} rel {abs(ab.mean()-ao.mean())/ao.mean():.2e} rmse {np.sqrt(np.mean((ab-ao)**2)):.3e} "
          f"pixels differing > 0.05: {(d > 0.05).mean():.4f}  bit-identical px {(b.accum.cpu().numpy().view(np.uint32) == o.accum.cpu().numpy().view(np.uint32)).all(axis=-1).mean():.3f}", flush=True)

cmp("cornell mode0 256x256x16", 256, 256, 16, None)
cmp("cornell mode1 256x256x16", 256, 256, 16, (0, 1))
for T in (20_000, 2_000_000, 50_000_000):
    verts, mats = host.synthetic_mesh(bctx, T, 0)
    torch.cuda.synchronize()
    cmp(f"synthetic {T} mode0 512x288x16", 512, 288, 16, None, verts, mats)
    cmp(f"synthetic {T} mode1 512x288x16", 512, 288, 16, (0, 1), verts, mats)
    del verts, mats
    torch.cuda.empty_cache()
