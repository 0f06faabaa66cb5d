#!/usr/bin/env python3
"""Dump hit records of identical ray batches from OptiX (built-in triangles, oracle/optix_ref query programs) and from b200rt, so that
the distribution of |t_b200rt - t_optix| in ulps can be studied offline against candidate formulas for t (tools/t_formula_study.py).
GPU box only.   python tools/dump_t_compare.py [--rays 262144] [--out gpurun_out/t_compare.npz]"""
import argparse, pathlib, sys
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from optix_raytracer_b200 import host  # noqa: E402
from oracle.optix_ref import backend as ob  # noqa: E402
from tests import common  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1 << 18)
ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "t_compare.npz"))
a = ap.parse_args()
ok, why = ob.available(0)
if not ok:
    print("OptiX unavailable:", why); sys.exit(0)
bctx, octx = host.Context(0), ob.OptixContext(0)
rng = np.random.default_rng(17)
out = {}
# Cornell (GAS)
b, o = host.PathTracer(bctx, 32, 32, 1), host.PathTracer(octx, 32, 32, 1)
rays = common.random_rays(rng, a.rays, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
d = bctx.to_device(rays)
got = host.ext_hits_to_numpy(bctx.trace_closest(b.accel, d)); ref = host.ext_hits_to_numpy(octx.trace_closest(o.accel, d))
out.update({"cornell_rays": rays, "cornell_verts": b.scene["vertices"].astype(np.float32)})
for k in ("t", "prim", "b1", "b2"):
    out[f"cornell_b_{k}"], out[f"cornell_o_{k}"] = got[k], ref[k]
# Duck (IAS, scaled instance)
sc = common.duck_scene()
brc, orc_ = host.Raycaster(bctx, sc), host.Raycaster(octx, sc)
rr = common.random_rays(rng, a.rays, brc.bbmin, brc.bbmax)
d = bctx.to_device(rr)
got = host.ext_hits_to_numpy(bctx.trace_closest(brc.ias, d)); ref = host.ext_hits_to_numpy(octx.trace_closest(orc_.ias, d, is_ias=True))
tris, _ = common.deindex(sc["meshes"][0]["primitives"][0])
out.update({"duck_rays": rr, "duck_tris": tris, "duck_xform": sc["instances"][0]["transform"].astype(np.float32)})
for k in ("t", "prim", "b1", "b2"):
    out[f"duck_b_{k}"], out[f"duck_o_{k}"] = got[k], ref[k]
np.savez_compressed(a.out, **out)
for name in ("cornell", "duck"):
    tb, to = out[f"{name}_b_t"], out[f"{name}_o_t"]
    hit = (tb >= 0) & (to >= 0) & (out[f"{name}_b_prim"] == out[f"{name}_o_prim"])
    u = np.abs(tb[hit].view(np.int32).astype(np.int64) - to[hit].view(np.int32).astype(np.int64))
    print(name, "hits", int(hit.sum()), "bit-identical %.4f  <=1ulp %.4f  <=4ulp %.4f  max %d" % ((u == 0).mean(), (u <= 1).mean(), (u <= 4).mean(), u.max()))
