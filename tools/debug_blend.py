import sys, pathlib, ctypes as C
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host
from oracle import pyoracle as orc
from tests import common
ctx = host.Context(0)
for ds in (False, True):
    sc = common.duck_scene(textured=False)
    sc["materials"][0].update({"alpha_mode": 2, "double_sided": ds, "base_color": [1.0, 0.9, 0.8, 0.6]})
    w, h = 160, 120
    mv = host.MeshViewer(ctx, sc, w, h)
    prim = sc["meshes"][0]["primitives"][0]
    tris, nrm = common.deindex(prim)
    scene = orc.Scene(tris, None, instances=[sc["instances"][0]["transform"][:3, :].reshape(12)], geom_flags=4 if ds else 0)
    p = orc.WhittedParams()
    p.width, p.height = w, h
    for k in ("eye", "U", "V", "W", "miss_color"):
        setattr(p, k, getattr(mv.params, k))
    m = sc["materials"][0]
    p.base_color = (C.c_float * 4)(*m["base_color"]); p.metallic, p.roughness = m["metallic"], m["roughness"]
    p.emissive = (C.c_float * 3)(*m["emissive_factor"]); p.alpha_mode = 2
    lights = mv.d_lights.cpu().numpy().tobytes()
    mv.launch_subframe(0); torch.cuda.synchronize()
    p.subframe_index = 0
    accum, frame, nrays = scene.whitted(p, lights, normals=nrm)
    got = mv.accum.cpu().numpy()
    d = (got.view(np.uint32) != accum.view(np.uint32)).any(-1)
    print("double_sided", ds, "rays", nrays, "differing pixels", d.sum(), "of covered", (accum[..., :3] != np.float32(0.1)).any(-1).sum())
    ys, xs = np.nonzero(d)
    for y, x in list(zip(ys, xs))[:8]:
        print(y, x, got[y, x], accum[y, x])
    print("max abs diff", np.abs(got - accum).max())
