import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from optix_raytracer_b200 import host
from oracle.optix_ref import backend as ob
from tests import common
bctx, octx = host.Context(0), ob.OptixContext(0)
sc = common.duck_scene()
brc = host.Raycaster(bctx, sc)
n = brc.buffer_rays(1040)
brc.launch(); torch.cuda.synchronize(); print("b200rt raycast ok", flush=True)
orc = host.Raycaster(octx, sc)
orc.buffer_rays(1040)
torch.cuda.synchronize(); print("optix setup ok", flush=True)
# 1. my query programs on the OptiX GAS alone
r = common.random_rays(np.random.default_rng(0), 1 << 16, sc["meshes"][0]["aabb"][0], sc["meshes"][0]["aabb"][1])
e = octx.trace_closest(orc.mesh_accels[0], bctx.to_device(r), is_ias=False); print("optix query on GAS ok", int((e[:, 1] >= 0).sum()), flush=True)
e = octx.trace_closest(orc.ias, brc.rays, is_ias=True); print("optix query on IAS ok", int((e[:, 1] >= 0).sum()), flush=True)
orc.launch(want_ext=False); torch.cuda.synchronize(); print("optix raycast ok", flush=True)
