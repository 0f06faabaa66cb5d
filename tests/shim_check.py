"""Run as a subprocess with B200RT_OPTIX_SHIM=1 (tests/test_gpu_parity.py): oracle/optix_ref/optix_harness.cpp — a host program that
makes the OptiX API calls of the reference samples (optixInit, optixDeviceContextCreate, optixModuleCreate, optixProgramGroupCreate,
optixPipelineCreate, optixAccelBuild / Compact, optixSbtRecordPackHeader, optixLaunch) — runs on the product's optixQueryFunctionTable
(optix_raytracer_b200/optix_shim/libnvoptix.so.1) and must produce exactly what the native C ABI produces."""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host  # noqa: E402
from oracle.optix_ref import backend as ob  # noqa: E402
from tests import common  # noqa: E402

ok, why = ob.available(0)
assert ok, why
assert ob.shim_active()
bctx, sctx = host.Context(0), ob.OptixContext(0)
assert sctx.olib.oref_rtcore_version() == 0


def same(a, b, what):
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"{what}: the OptiX-API path and the native ABI differ"
    print("identical:", what)


# optixPathTracer and optixMultiGPU programs
for mg in (None, (0, 1)):
    b, s = host.PathTracer(bctx, 96, 64, 4, multigpu=mg), host.PathTracer(sctx, 96, 64, 4, multigpu=mg)
    b.sample_groups = 1   # what optixLaunch on the shim uses unless B200RT_SAMPLE_GROUPS says otherwise: the reference's summation order
    for sub in range(2):
        b.launch_subframe(sub)
        s.launch_subframe(sub)
    torch.cuda.synchronize()
    assert float(b.accum.abs().sum()) > 0
    same(b.accum, s.accum, f"Cornell accum (multigpu={mg})")
    same(b.frame, s.frame, f"Cornell frame (multigpu={mg})")
# optixRaycasting, with the alpha-mask any-hit program
b, s = host.Raycaster(bctx, common.duck_alpha_scene(1)), host.Raycaster(sctx, common.duck_alpha_scene(1))
b.buffer_rays(256)
s.buffer_rays(256)
b.launch(want_ext=False)
s.launch(want_ext=False)
torch.cuda.synchronize()
assert float((b.hits[:, 0] >= 0).float().mean()) > 0.1
same(b.hits, s.hits, "optixRaycasting Hit buffer")
same(b.hits_translated, s.hits_translated, "optixRaycasting Hit buffer (translated batch)")
# optixMeshViewer (whitted), BLEND material
sc = common.duck_alpha_scene(2)
b, s = host.MeshViewer(bctx, sc, 160, 120), host.MeshViewer(sctx, sc, 160, 120)
for sub in range(2):
    b.launch_subframe(sub)
    s.launch_subframe(sub)
torch.cuda.synchronize()
same(b.accum, s.accum, "whitted accum")
# imgui_test
cam = host.playground_camera(eye=(0.3, 0.6, -1.2), up=(0.0, 1.0, 0.000073), lookat=(0.0, 0.1, 0.0), fov=50.0)
b, s = host.Playground(bctx, 128, 96, spf=2, rows=12, camera=cam), host.Playground(sctx, 128, 96, spf=2, rows=12, camera=cam)
b.launch_frame(dirty=True)
s.launch_frame(dirty=True)
torch.cuda.synchronize()
same(b.film, s.film, "imgui_test film")
# a pipeline of programs the library has no restatement of is refused, not run wrongly
try:
    sctx.prepare_programs("query_gas")
    raise SystemExit("a pipeline of unknown programs was accepted")
except ob.OptixError as e:
    assert "7800" in str(e), e
print("SHIM OK")
