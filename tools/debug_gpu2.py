import sys, numpy as np, torch, pathlib
sys.path.insert(0, '/root/repo')
from optix_raytracer_b200 import _lib
_lib.LIB_PATH = pathlib.Path('/root/repo/gpurun_dbg_libb200rt.so')
from optix_raytracer_b200 import host
ctx = host.Context(0)
pt = host.PathTracer(ctx, 32, 32, 1, compact=False)
rays = np.array([[1.9722353e+02,4.0069336e+02,3.2397247e+02,0.01,0,0,1,1e16]], np.float32)
got = host.ext_hits_to_numpy(ctx.trace_closest(pt.accel, ctx.to_device(rays)))
torch.cuda.synchronize()
print(got)
