import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from optix_raytracer_b200 import host
from oracle import pyoracle as orc
from tests import common
ctx = host.Context(0, log_level=4)
pt = host.PathTracer(ctx, 32, 32, 1, compact=False)
torch.cuda.synchronize()
blob = pt.accel.buf.cpu().numpy()
gas = common.decode_gas(blob)
print({k: v for k, v in gas.items() if k not in ('nodes', 'tris')})
for i in range(gas['num_nodes']):
    raw = gas['nodes'][i]
    print(i, 'P', raw[0:12].view(np.float32), 'e', raw[12:15], 'imask', bin(raw[15]), 'cb,tb', raw[16:24].view(np.uint32), 'meta', [hex(x) for x in raw[24:32]])
    print('   qlo', raw[32:56].reshape(3, 8).tolist(), 'qhi', raw[56:80].reshape(3, 8).tolist())
print('tri prims', gas['tris'][:, 0, 3].view(np.uint32), 'ord', gas['tris'][:, 2, 3].view(np.uint32))
try:
    print('validate', common.validate_gas(gas))
except AssertionError as e:
    print('VALIDATION FAILED', e)
sc = pt.scene
scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
rng = np.random.default_rng(1)
rays = common.random_rays(rng, 2000, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
got = host.ext_hits_to_numpy(ctx.trace_closest(pt.accel, ctx.to_device(rays)))
ref = scene.trace(rays)
bad = np.nonzero((got['prim'] != ref['prim']))[0]
print('bad', bad.size, 'of', rays.shape[0])
for i in bad[:8]:
    print(i, rays[i], 'got', got['t'][i], got['prim'][i], 'ref', ref['t'][i], ref['prim'][i])
