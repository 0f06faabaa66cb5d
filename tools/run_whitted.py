import pathlib, sys
import torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host
from tests import common
ctx = host.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "opaque"
sc = common.duck_scene() if which == "opaque" else common.duck_alpha_scene(1 if which == "mask" else 2)
if which == "raycast":
    rc = host.Raycaster(ctx, common.duck_scene()); rc.buffer_rays(1040)
    for i in range(4):
        rc.launch(want_ext=False)
else:
    mv = host.MeshViewer(ctx, sc, 1920, 1080)
    for i in range(4):
        mv.launch_subframe(i)
torch.cuda.synchronize()
