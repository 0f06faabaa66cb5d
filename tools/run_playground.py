import pathlib, sys
import torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host
ctx = host.Context(0)
kw = {} if (len(sys.argv) < 2 or sys.argv[1] == "ref") else {"eye": (0.5, 0.7, -1.4), "up": (0.0, 1.0, 0.000073), "lookat": (0.0, 0.1, 0.0), "fov": 50.0}
cam = host.playground_camera(aperture=0.0, **kw)
pg = host.Playground(ctx, 1920, 1080, spf=8, rows=132, camera=cam)
for i in range(3):
    pg.launch_frame(dirty=True)
torch.cuda.synchronize()
