#!/usr/bin/env python3
"""Turn the reference's Duck.gltf (SDK/data/Duck/Duck.gltf + Duck0.bin) into a compact test fixture,
tests/golden/duck_mesh.npz, using the package's own glTF loader (optix_raytracer_b200.host.load_gltf).
The GPU box has no /root/reference, so the parity tests and bench read this file instead.
Contents: per-primitive positions / normals / TEXCOORD_0 / u16 indices of mesh 0, the instance's 4x4 node transform,
the mesh's object-space AABB from the accessor min/max (what sutil::Scene uses, Scene.cpp:474-489), the glTF camera
(eye, up, fovY as processGLTFNode derives them, Scene.cpp:166-192) and the base-colour texture box-filtered from
512x512 to 128x128 RGBA8 (a small stand-in for Duck.png; the whitted parity tests only need *a* texture both engines sample)."""
import pathlib, sys
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200.host import load_gltf

sc = load_gltf("/root/reference/SDK/data/Duck/Duck.gltf")
assert len(sc["meshes"]) == 1 and len(sc["meshes"][0]["primitives"]) == 1 and len(sc["instances"]) == 1
p = sc["meshes"][0]["primitives"][0]
inst = sc["instances"][0]
img = sc["images"][0].astype(np.float32).reshape(128, 4, 128, 4, 4).mean(axis=(1, 3))
tex = np.clip(np.rint(img), 0, 255).astype(np.uint8)
cam = sc["cameras"][0]
out = ROOT / "tests" / "golden" / "duck_mesh.npz"
np.savez_compressed(out, positions=p["positions"], normals=p["normals"], indices=p["indices"].astype(np.uint16), texcoords0=p["texcoords"][0],
                    texture_rgba8=tex, cam_eye=np.asarray(cam["eye"], np.float32), cam_up=np.asarray(cam["up"], np.float32),
                    cam_fov_y=np.float32(cam["fov_y"]),
                    transform=inst["transform"], aabb_lo=sc["meshes"][0]["aabb"][0], aabb_hi=sc["meshes"][0]["aabb"][1],
                    world_lo=inst["world_aabb"][0], world_hi=inst["world_aabb"][1])
print("wrote", out, out.stat().st_size, "bytes;", p["positions"].shape, p["indices"].shape, inst["transform"], inst["world_aabb"])
