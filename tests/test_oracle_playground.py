"""imgui_test ("playground", BASELINE.json configs[3]) on the CPU: the oracle restatement and the host-side mirrors against golden
vectors written by the reference's own classes (tools/make_golden.py through oracle/ref_shim_pg.cpp), plus size-independent
properties of a rendered frame."""
import ctypes as C
import json
import pathlib
import struct

import numpy as np
import pytest

from oracle import pyoracle as orc

ROOT = pathlib.Path(__file__).resolve().parents[1]
KAT = json.loads((ROOT / "tests" / "golden" / "kat.json").read_text())


def _ulps(a_bits, b):
    a = np.array(a_bits, np.uint32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b).max()


def test_host_camera_and_lights_are_byte_identical_to_the_reference_objects():
    torch = pytest.importorskip("torch")
    from optix_raytracer_b200 import host
    for c in KAT["playground_cameras"]:
        got = host.playground_camera(c["eye"], c["up"], c["lookat"], c["aperture"], c["fd"], c["fov"], c["ortho"])
        assert got.hex() == c["bytes"]
    for l in KAT["playground_lights"]:
        got = host.playground_light(l["kind"], l["a"], l["lumi"], l["scalar"])
        n = 36 if l["kind"] == 0 else 40  # PointLight leaves the last 4 bytes of the union untouched
        assert got[:n].hex() == l["bytes"][:2 * n] and got[40:].hex() == l["bytes"][80:]
    ref_default = "".join(l["bytes"] for l in KAT["playground_lights"][:4])
    assert host.playground_default_lights().hex() == ref_default
    pl = KAT["playground_layout"]
    assert C.sizeof(host.PGParams) == pl["Params"]
    for name in ("image_width", "image_height", "samples_per_frame", "camera", "dt", "dirty", "image", "film", "tfactor", "handle", "normals",
                 "vertices", "mat_indices", "nmat_indices", "lights", "nlights", "materials", "nmaterials"):
        assert getattr(host.PGParams, name).offset == pl[name], name
    assert (pl["sizeof_Camera"], pl["sizeof_LightVariant"], pl["sizeof_DiffuseMaterial"]) == (92, 44, 12)
    assert host.playground_default_materials().shape == (28, 3)


def test_oracle_ray_generation_and_lights_track_the_reference_host_evaluation():
    """Camera::compute_ray / LightVariant::wi are __host__ __device__ in the reference: its host compiler evaluates them without
    fma contraction, the contract places the fmas nvcc would — same RNG consumption, values within a few ulp.  Where two rnd(seed)
    calls are arguments of one make_float2 / make_float3 the host compiler (gcc) evaluates them right to left while the device
    code — the PTX nvcc makes of optixTriangle.cu draws x first (oracle/_ref/optixTriangle.ptx) — goes left to right, so for those
    cases (aperture > 0, jittered lights) only the RNG consumption is compared here; the values are compared on the GPU against the
    reference program running on OptiX (tests/test_gpu_optix_parity.py)."""
    L = orc.lib()
    L.orc_playground_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    L.orc_playground_light.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    for c in KAT["playground_cameras"]:
        cam = np.frombuffer(bytes.fromhex(c["bytes"]), np.uint8).copy()
        for r in c["rays"]:
            seed = C.c_uint32(r["seed"])
            org, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
            L.orc_playground_ray(cam.ctypes.data, r["ix"], r["iy"], r["w"], r["h"], C.byref(seed), org.ctypes.data, d.ctypes.data)
            assert seed.value == r["seed_after"]
            ref_o = np.array(r["org_bits"], np.uint32).view(np.float32)
            ref_d = np.array(r["dir_bits"], np.uint32).view(np.float32)
            if c["aperture"] == 0.0 or c["ortho"]:
                assert np.allclose(org, ref_o, rtol=0, atol=2e-6 * max(1.0, np.abs(ref_o).max())) and np.allclose(d, ref_d, rtol=0, atol=3e-7)
    for l in KAT["playground_lights"]:
        lb = np.frombuffer(bytes.fromhex(l["bytes"]), np.uint8).copy()
        seed = C.c_uint32(l["seed"])
        wi, lumi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        p = np.array(l["p"], np.float32)
        L.orc_playground_light(lb.ctypes.data, p.ctypes.data, C.byref(seed), wi.ctypes.data, lumi.ctypes.data)
        assert seed.value == l["seed_after"]
        assert _ulps(l["lumi_bits"], lumi) == 0
        if l["kind"] == 0:
            assert _ulps(l["wi_bits"], wi) <= 2
        else:  # same three draws, x and z swapped by the host's right-to-left argument evaluation
            ref_wi = np.array(l["wi_bits"], np.uint32).view(np.float32)
            assert abs(wi[1] - ref_wi[1]) < 1e-6 and np.abs(wi - ref_wi).max() < 2.0 * l["scalar"] + 1e-6


def _frame(scene, nrm, mats, cam, w, h, spf, dt, dirty, film=None):
    torch = pytest.importorskip("torch")
    from optix_raytracer_b200 import host
    return scene.playground(cam, host.playground_default_lights(), host.playground_default_materials(), nrm, mats, w, h, spf, dt, dirty, film=film)


def test_playground_frame_properties():
    torch = pytest.importorskip("torch")
    from optix_raytracer_b200 import host
    verts, nrm, mats = orc.playground_scene(6)
    assert verts.shape[0] == 25 * 4 * 36 + 800 and set(np.unique(mats)) == set(range(25)) | {26}
    assert np.allclose(np.linalg.norm(nrm.reshape(-1, 3), axis=1), 1.0, atol=1e-6)
    scene = orc.Scene(verts)
    scene.set_geometry_flags(0)  # OPTIX_GEOMETRY_FLAG_NONE, as TriangleGAS builds it
    cam = host.playground_camera(eye=(0.3, 0.6, -1.2), up=(0.0, 1.0, 0.000073), lookat=(0.0, 0.1, 0.0), fov=50.0)
    w, h = 48, 36
    f1, img1, n1 = _frame(scene, nrm, mats, cam, w, h, 1, 3, True)
    assert np.isfinite(f1).all() and n1 > w * h  # some pixels hit and spawn 5 probes
    # aperture 0: every sample of a launch is the same ray with the same closest-hit seed tea<4>(pixel, dt) -> film(spf=3) = ((p + p) + p)
    f3_, img3, n3 = _frame(scene, nrm, mats, cam, w, h, 3, 3, True)
    assert np.array_equal(f3_, (f1 + f1) + f1) and n3 == 3 * n1
    # progressive accumulation (dirty = false adds to the film); dt re-seeds the closest-hit draws, so the second frame differs
    f_b, _, _ = _frame(scene, nrm, mats, cam, w, h, 1, 4, False, film=f1.copy())
    f_b_alone, _, _ = _frame(scene, nrm, mats, cam, w, h, 1, 4, True)
    assert np.array_equal(f_b, f1 + f_b_alone) and not np.array_equal(f_b_alone, f1)
    assert img3.shape == (h, w, 4) and (img3[..., 3] == 255).all()
    # misses show the reference's direction gradient background, 0.5 * dir + 0.5 in [0, 1]
    assert f1.min() >= 0.0 and f1.max() <= 1.0
    f1, _, _ = _frame(scene, nrm, mats, cam, w, h, 1, 1, True)
    # a DISABLE_ANYHIT geometry is invisible to the light probes (CULL_DISABLED_ANYHIT) but not to the bounce probe: hit pixels get brighter or equal
    scene.set_geometry_flags(1)
    f_cull, _, _ = _frame(scene, nrm, mats, cam, w, h, 1, 1, True)
    assert (f_cull >= f1 - 1e-7).all() and (f_cull > f1).any()
