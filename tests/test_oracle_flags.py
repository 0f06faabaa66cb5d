"""Ray / instance / geometry flag semantics of the oracle's traversal (oracle.cpp: cull_word, trace_geom, trace_scene), checked on
hand-made cases against what the OptiX headers document: face culling and its two instance overrides
(reference include/optix_types.h:1093-1097, 1819-1825), the any-hit state of a triangle under the precedence
ray flags > instance flags > geometry flags (include/optix_types.h:1102-1108, 1800-1806) as seen by CULL_DISABLED_ANYHIT /
CULL_ENFORCED_ANYHIT (:1832-1839), and the visibility mask.  Combinations OptiX declares mutually exclusive (ray DISABLE / ENFORCE_ANYHIT
with the CULL_*_ANYHIT flags, both face-cull flags at once) have the meaning include/b200rt.h gives them.  The GPU path is compared with the oracle on random scenes in
tests/test_gpu_parity.py::test_ray_and_instance_flags_vs_oracle."""
import numpy as np
import pytest

from oracle import pyoracle as orc

# OptixRayFlags / OptixInstanceFlags / OptixGeometryFlags
R_DIS_AH, R_ENF_AH, R_CULL_BACK, R_CULL_FRONT, R_CULL_DIS_AH, R_CULL_ENF_AH = 1, 2, 16, 32, 64, 128
I_NO_CULL, I_FLIP, I_DIS_AH, I_ENF_AH = 1, 2, 4, 8
G_DIS_AH, G_NO_CULL = 1, 4

IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
# counter-clockwise seen from +z: a ray travelling in -z sees the front face
TRI = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
FROM_FRONT = np.array([[0.2, 0.2, 1, 0, 0, 0, -1, 1e16]], np.float32)
FROM_BACK = np.array([[0.2, 0.2, -1, 0, 0, 0, 1, 1e16]], np.float32)


def hit(scene, ray, flags):
    h = scene.trace(ray, ray_flags=flags)["t"][0] >= 0
    assert h == bool(scene.trace(ray, any_hit=True, ray_flags=flags)["occluded"][0])
    return bool(h)


@pytest.mark.parametrize("instanced", [False, True])
def test_face_culling(instanced):
    s = orc.Scene(TRI, None, instances=[IDENT] if instanced else ())
    assert hit(s, FROM_FRONT, 0) and hit(s, FROM_BACK, 0)
    assert hit(s, FROM_FRONT, R_CULL_BACK) and not hit(s, FROM_BACK, R_CULL_BACK)
    assert not hit(s, FROM_FRONT, R_CULL_FRONT) and hit(s, FROM_BACK, R_CULL_FRONT)
    assert not hit(s, FROM_FRONT, R_CULL_BACK | R_CULL_FRONT) and not hit(s, FROM_BACK, R_CULL_BACK | R_CULL_FRONT)
    # OPTIX_GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING exempts the geometry
    s = orc.Scene(TRI, None, instances=[IDENT] if instanced else (), geom_flags=G_DIS_AH | G_NO_CULL)
    assert hit(s, FROM_BACK, R_CULL_BACK) and hit(s, FROM_FRONT, R_CULL_FRONT)


def test_instance_face_flags():
    flip = orc.Scene(TRI, None, instances=[(IDENT, I_FLIP, 1)])
    assert not hit(flip, FROM_FRONT, R_CULL_BACK) and hit(flip, FROM_BACK, R_CULL_BACK)
    assert hit(flip, FROM_FRONT, R_CULL_FRONT) and not hit(flip, FROM_BACK, R_CULL_FRONT)
    nocull = orc.Scene(TRI, None, instances=[(IDENT, I_NO_CULL, 1)])
    assert hit(nocull, FROM_BACK, R_CULL_BACK) and hit(nocull, FROM_FRONT, R_CULL_FRONT)
    # disabling face culling leaves the any-hit culls alone
    assert not hit(nocull, FROM_FRONT, R_CULL_DIS_AH)


def test_anyhit_state_precedence():
    for gflags, base_off in ((G_DIS_AH, True), (0, False)):
        for iflags, inst_off in ((0, None), (I_DIS_AH, True), (I_ENF_AH, False)):
            s = orc.Scene(TRI, None, instances=[(IDENT, iflags, 1)], geom_flags=gflags)
            for rflags, ray_off in ((0, None), (R_DIS_AH, True), (R_ENF_AH, False)):
                off = ray_off if ray_off is not None else inst_off if inst_off is not None else base_off
                assert hit(s, FROM_FRONT, rflags | R_CULL_DIS_AH) == (not off), (gflags, iflags, rflags)
                assert hit(s, FROM_FRONT, rflags | R_CULL_ENF_AH) == off, (gflags, iflags, rflags)
                assert hit(s, FROM_FRONT, rflags)


def test_per_triangle_geometry_flags_and_mask():
    tris = np.concatenate([TRI, TRI + np.array([0, 0, -0.5], np.float32)])  # second triangle half a unit behind the first
    s = orc.Scene(tris, None, geom_flags=np.array([G_DIS_AH, 0], np.uint8))
    assert s.trace(FROM_FRONT)["prim"][0] == 0
    assert s.trace(FROM_FRONT, ray_flags=R_CULL_DIS_AH)["prim"][0] == 1      # the any-hit-disabled one is skipped
    assert s.trace(FROM_FRONT, ray_flags=R_CULL_ENF_AH)["prim"][0] == 0
    assert s.trace(FROM_BACK, ray_flags=R_CULL_ENF_AH)["prim"][0] == 0
    # visibility mask 0: the instance does not exist for the ray; instance index counts it all the same
    s = orc.Scene(TRI, None, instances=[(IDENT, 0, 0), (IDENT, 0, 1)])
    r = s.trace(FROM_FRONT)
    assert r["t"][0] == 1.0 and r["inst"][0] == 1
    assert not hit(orc.Scene(TRI, None, instances=[(IDENT, 0, 0)]), FROM_FRONT, 0)


def test_ray_visibility_mask():
    """optixTrace's 8-bit visibilityMask against OptixInstance::visibilityMask: an instance is traversed when the two share a bit
    (reference include/optix_types.h OptixVisibilityMask; SDK/imgui_test/optixTriangle.cu:130 traces with 255, the other samples with 1).
    include/b200rt.h carries the ray's mask in bits 16-23 of the flags word, XOR 1, so that a word without the field means mask 1."""
    vis = lambda m: ((m ^ 1) & 0xff) << 16
    behind = TRI + np.array([0, 0, -0.5], np.float32)
    s = orc.Scene(np.concatenate([TRI, behind]), None, instances=[(IDENT, 0, 2), (np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, -2], np.float32), 0, 5)])
    # instance 0 (mask 2) holds both triangles at z = 0 / -0.5; instance 1 (mask 5) the same, two units further back
    assert s.trace(FROM_FRONT)["inst"][0] == 1 and s.trace(FROM_FRONT)["t"][0] == 3.0          # default mask 1: only instance 1
    assert s.trace(FROM_FRONT, ray_flags=vis(1))["inst"][0] == 1
    assert s.trace(FROM_FRONT, ray_flags=vis(2))["inst"][0] == 0 and s.trace(FROM_FRONT, ray_flags=vis(2))["t"][0] == 1.0
    assert s.trace(FROM_FRONT, ray_flags=vis(255))["inst"][0] == 0
    assert s.trace(FROM_FRONT, ray_flags=vis(4))["inst"][0] == 1
    assert not hit(s, FROM_FRONT, vis(8)) and not hit(s, FROM_FRONT, vis(0))
    # the field rides along with the other flags
    assert not hit(s, FROM_BACK, vis(2) | R_CULL_BACK) and hit(s, FROM_BACK, vis(2))


def test_flag_word_of_the_library_equals_the_oracles():
    """accel.h: cull_word (what the device triangle tests see) against oracle.cpp: cull_word, over every combination of the eight ray-flag
    bits and the four instance-flag bits that concern triangles — a host function of the C ABI, no GPU involved."""
    from optix_raytracer_b200 import _lib
    lib, o = _lib.load(), orc.lib()
    for rf in range(256):
        for inf in range(16):
            w = lib.b200rt_triangle_flag_word(rf, inf)
            assert w == o.orc_cull_word(rf, inf), (rf, inf)
            assert w & ~0xf3 == 0
    # documented cases: ray flags win over instance flags; FLIP swaps the face-cull bits; DISABLE_TRIANGLE_FACE_CULLING drops them only
    assert lib.b200rt_triangle_flag_word(R_ENF_AH | R_CULL_BACK, I_DIS_AH | I_FLIP) == (2 | R_CULL_FRONT)
    assert lib.b200rt_triangle_flag_word(R_CULL_BACK | R_CULL_DIS_AH, I_NO_CULL | I_ENF_AH) == (2 | R_CULL_DIS_AH)
