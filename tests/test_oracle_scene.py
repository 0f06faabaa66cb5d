"""Oracle self-checks that need no GPU: its BVH agrees with brute force (the hit rule is structure
independent), instance transforms, the raycasting quirks, and the path tracer's invariants."""
import ctypes as C

import numpy as np

from oracle import pyoracle as orc
from tests import common


def test_oracle_bvh_equals_brute_force():
    rng = np.random.default_rng(5)
    n = 3000
    c = rng.random((n, 1, 3), dtype=np.float32) * 10
    tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 0.6).astype(np.float32)
    tris[10] = tris[9]
    sc = orc.Scene(tris)
    rays = common.random_rays(rng, 4000, [0, 0, 0], [10, 10, 10])
    a = sc.trace(rays)
    st = sc.trace(rays, stats=True)
    assert st["node_visits"] > 0 and st["tri_tests"] < n * rays.shape[0] / 10
    sc.set_brute(True)
    b = sc.trace(rays)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert (a["t"] >= 0).mean() > 0.3
    occ_a = sc.trace(rays, any_hit=True)["occluded"]
    assert np.array_equal(occ_a, a["t"] >= 0)


def test_duplicate_triangles_resolve_to_lowest_ordinal():
    tri = np.array([[[0, 0, 1], [1, 0, 1], [0, 1, 1]]], np.float32)
    sc = orc.Scene(np.concatenate([tri, tri, tri]))
    rays = np.array([[0.2, 0.2, 0, 0, 0, 0, 1, 1e16]], np.float32)
    r = sc.trace(rays)
    assert r["prim"][0] == 0 and r["t"][0] == 1.0
    # exact tmax is excluded, tmin is excluded
    rays2 = np.array([[0.2, 0.2, 0, 0, 0, 0, 1, 1.0], [0.2, 0.2, 0, 1.0, 0, 0, 1, 10.0]], np.float32)
    assert np.all(sc.trace(rays2)["t"] == -1)


def test_shared_edge_is_watertight():
    """Rays through the shared diagonal of a quad must hit one of the two triangles."""
    quad = np.array([[[0, 0, 5], [1, 0, 5], [1, 1, 5]], [[0, 0, 5], [1, 1, 5], [0, 1, 5]]], np.float32)
    sc = orc.Scene(quad)
    s = np.linspace(0.01, 0.99, 997, dtype=np.float32)
    rays = np.zeros((s.size, 8), np.float32)
    rays[:, 0] = s; rays[:, 1] = s; rays[:, 6] = 1; rays[:, 7] = 1e16
    t = sc.trace(rays)["t"]
    assert np.all(t > 0) and np.all(np.abs(t - 5.0) < 1e-5)


def test_back_face_culling_flags():
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)  # counter-clockwise seen from +z
    sc = orc.Scene(tri)
    front = np.array([[0.2, 0.2, 1, 0, 0, 0, -1, 1e16]], np.float32)
    back = np.array([[0.2, 0.2, -1, 0, 0, 0, 1, 1e16]], np.float32)
    assert sc.trace(front, ray_flags=16)["t"][0] == 1.0 and sc.trace(back, ray_flags=16)["t"][0] == -1.0
    assert sc.trace(front, ray_flags=32)["t"][0] == -1.0 and sc.trace(back, ray_flags=32)["t"][0] == 1.0


def test_instance_transform_shares_t_between_spaces():
    tri = np.array([[[0, 0, 100], [100, 0, 100], [0, 100, 100]]], np.float32)
    m = np.array([0.01, 0, 0, 1, 0, 0.01, 0, 2, 0, 0, 0.01, 3], np.float32)
    sc = orc.Scene(tri, instances=[m])
    rays = np.array([[1.2, 2.2, 0, 0, 0, 0, 1, 1e16]], np.float32)
    r = sc.trace(rays)
    assert r["inst"][0] == 0 and abs(r["t"][0] - 4.0) < 1e-5
    inv = orc.invert34(m)
    assert np.allclose(inv[[0, 5, 10]], 100) and np.allclose(inv[[3, 7, 11]], [-100, -200, -300])


def test_duck_raycast_quirks():
    sc = common.duck_scene()
    tris, nrm = common.deindex(sc["meshes"][0]["primitives"][0])
    scene = orc.Scene(tris, None, instances=[sc["instances"][0]["transform"][:3, :].reshape(12)])
    lo, hi = sc["instances"][0]["world_aabb"]
    w = 96
    h = int(np.float32(w) * (hi - lo)[1] / (hi - lo)[0])
    x0, y0, z, dx, dy = orc.raycast_ortho_scalars(lo, hi, w, h)
    rays = orc.raycast_create_rays(w, h, x0, y0, z, dx, dy)
    assert np.all(rays[:, 3] == 0) and np.all(rays[:, 6] == 1) and np.all(rays[:, 7] == np.float32(1e34))
    hits, ext = scene.raycast_hits(rays, nrm)
    t_ext = ext[:, 0].view(np.float32)
    hit = t_ext >= 0
    assert 0.2 < hit.mean() < 0.8
    # Hit.t = float(unsigned(t)): the reference's truncation quirk (optixRaycasting.cu:76,82)
    assert np.array_equal(hits[hit, 0], np.floor(t_ext[hit]))
    assert np.all(hits[~hit] == np.array([-1, 1, 0, 0], np.float32))
    assert np.allclose(np.linalg.norm(hits[hit, 1:], axis=1), 1, atol=1e-5)
    img = orc.raycast_shade(hits)
    assert np.all(img[~hit] == np.float32(0.2)) and img[hit].min() >= 0 and img[hit].max() <= 1


def _cornell_params(w, h, spl, mode, sub=0):
    from optix_raytracer_b200 import host
    sc = host.load_cornell()
    cam, lt = sc["camera"], sc["light"]
    U, V, W = orc.camera_uvw(cam["eye"], cam["lookat"], cam["up"], cam["fov_y"], w / float(h))
    p = orc.PTParams()
    p.subframe_index, p.width, p.height, p.samples_per_launch, p.nmat, p.mode = sub, w, h, spl, 4, mode
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    p.eye, p.U, p.V, p.W = f3(cam["eye"]), f3(U), f3(V), f3(W)
    p.light_corner, p.light_v1, p.light_v2 = f3(lt["corner"]), f3(lt["v1"]), f3(lt["v2"])
    p.light_normal, p.light_emission, p.bg = f3(host.light_normal(lt["v1"], lt["v2"])), f3(lt["emission"]), f3([0, 0, 0])
    return sc, p


def test_cornell_pathtracer_is_deterministic_and_thread_independent():
    sc, p = _cornell_params(48, 48, 4, 0)
    scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
    a1, f1, s1 = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"], threads=1)
    a2, f2, s2 = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"], threads=7)
    assert s1 == s2 and np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(f1, f2)
    assert np.all(a1[..., 3] == 1) and np.all(f1[..., 3] == 255)
    mean = a1[..., :3].mean(axis=(0, 1))
    assert 0.1 < mean[0] < 0.4 and mean[2] < mean[0]          # warm light, little blue
    # light pixels saturate; segments per path are plausible for Russian roulette with albedo <= 0.8
    assert 2.0 < s1 / (48 * 48 * 4) < 12.0


def test_cornell_multigpu_mode_depth_cap():
    sc, p = _cornell_params(32, 32, 2, 1)
    scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
    a, f, segs = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"])
    # at most 4 radiance + 4 shadow segments per path
    assert segs <= 32 * 32 * 2 * 8


def test_running_mean_over_subframes():
    sc, p = _cornell_params(24, 24, 2, 0)
    scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
    acc0, _, _ = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"])
    p.subframe_index = 1
    cur, _, _ = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"])  # accum=None -> zeros as "prev"
    acc1, _, _ = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"], accum=acc0.copy())
    # lerp(prev, cur, 1/2) where `cur` is what subframe 1 alone produced: cur_alone = 2 * lerp(0, cur, 1/2)
    assert np.allclose(acc1[..., :3], 0.5 * acc0[..., :3] + cur[..., :3], rtol=1e-5, atol=1e-6)


def test_synthetic_mesh_is_closed_and_counts_add_up():
    for total in (2, 100, 5000, 123_457):
        tris, mats = orc.synth_mesh(total)
        assert tris.shape == (total, 3, 3) and mats.shape == (total,)
        assert np.isfinite(tris).all()
        assert tris.min() >= -1e-3 and tris[..., 0].max() <= 556.001 and tris[..., 2].max() <= 559.201
        assert (mats == 3).sum() >= 2 or total < 64


def test_oracle_bvh_equals_brute_force_on_the_edge_case_scene():
    """The oracle's own median-split BVH must not change a single hit record against its brute-force loop (the exact definition) on the
    rays traversals usually disagree on: shared edges and vertices, rays in box face planes, degenerate intervals, geometry a million
    units from the origin (tests/common.py: edge_case_scene; the GPU runs the same set in tests/test_gpu_parity.py)."""
    tris, rays = common.edge_case_scene()
    scene = orc.Scene(tris)
    a = scene.trace(rays)
    occ_a = scene.trace(rays, any_hit=True)["occluded"]
    scene.set_brute(True)
    b = scene.trace(rays)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert np.array_equal(occ_a, scene.trace(rays, any_hit=True)["occluded"])
    assert (a["t"] >= 0).sum() > 2000
    # a vertex shared by six triangles of the coplanar grid goes to the lowest ordinal
    hit = scene.trace(np.array([[1.0, 1.0, 5, 0, 0, 0, -1, 100]], np.float32))
    assert hit["t"][0] == 4.0 and hit["prim"][0] == min(i for i in range(tris.shape[0]) if (tris[i] == np.array([1, 1, 1], np.float32)).all(axis=1).any())
