"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical
inputs.  Bit-exact for hit records (t, primitive, instance, barycentrics) and accumulated radiance; the
sRGB u8 frame is allowed 1 LSB (device __powf vs libm powf, documented in DESIGN.md)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from tests import common  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from optix_raytracer_b200 import host
    c = host.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


def _assert_hits_equal(got, ref, what):
    for k in ("prim", "inst", "t", "b1", "b2"):
        a = got[k].view(np.uint32) if got[k].dtype == np.float32 else got[k]
        b = ref[k].view(np.uint32) if ref[k].dtype == np.float32 else ref[k]
        hit = ref["t"] >= 0
        if k in ("b1", "b2", "prim", "inst"):
            a, b = a[hit], b[hit]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"{what}: {bad.size} rays differ in {k}; first {bad[:5]}: got {got[k][hit][bad[:5]] if k!='t' else got[k][bad[:5]]} ref {ref[k][hit][bad[:5]] if k!='t' else ref[k][bad[:5]]}"


@pytest.mark.parametrize("node_format", ["q8", "f32"])
def test_cornell_random_rays_closest_and_any(ctx, orc, node_format, monkeypatch):
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_NODE_FORMAT", node_format)
    pt = host.PathTracer(ctx, 32, 32, 1)
    sc = pt.scene
    scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
    rng = np.random.default_rng(1)
    rays = common.random_rays(rng, 200_000, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
    d_rays = ctx.to_device(rays)
    got = host.ext_hits_to_numpy(ctx.trace_closest(pt.accel, d_rays))
    ref = scene.trace(rays)
    assert (ref["t"] >= 0).mean() > 0.5
    _assert_hits_equal(got, ref, "cornell closest")
    # occlusion rays with finite tmax
    rays2 = rays.copy()
    rays2[:, 7] = rng.random(rays.shape[0], dtype=np.float32) * 800
    occ = ctx.trace_any(pt.accel, ctx.to_device(rays2)).cpu().numpy().astype(bool)
    ref_occ = scene.trace(rays2, any_hit=True)["occluded"]
    assert np.array_equal(occ, ref_occ)
    assert 0.1 < occ.mean() < 0.9


@pytest.mark.parametrize("mode,groups,ray_sort", [(0, 1, 0), (1, 1, 0), (0, 3, 0), (1, 4, 0), (0, 8, 0), (0, 4, 1), (1, 2, 1)])
def test_cornell_pathtracer_bit_exact(ctx, orc, mode, groups, ray_sort):
    """groups = b200rt_pt_options.sample_groups: 1 is the reference's flat summation order; 3 (uneven split of 4 samples), 4 and 8
    (more groups than samples: empty groups) run the samples of a pixel in parallel lanes and must match the oracle's grouped sum.
    ray_sort = b200rt_pt_options.ray_sort: the rays of an iteration are radix-sorted by origin cell and direction before they are
    traced (image large enough for iterations above the 65536-ray threshold); a lane owns its state, so nothing may change."""
    from optix_raytracer_b200 import host
    w, h, spl = (320, 240, 4) if ray_sort else (96, 80, 4)
    pt = host.PathTracer(ctx, w, h, spl, multigpu=(0, 1) if mode else None)
    pt.sample_groups = groups
    pt.ray_sort = ray_sort
    sc = pt.scene
    scene = orc.Scene(sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"])
    ref_accum = None
    for sub in range(2):  # second subframe exercises the running-mean lerp
        st = pt.launch_subframe(sub, collect_stats=True)
        torch.cuda.synchronize()
        p = common.oracle_pt_params(orc, pt.params, mode)
        p.groups = groups
        ref_accum, ref_frame, segs = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"], accum=ref_accum)
        assert st.radiance_segments + st.shadow_segments == segs
        accum = pt.accum.cpu().numpy()
        if mode:  # per-sample layout -> pixel order
            img = np.zeros((h, w, 4), np.float32)
            idx = pt.sample_index.cpu().numpy()
            ok = (idx[:, 0] < w) & (idx[:, 1] < h)
            img[idx[ok, 1], idx[ok, 0]] = accum[ok]
            accum = img
        assert np.array_equal(accum.view(np.uint32), ref_accum.view(np.uint32)), f"subframe {sub}: accum differs"
        frame = pt.frame.cpu().numpy().astype(np.int32)
        if mode:
            # optixMultiGPU adds deviceColor(device_idx) to the displayed colour only; device_idx=3 adds nothing
            pass
        assert np.abs(frame - ref_frame.astype(np.int32)).max() <= 1


def test_duck_raycast_bit_exact(ctx, orc):
    from optix_raytracer_b200 import host
    sc = common.duck_scene()
    rc = host.Raycaster(ctx, sc)
    n = rc.buffer_rays(320)
    rc.launch()
    torch.cuda.synchronize()
    prim = sc["meshes"][0]["primitives"][0]
    tris, nrm = common.deindex(prim)
    scene = orc.Scene(tris, None, instances=[sc["instances"][0]["transform"][:3, :].reshape(12)])
    # ray generation restated by the oracle must match the device kernel bit for bit
    x0, y0, z, dx, dy = orc.raycast_ortho_scalars(rc.bbmin, rc.bbmax, rc.width, rc.height, 0.05)
    ref_rays = orc.raycast_create_rays(rc.width, rc.height, x0, y0, z, dx, dy)
    assert np.array_equal(rc.rays.cpu().numpy().view(np.uint32), ref_rays.view(np.uint32))
    ref_rays_t = orc.raycast_translate(ref_rays, rc.translate_offset)
    assert np.array_equal(rc.rays_translated.cpu().numpy().view(np.uint32), ref_rays_t.view(np.uint32))
    for rays, hits, ext, name in ((ref_rays, rc.hits, rc.ext, "original"), (ref_rays_t, rc.hits_translated, rc.ext_translated, "translated")):
        ref_hits, ref_ext = scene.raycast_hits(rays, nrm)
        got_ext = host.ext_hits_to_numpy(ext)
        ref = {"t": ref_ext[:, 0].view(np.float32), "prim": ref_ext[:, 1], "inst": ref_ext[:, 2], "b1": ref_ext[:, 3].view(np.float32),
               "b2": ref_ext[:, 4].view(np.float32)}
        assert (ref["t"] >= 0).mean() > 0.2
        _assert_hits_equal(got_ext, ref, f"duck {name}")
        assert np.array_equal(hits.cpu().numpy().view(np.uint32), ref_hits.view(np.uint32)), f"duck {name}: Hit buffer differs"
        img = rc.shade(hits).cpu().numpy()
        assert np.array_equal(img.view(np.uint32), orc.raycast_shade(ref_hits).view(np.uint32))


@pytest.mark.parametrize("node_format,hierarchy", [("q8", "lbvh"), ("f32", "lbvh"), ("q8", "ploc"), ("f32", "ploc")])
def test_triangle_soup_bvh_vs_oracle(ctx, orc, node_format, hierarchy, monkeypatch):
    """A random soup big enough for a multi-level wide BVH, including degenerate and duplicate triangles; both node encodings and both
    binary hierarchies the collapse can start from (Karras tree over Morton codes, PLOC clustering: bvh_build.cu)."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_NODE_FORMAT", node_format)
    monkeypatch.setenv("B200RT_HIERARCHY", hierarchy)
    rng = np.random.default_rng(7)
    n = 30_000
    c = rng.random((n, 1, 3), dtype=np.float32) * 10
    tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 0.4).astype(np.float32)
    tris[100] = tris[99]            # exact duplicate: tie broken by ordinal
    tris[200, 2] = tris[200, 1]     # zero-area triangle
    verts = ctx.to_device(tris.reshape(-1, 3))
    accel = ctx.build_accel([ctx.triangle_input(verts, vertex_stride=12)])
    info = accel.info()
    assert info.num_triangles == n and info.num_nodes > 100
    scene = orc.Scene(tris)
    rays = common.random_rays(rng, 100_000, [0, 0, 0], [10, 10, 10])
    got = host.ext_hits_to_numpy(ctx.trace_closest(accel, ctx.to_device(rays)))
    ref = scene.trace(rays)
    _assert_hits_equal(got, ref, "soup")


@pytest.mark.parametrize("hierarchy", ["lbvh", "ploc"])
def test_edge_case_rays_and_geometry_vs_oracle(ctx, orc, hierarchy, monkeypatch):
    """Where traversals usually disagree: rays aimed exactly at shared edges and vertices of a coplanar grid (watertightness: the hit goes
    to the lower ordinal, never through the crack), axis-parallel rays lying in the plane of box faces, zero direction components,
    tmin == tmax, tmax < tmin, tmax = +inf, origins on the surface, geometry a million units from the origin, slivers and a triangle
    fan sharing one vertex.  Hit records and occlusion flags must equal the oracle's bit for bit."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_HIERARCHY", hierarchy)
    tris, rays = common.edge_case_scene()
    accel = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
    scene = orc.Scene(tris)
    got = host.ext_hits_to_numpy(ctx.trace_closest(accel, ctx.to_device(rays)))
    ref = scene.trace(rays)
    assert (ref["t"] >= 0).sum() > 2000
    _assert_hits_equal(got, ref, "edge cases")
    occ = ctx.trace_any(accel, ctx.to_device(rays)).cpu().numpy().astype(bool)
    assert np.array_equal(occ, scene.trace(rays, any_hit=True)["occluded"])


@pytest.mark.parametrize("node_format", ["q8", "f32"])
def test_scene_far_from_origin_vs_oracle(ctx, orc, node_format, monkeypatch):
    """A scene that lives only a million units from the origin (coordinate spacing 0.0625, extent 4): the padding of the scene bounds
    and of every child box is below that spacing and has to be applied with outward rounding, or rays lying exactly in a face plane
    of a box (grid vertices and edges, zero direction components) are culled.  Both node encodings, straight and IAS-instanced."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_NODE_FORMAT", node_format)
    tris, rays = common.edge_case_scene()
    tris = tris[1216:]
    far = rays[np.abs(rays[:, 0]) > 1e5]
    assert tris.shape[0] == 32 and far.shape[0] == 50
    rng = np.random.default_rng(3)
    rnd = common.random_rays(rng, 5000, [1e6, 1e6, 1e6 - 1], [1e6 + 4, 1e6 + 4, 1e6 + 1])
    rays = np.concatenate([far, rnd]).astype(np.float32)
    gas = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    ias = ctx.build_accel([ctx.instance_input([(ident, 0, gas)])], compact=False)
    for accel, scene, what in ((gas, orc.Scene(tris), "gas"), (ias, orc.Scene(tris, None, instances=[ident]), "ias")):
        ref = scene.trace(rays)
        assert (ref["t"][:50] >= 0).sum() >= 40 and (ref["t"] >= 0).sum() > 500
        _assert_hits_equal(host.ext_hits_to_numpy(ctx.trace_closest(accel, ctx.to_device(rays))), ref, f"far scene {what}")
        occ = ctx.trace_any(accel, ctx.to_device(rays)).cpu().numpy().astype(bool)
        assert np.array_equal(occ, scene.trace(rays, any_hit=True)["occluded"])


def test_multi_instance_ias_vs_oracle(ctx, orc):
    """One GAS instanced five times (rotated, scaled, overlapping) plus an instance whose visibility mask is 0: hit records (t, primitive,
    instance, barycentrics) and occlusion flags are those of the oracle — closest over instances = min (t, instance index, ordinal); the
    bounds test that drops a ray at an instance it passes by (trav_begin<BOUNDS>) must not change anything."""
    from optix_raytracer_b200 import host
    rng = np.random.default_rng(17)
    n = 3000
    c = rng.random((n, 1, 3), dtype=np.float32) * 2 - 1
    tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 0.15).astype(np.float32)
    gas = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])

    def xf(angle, scale, t):
        ca, sa = np.cos(angle), np.sin(angle)
        m = np.array([[ca * scale, 0, sa * scale, t[0]], [0, scale * 0.8, 0, t[1]], [-sa * scale, 0, ca * scale, t[2]]], np.float32)
        return m.reshape(12)
    xfs = [xf(0.0, 1.0, (0, 0, 0)), xf(0.7, 0.5, (1.5, 0.2, 0)), xf(-1.1, 1.3, (-1, -0.5, 2)), xf(2.0, 0.9, (0.5, 0.5, 0.5)), xf(0.3, 2.0, (0, 3, 0))]
    inst = [(m, 0, gas) for m in xfs] + [(xf(0.1, 5.0, (0, 0, 0)), 0, gas, 0)]   # the last one is invisible (mask 0)
    ias = ctx.build_accel([ctx.instance_input(inst)], compact=False)
    assert ias.info().num_instances == 6
    scene = orc.Scene(tris, None, instances=xfs)
    rays = common.random_rays(rng, 200_000, [-4, -3, -4], [4, 6, 5])
    got = host.ext_hits_to_numpy(ctx.trace_closest(ias, ctx.to_device(rays)))
    ref = scene.trace(rays)
    assert (ref["t"] >= 0).mean() > 0.1 and len(np.unique(ref["inst"][ref["t"] >= 0])) == 5
    _assert_hits_equal(got, ref, "multi-instance")
    rays[:, 7] = rng.random(rays.shape[0], dtype=np.float32) * 6
    occ = ctx.trace_any(ias, ctx.to_device(rays)).cpu().numpy().astype(bool)
    assert np.array_equal(occ, scene.trace(rays, any_hit=True)["occluded"])


def test_ray_and_instance_flags_vs_oracle(ctx, orc):
    """Every ray-flag family against instances that carry every instance-flag family, over geometry whose two SBT records differ in
    their geometry flags: face culling (with OPTIX_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING / FLIP_TRIANGLE_FACING and
    OPTIX_GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING), the any-hit state under ray > instance > geometry precedence as seen by
    CULL_DISABLED_ANYHIT / CULL_ENFORCED_ANYHIT, an invisible instance.  Hit records and occlusion flags equal the oracle's
    (semantics pinned on hand-made cases in tests/test_oracle_flags.py)."""
    from optix_raytracer_b200 import host
    rng = np.random.default_rng(23)
    n = 2500
    c = rng.random((n, 1, 3), dtype=np.float32) * 2 - 1
    tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 0.25).astype(np.float32)
    sbt = (rng.random(n) < 0.5).astype(np.uint32)
    rec_flags = [1, 4]  # record 0: DISABLE_ANYHIT; record 1: any-hit enabled + DISABLE_TRIANGLE_FACE_CULLING
    gas = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12, sbt_index=ctx.to_device(sbt), num_sbt=2,
                                              flags=rec_flags)])
    tri_flags = np.array(rec_flags, np.uint8)[sbt]

    def xf(angle, scale, t):
        ca, sa = np.cos(angle), np.sin(angle)
        return np.array([[ca * scale, 0, sa * scale, t[0]], [0, scale, 0, t[1]], [-sa * scale, 0, ca * scale, t[2]]], np.float32).reshape(12)
    xfs = [xf(0.0, 1.0, (0, 0, 0)), xf(0.9, 0.8, (1.0, 0.3, 0)), xf(-0.6, 1.2, (-1, -0.4, 1)), xf(1.7, 0.9, (0.4, 0.6, 0.2)), xf(0.2, 1.5, (0, 1.5, 0)),
           xf(2.5, 1.1, (0.5, -1, -0.5)), xf(0.1, 4.0, (0, 0, 0))]
    iflags = [0, 1, 2, 4, 8, 2 | 8, 0]
    masks = [1, 1, 1, 1, 1, 255, 0]
    ias = ctx.build_accel([ctx.instance_input([(m, 0, gas, k, f) for m, f, k in zip(xfs, iflags, masks)])], compact=False)
    scene_gas = orc.Scene(tris, sbt, geom_flags=tri_flags)
    scene_ias = orc.Scene(tris, sbt, instances=[(m, f, k) for m, f, k in zip(xfs, iflags, masks)], geom_flags=tri_flags)
    rays = common.random_rays(rng, 30_000, [-3, -3, -3], [3, 4, 3])
    rays_any = rays.copy()
    rays_any[:, 7] = rng.random(rays.shape[0], dtype=np.float32) * 5
    d_rays, d_rays_any = ctx.to_device(rays), ctx.to_device(rays_any)
    seen = set()
    for rf in (0, 16, 32, 48, 64, 128, 1 | 64, 2 | 64, 1 | 128, 2 | 128, 16 | 64, 32 | 128 | 2, 1, 2):
        for accel, scene, what in ((gas, scene_gas, "gas"), (ias, scene_ias, "ias")):
            got = host.ext_hits_to_numpy(ctx.trace_closest(accel, d_rays, ray_flags=rf))
            ref = scene.trace(rays, ray_flags=rf)
            _assert_hits_equal(got, ref, f"{what} ray_flags {rf:#x}")
            occ = ctx.trace_any(accel, d_rays_any, ray_flags=rf | 4).cpu().numpy().astype(bool)
            assert np.array_equal(occ, scene.trace(rays_any, any_hit=True, ray_flags=rf)["occluded"]), f"{what} occlusion, ray_flags {rf:#x}"
            seen.add((what, rf, int((ref["t"] >= 0).sum())))
    # the ray's 8-bit visibility mask (bits 16-23 of the flags word, XOR 1): instance k is traversed when (its mask & the ray's) != 0
    masks2 = [1, 2, 4, 3, 6, 255, 0]
    ias2 = ctx.build_accel([ctx.instance_input([(m, 0, gas, k, f) for m, f, k in zip(xfs, iflags, masks2)])], compact=False)
    scene_ias2 = orc.Scene(tris, sbt, instances=[(m, f, k) for m, f, k in zip(xfs, iflags, masks2)], geom_flags=tri_flags)
    for vis in (1, 2, 4, 6, 8, 255, 0):
        rf = ((vis ^ 1) & 0xff) << 16
        ref = scene_ias2.trace(rays, ray_flags=rf)
        _assert_hits_equal(host.ext_hits_to_numpy(ctx.trace_closest(ias2, d_rays, ray_flags=rf)), ref, f"ray visibility mask {vis}")
        occ = ctx.trace_any(ias2, d_rays_any, ray_flags=rf | 4).cpu().numpy().astype(bool)
        assert np.array_equal(occ, scene_ias2.trace(rays_any, any_hit=True, ray_flags=rf)["occluded"]), f"occlusion, ray visibility mask {vis}"
        visible = {k for k, m in enumerate(masks2) if m & vis}
        assert set(np.unique(ref["inst"][ref["t"] >= 0]).tolist()) == visible, f"ray visibility mask {vis}: instances hit"
    # the flags do something: the hit counts differ between the settings, and the invisible instance is never reported
    assert len({c for w, rf, c in seen if w == "ias"}) >= 6
    assert not np.any(scene_ias.trace(rays)["inst"][scene_ias.trace(rays)["t"] >= 0] == 6)


@pytest.mark.parametrize("node_format", ["q8", "f32"])
def test_synthetic_mesh_generator_matches_oracle_and_traces(ctx, orc, node_format, monkeypatch):
    """The procedural scene of BASELINE.json configs[4]: device generator == oracle restatement bit for bit,
    the built BVH satisfies the structural invariants, and path tracing it (optixMultiGPU programs) is bit-exact."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_NODE_FORMAT", node_format)
    T = 200_000
    verts, mats = host.synthetic_mesh(ctx, T, 0)
    torch.cuda.synchronize()
    tris, omats = orc.synth_mesh(T, 0)
    got = verts.cpu().numpy().reshape(T, 3, 4)
    assert np.array_equal(got[:, :, :3].view(np.uint32), tris.view(np.uint32))
    assert np.array_equal(mats.cpu().numpy().astype(np.uint32), omats)
    w, h, spl = 64, 48, 2
    pt = host.PathTracer(ctx, w, h, spl, vertices=verts, mat_indices=mats, multigpu=(0, 1), compact=True)
    gas = common.decode_gas(pt.accel.buf.cpu().numpy())
    assert gas["num_tris"] == T
    depth, leaves = common.validate_gas(gas)
    assert depth <= 32
    st = pt.launch_subframe(0, collect_stats=1)
    torch.cuda.synchronize()
    scene = orc.Scene(tris, omats)
    p = common.oracle_pt_params(orc, pt.params, 1)
    sc = pt.scene
    ref_accum, ref_frame, segs = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"])
    assert st.radiance_segments + st.shadow_segments == segs
    img = np.zeros((h, w, 4), np.float32)
    idx = pt.sample_index.cpu().numpy()
    ok = (idx[:, 0] < w) & (idx[:, 1] < h)
    img[idx[ok, 1], idx[ok, 0]] = pt.accum.cpu().numpy()[ok]
    assert np.array_equal(img.view(np.uint32), ref_accum.view(np.uint32))


@pytest.mark.parametrize("node_format,hierarchy", [("q8", "lbvh"), ("f32", "lbvh"), ("q8", "ploc"), ("f32", "ploc")])
def test_bvh_invariants_and_compaction(ctx, node_format, hierarchy, monkeypatch):
    """Both node encodings (accel.h): Node8 with 8-bit boxes and Node8F with fp32 boxes (forced through B200RT_NODE_FORMAT), from both
    binary hierarchies (B200RT_HIERARCHY); builds are deterministic (compacted == uncompacted node for node)."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_NODE_FORMAT", node_format)
    monkeypatch.setenv("B200RT_HIERARCHY", hierarchy)
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 5, 33, 1000):
        tris = (rng.random((n, 3, 3), dtype=np.float32) * 4).astype(np.float32)
        verts = ctx.to_device(tris.reshape(-1, 3))
        a = ctx.build_accel([ctx.triangle_input(verts, vertex_stride=12)], compact=False)
        b = ctx.build_accel([ctx.triangle_input(verts, vertex_stride=12)], compact=True)
        torch.cuda.synchronize()
        ga, gb = common.decode_gas(a.buf.cpu().numpy()), common.decode_gas(b.buf.cpu().numpy())
        common.validate_gas(ga)
        common.validate_gas(gb)
        assert np.array_equal(ga["nodes"], gb["nodes"]) and np.array_equal(ga["tris"].view(np.uint32), gb["tris"].view(np.uint32))
        assert b.buf.numel() <= a.buf.numel()
        assert gb["total_bytes"] == 128 + gb["node_bytes"] * gb["num_nodes"] + 48 * n


@pytest.mark.parametrize("aperture,ortho", [(0.0, False), (0.05, False), (0.0, True)])
def test_playground_bit_exact(ctx, orc, aperture, ortho):
    """imgui_test (BASELINE.json configs[3]): scene generator == oracle restatement, film bit-exact over two frames (dirty, then
    accumulating), image within 1 LSB."""
    from optix_raytracer_b200 import host
    w, h, spf, rows = 64, 48, 3, 10
    cam = host.playground_camera(eye=(0.3, 0.6, -1.2), up=(0.0, 1.0, 0.000073), lookat=(0.0, 0.1, 0.0), fov=50.0, aperture=aperture, ortho=ortho)
    pg = host.Playground(ctx, w, h, spf=spf, rows=rows, camera=cam)
    verts, nrm, mats = orc.playground_scene(rows)
    torch.cuda.synchronize()
    assert np.array_equal(pg.vertices.cpu().numpy().reshape(-1, 3, 3).view(np.uint32), verts.view(np.uint32))
    assert np.array_equal(pg.normals.cpu().numpy().reshape(-1, 3, 3).view(np.uint32), nrm.view(np.uint32))
    assert np.array_equal(pg.mat_indices.cpu().numpy(), mats)
    scene = orc.Scene(verts)
    scene.set_geometry_flags(0)
    film = None
    for frame, dirty in enumerate((True, False)):
        pg.launch_frame(dirty=dirty)
        torch.cuda.synchronize()
        dt = pg.params.dt
        assert dt == spf * (frame + 1)
        film, image, nrays = scene.playground(cam, pg.lights_bytes, pg.materials, nrm, mats, w, h, spf, dt, dirty, film=film)
        got = pg.film.cpu().numpy()
        assert (got > 0).mean() > 0.9
        assert np.array_equal(got.view(np.uint32), film.view(np.uint32)), f"frame {frame}: film differs"
        assert np.abs(pg.image.cpu().numpy().astype(np.int32) - image.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("alpha_mode", [0, 1])
def test_raycast_launch_shapes_agree(ctx, alpha_mode):
    """optixRaycasting's launch takes optixLaunch's width x height; the one-ray-per-thread kernel gives a warp an 8 x 4 tile of launch
    indices (raycast.cu).  Sizes that are no multiple of the tile, and the flat n x 1 launch of the same buffer (rows of 32 indices), must
    fill the same Hit / ext records — with and without an any-hit program (alpha MASK)."""
    from optix_raytracer_b200 import host
    sc = common.duck_scene() if alpha_mode == 0 else common.duck_alpha_scene(1)
    rc = host.Raycaster(ctx, sc)
    for width in (37, 101):
        n = rc.buffer_rays(width)
        assert rc.height % 4 != 0 or width % 8 != 0
        rc.launch()
        torch.cuda.synchronize()
        tiled_hits, tiled_ext = rc.hits.clone(), rc.ext.clone()
        assert (tiled_hits[:, 0] >= 0).float().mean() > 0.2
        rc.hits.fill_(-7.0); rc.ext.fill_(-7)
        ctx.launch_raycast(rc.programs, rc.d_params.data_ptr(), rc.sbt, n, 1, rc.ext.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(tiled_hits.view(torch.int32), rc.hits.view(torch.int32)), f"width {width}: Hit buffers of the two launch shapes differ"
        assert torch.equal(tiled_ext, rc.ext), f"width {width}: ext records of the two launch shapes differ"
    rc.close()


@pytest.mark.parametrize("alpha_mode,double_sided,size", [(0, False, (160, 120)), (0, False, (157, 83)), (2, False, (160, 120)), (2, True, (160, 120))])
def test_whitted_untextured_bit_exact(ctx, orc, alpha_mode, double_sided, size):
    """optixMeshViewer (BASELINE.json configs[2]) on the Duck with its texture removed: accum bit-exact against the oracle over
    two subframes (pixel-centre ray, then jittered + running mean), frame within 1 LSB.  alpha_mode 2 = ALPHA_MODE_BLEND with
    base-colour alpha 0.6: every hit continues behind itself (whitted.cu:266-286); doubleSided lifts the back-face culling of the
    radiance rays, so the chain gets a second level inside the duck.  Textured shading is compared on the GPU against the
    reference programs on OptiX (tests/test_gpu_optix_parity.py)."""
    from optix_raytracer_b200 import host
    sc = common.duck_scene(textured=False)
    if alpha_mode:
        sc["materials"][0].update({"alpha_mode": alpha_mode, "double_sided": double_sided, "base_color": [1.0, 0.9, 0.8, 0.6]})
    w, h = size   # 157 x 83: no multiple of RAYGEN's 8 x 4 pixel tiles (whitted.cu)
    mv = host.MeshViewer(ctx, sc, w, h)
    prim = sc["meshes"][0]["primitives"][0]
    tris, nrm = common.deindex(prim)
    scene = orc.Scene(tris, None, instances=[sc["instances"][0]["transform"][:3, :].reshape(12)],
                      geom_flags=4 if double_sided else 0)
    p = orc.WhittedParams()
    p.width, p.height = w, h
    for k in ("eye", "U", "V", "W", "miss_color"):
        setattr(p, k, getattr(mv.params, k))
    m = sc["materials"][0]
    p.base_color = (C.c_float * 4)(*m["base_color"])
    p.metallic, p.roughness = m["metallic"], m["roughness"]
    p.emissive = (C.c_float * 3)(*m["emissive_factor"])
    p.alpha_mode = alpha_mode
    lights = mv.d_lights.cpu().numpy().tobytes()
    accum = None
    opaque_rays = None
    for sub in range(2):
        mv.launch_subframe(sub)
        torch.cuda.synchronize()
        p.subframe_index = sub
        accum, frame, nrays = scene.whitted(p, lights, normals=nrm, accum=accum)
        got = mv.accum.cpu().numpy()
        assert nrays > w * h and (got[..., :3] != np.float32(0.1)).any(axis=-1).mean() > 0.03  # the duck covers part of the image
        assert np.array_equal(got.view(np.uint32), accum.view(np.uint32)), f"subframe {sub}: accum differs"
        assert np.abs(mv.frame.cpu().numpy().astype(np.int32) - frame.astype(np.int32)).max() <= 1
    mv.close()


def test_unmodified_optix_host_calls_run_on_the_function_table_shim():
    """SURVEY 8(b) layer 2: a library with the soname libnvoptix.so.1 answers optixQueryFunctionTable, so the reference's own call sequence
    (here: oracle/optix_ref/optix_harness.cpp, built against the reference's optix_stubs.h) runs on b200rt and gives bit-identical
    results to the native C ABI — Cornell (both program sets), optixRaycasting with an alpha mask, whitted with BLEND, imgui_test."""
    import os
    import pathlib
    import subprocess
    import sys
    root = pathlib.Path(__file__).resolve().parents[1]
    if not (root / "oracle" / "_ref" / "liboptixref.so").exists():
        pytest.skip("oracle/_ref/liboptixref.so (the OptiX host harness) is built only where /root/reference exists")
    env = dict(os.environ, B200RT_OPTIX_SHIM="1")
    r = subprocess.run([sys.executable, str(root / "tests" / "shim_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]



@pytest.mark.parametrize("hierarchy", ["lbvh", "ploc"])
def test_builder_takes_any_input_and_keeps_the_tree_within_the_traversal_stack(ctx, orc, hierarchy, monkeypatch):
    """Inputs that make a tall binary hierarchy: positions in geometric progression (every Morton bit splits off one triangle: a chain),
    a pile of coincident triangles (identical keys: the index bits decide), and both together.  OptiX builds anything, so does this
    builder: the collapse opens the tallest subtrees first where the surface-area order would run out of levels (bvh_build.cu), the
    tree stays within TRAV_STACK / 2 levels and the hits equal the oracle's.  Run once more with the depth target lowered, which forces
    the guard to act on an ordinary soup."""
    from optix_raytracer_b200 import host
    monkeypatch.setenv("B200RT_HIERARCHY", hierarchy)
    rng = np.random.default_rng(5)
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32) * 0.01
    chain = np.stack([base * np.float32(2.0 ** -k) + np.float32(100.0 * 2.0 ** -k) for k in range(60)]).astype(np.float32)
    pile = np.repeat((base + 3.0)[None], 5000, axis=0).astype(np.float32)
    soup = (rng.random((20_000, 1, 3), dtype=np.float32) * 100 + (rng.random((20_000, 3, 3), dtype=np.float32) - 0.5)).astype(np.float32)
    tris = np.concatenate([chain, pile, soup]).astype(np.float32)
    rays = np.concatenate([common.random_rays(rng, 60_000, [0, 0, 0], [100, 100, 100]),
                           common.random_rays(rng, 20_000, [2.9, 2.9, 2.9], [3.1, 3.1, 3.1]),
                           common.random_rays(rng, 20_000, [0, 0, 0], [0.5, 0.5, 0.5])]).astype(np.float32)
    scene = orc.Scene(tris)
    ref = scene.trace(rays)
    depths = []
    for target in (None, "8"):
        if target:
            monkeypatch.setenv("B200RT_MAX_WIDE_DEPTH", target)
        accel = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
        gas = common.decode_gas(accel.buf.cpu().numpy())
        depth, _ = common.validate_gas(gas)
        assert depth == gas["depth"] and depth <= 32
        depths.append(depth)
        got = host.ext_hits_to_numpy(ctx.trace_closest(accel, ctx.to_device(rays)))
        _assert_hits_equal(got, ref, f"tall hierarchy ({hierarchy}, depth {depth})")
    assert depths[1] <= depths[0]


def test_a_million_coincident_triangles_build_and_trace(ctx, orc):
    from optix_raytracer_b200 import host
    tri = np.array([[[0, 0, 1], [2, 0, 1], [0, 2, 1]]], np.float32)
    tris = np.repeat(tri, 1_000_000, axis=0)
    accel = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
    info = accel.info()
    assert info.num_triangles == 1_000_000 and info.depth <= 32
    rays = np.array([[0.5, 0.5, 5, 0, 0, 0, -1, 100], [3, 3, 5, 0, 0, 0, -1, 100], [0.25, 0.25, 0, 0, 0, 0, 1, 100]], np.float32)
    got = host.ext_hits_to_numpy(ctx.trace_closest(accel, ctx.to_device(rays)))
    # the closest hit among equal t is the lowest ordinal: primitive 0
    assert got["t"][0] == 4.0 and got["prim"][0] == 0 and got["t"][1] < 0 and got["t"][2] == 1.0 and got["prim"][2] == 0


def test_raycast_helper_kernels_equal_the_reference_kernels(ctx):
    """b200rt_create_rays_ortho / b200rt_translate_rays / b200rt_shade_hits against the reference's OWN plain-CUDA
    kernels (SDK/optixRaycasting/optixRaycastingKernels.cu:42-115), compiled where they lie with the reference's nvcc flags into
    oracle/_ref/librefraycast.so (oracle/ref_raycast_kernels.cu): ray buffers and the shaded image are identical bit for bit, for the
    Duck's bounding box at the sample's default width and for an awkward box / size."""
    import ctypes as C
    import pathlib
    from optix_raytracer_b200 import host
    so = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "librefraycast.so"
    if not so.exists():
        pytest.skip("oracle/_ref/librefraycast.so not built (make -C oracle needs /root/reference)")
    ref = C.CDLL(str(so))
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    dev = ctx.torch_device
    rng = np.random.default_rng(41)
    sc = common.duck_scene()
    cases = [(1040, np.asarray(sc["bbox_min"], np.float32), np.asarray(sc["bbox_max"], np.float32))] if "bbox_min" in sc else []
    cases += [(1040, np.array([-86.3, 9.9, -61.4], np.float32), np.array([96.2, 164.0, 53.9], np.float32)),
              (333, np.array([1e3 + 0.1, -7.77, 0.0], np.float32), np.array([1e3 + 3.3, 1.23, 1e-3], np.float32))]
    for width, bbmin, bbmax in cases:
        span = bbmax - bbmin
        height = int(np.float32(width) * span[1] / span[0])
        n = width * height
        mine = torch.zeros((n, 8), dtype=torch.float32, device=dev)
        theirs = torch.zeros((n, 8), dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.b200rt_create_rays_ortho(ctx.h, ctx.stream, mine.data_ptr(), width, height, f3(bbmin), f3(bbmax), 0.05), "create_rays_ortho")
        torch.cuda.synchronize()
        assert ref.ref_create_rays_ortho(C.c_void_p(theirs.data_ptr()), width, height, f3(bbmin), f3(bbmax), C.c_float(0.05)) == 0
        assert torch.equal(mine.view(torch.int32), theirs.view(torch.int32)), f"createRaysOrtho differs ({width}x{height})"
        off = (span * np.array([0.2, 0, 0], np.float32)).astype(np.float32)
        ctx.check(ctx.lib.b200rt_translate_rays(ctx.h, ctx.stream, mine.data_ptr(), n, f3(off)), "translate_rays")
        torch.cuda.synchronize()
        assert ref.ref_translate_rays(C.c_void_p(theirs.data_ptr()), n, f3(off)) == 0
        assert torch.equal(mine.view(torch.int32), theirs.view(torch.int32)), "translateRays differs"
        # hits: a mix of misses (t = -1) and unit normals
        hits = np.zeros((n, 4), np.float32)
        hits[:, 0] = np.where(rng.random(n) < 0.4, -1.0, rng.random(n) * 100).astype(np.float32)
        nrm = rng.normal(size=(n, 3)).astype(np.float32)
        hits[:, 1:] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        d_hits = ctx.to_device(hits)
        img_mine = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        img_theirs = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.b200rt_shade_hits(ctx.h, ctx.stream, img_mine.data_ptr(), n, d_hits.data_ptr()), "shade_hits")
        torch.cuda.synchronize()
        assert ref.ref_shade_hits(C.c_void_p(img_theirs.data_ptr()), n, C.c_void_p(d_hits.data_ptr())) == 0
        assert torch.equal(img_mine.view(torch.int32), img_theirs.view(torch.int32)), "shadeHits differs"


def test_empty_and_tiny_inputs(ctx, orc):
    """The sizes no sample uses but a drop-in has to survive: a GAS without triangles (every ray misses, nothing is occluded), a ray
    buffer without rays, an IAS without instances, and GAS of one / two / three triangles (root = leaf; the smallest inner nodes) on both
    hierarchies — hit records equal the oracle's."""
    from optix_raytracer_b200 import host
    rng = np.random.default_rng(3)
    rays = common.random_rays(rng, 4096, [-2, -2, -2], [2, 2, 2])
    d_rays = ctx.to_device(rays)
    empty = ctx.build_accel([ctx.triangle_input(torch.zeros((0, 3), dtype=torch.float32, device=ctx.torch_device), vertex_stride=12)])
    assert empty.info().num_triangles == 0
    got = host.ext_hits_to_numpy(ctx.trace_closest(empty, d_rays))
    assert (got["t"] < 0).all() and not ctx.trace_any(empty, d_rays).cpu().numpy().any()
    ias0 = ctx.build_accel([ctx.instance_input([])], compact=False)
    assert (host.ext_hits_to_numpy(ctx.trace_closest(ias0, d_rays))["t"] < 0).all()
    none = ctx.to_device(np.zeros((0, 8), np.float32))
    for n in (1, 2, 3, 5):
        c = rng.random((n, 1, 3), dtype=np.float32) * 2 - 1
        tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 1.5).astype(np.float32)
        scene = orc.Scene(tris)
        ref = scene.trace(rays)
        for hier in ("lbvh", "ploc"):
            import os
            os.environ["B200RT_HIERARCHY"] = hier
            try:
                acc = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12)])
            finally:
                os.environ.pop("B200RT_HIERARCHY", None)
            _assert_hits_equal(host.ext_hits_to_numpy(ctx.trace_closest(acc, d_rays)), ref, f"{n} triangles, {hier}")
            assert ctx.trace_closest(acc, none).shape[0] == 0
            ias = ctx.build_accel([ctx.instance_input([(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), 0, acc), (np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), 0, empty)])],
                                  compact=False)
            got = host.ext_hits_to_numpy(ctx.trace_closest(ias, d_rays))
            assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32)), f"{n} triangles instanced next to an empty GAS"
        assert (ref["t"] >= 0).any()
