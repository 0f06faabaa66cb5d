"""GPU parity against the REAL reference: the reference's own OptiX device programs (PTX compiled from
/root/reference/SDK/optix*/...cu, oracle/_ref/) run by libnvoptix.so.1 on the same B200, driven with the
same host state as libb200rt.so (oracle/optix_ref/).  This is the check BASELINE.json's north_star asks for:

  * primary-ray hit records on identical ray batches: same hit/miss, same primitive, same instance, t within ulps
    (OptiX's triangle test is closed source, so t is compared by tolerance: >= 98 % of rays within 4 ulp);
  * optixRaycasting's Hit buffer: the reference's __closesthit__buffer_hit stores float(unsigned(t)) — identical
    for every ray — and the interpolated normal (fast-math build: <= 1e-5 absolute);
  * rendered images at the same seeds: relative mean radiance < 1e-3 and PSNR of the sRGB frame > 55 dB.

Skipped (with the reason) where OptiX cannot be initialised, e.g. a box without libnvoptix.so.1.
Measured values of a full-size run are committed in profiles/r01_optix_compare.json (tools/optix_compare.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from tests import common  # noqa: E402


@pytest.fixture(scope="module")
def ctxs():
    from optix_raytracer_b200 import host
    from oracle.optix_ref import backend as ob
    ok, why = ob.available(0)
    if not ok:
        pytest.skip(f"OptiX reference not usable here: {why}")
    octx = ob.OptixContext(0)
    # the libnvoptix.so.1 mapped into this process must be the DRIVER's: with optix_raytracer_b200/optix_shim on LD_LIBRARY_PATH these
    # tests would compare b200rt with itself
    mapped = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if "libnvoptix" in ln})
    assert mapped, "OptiX initialised but no libnvoptix is mapped?"
    assert not any("optix_shim" in m for m in mapped), f"the OptiX parity tests are running on the shim, not on the driver's OptiX: {mapped}"
    assert not ob.shim_active()
    return host.Context(0), octx


def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def _compare(bctx, octx, baccel, oaccel, rays, is_ias):
    from optix_raytracer_b200 import host
    d = bctx.to_device(rays)
    got = host.ext_hits_to_numpy(bctx.trace_closest(baccel, d))
    ref = host.ext_hits_to_numpy(octx.trace_closest(oaccel, d, is_ias=is_ias))
    hb, ho = got["t"] >= 0, ref["t"] >= 0
    assert ho.mean() > 0.2
    assert np.array_equal(hb, ho), f"{(hb != ho).sum()} rays disagree on hit/miss"
    return got, ref, ho


def test_cornell_hit_records_match_optix(ctxs):
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    b, o = host.PathTracer(bctx, 32, 32, 1), host.PathTracer(octx, 32, 32, 1)
    rng = np.random.default_rng(5)
    rays = common.random_rays(rng, 1 << 18, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
    got, ref, hit = _compare(bctx, octx, b.accel, o.accel, rays, False)
    assert np.array_equal(got["prim"][hit], ref["prim"][hit]), "primitive index differs from OptiX"
    u = _ulps(got["t"][hit], ref["t"][hit])
    assert (u <= 4).mean() >= 0.98, f"only {(u <= 4).mean():.4f} of the hits have t within 4 ulp of OptiX"
    assert np.abs(got["b1"][hit] - ref["b1"][hit]).max() < 1e-4 and np.abs(got["b2"][hit] - ref["b2"][hit]).max() < 1e-4
    # occlusion rays (TERMINATE_ON_FIRST_HIT) with finite tmax
    rays[:, 7] = rng.random(rays.shape[0], dtype=np.float32) * 800
    d = bctx.to_device(rays)
    ob_, oo = bctx.trace_any(b.accel, d).cpu().numpy().astype(bool), octx.trace_any(o.accel, d).cpu().numpy().astype(bool)
    assert (ob_ != oo).mean() < 1e-5


def test_duck_raycasting_matches_the_reference_programs_on_optix(ctxs):
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    sc = common.duck_scene()
    b, o = host.Raycaster(bctx, sc), host.Raycaster(octx, sc)
    # BASELINE.json configs[1] at full size: the sample's default width gives two batches of 1 006 720 rays
    assert b.buffer_rays(1040) > 1_000_000
    o.buffer_rays(1040)
    b.launch()
    o.launch(want_ext=False)
    torch.cuda.synchronize()
    for hb, ho in ((b.hits, o.hits), (b.hits_translated, o.hits_translated)):
        hb, ho = hb.cpu().numpy(), ho.cpu().numpy()
        assert (ho[:, 0] >= 0).mean() > 0.2
        assert np.array_equal(hb[:, 0].view(np.uint32), ho[:, 0].view(np.uint32)), "Hit.t (= float(unsigned(t))) differs from the reference"
        assert np.abs(hb[:, 1:] - ho[:, 1:]).max() <= 1e-5, "Hit.geom_normal differs from the reference"
    # extended records through OptiX's built-in triangles on the same IAS: primitive + instance identical
    got, ref, hit = _compare(bctx, octx, b.ias, o.ias, b.rays.cpu().numpy(), True)
    assert np.array_equal(got["prim"][hit], ref["prim"][hit]) and np.array_equal(got["inst"][hit], ref["inst"][hit])
    assert (_ulps(got["t"][hit], ref["t"][hit]) <= 8).all()


def test_alpha_mask_raycasting_matches_the_reference_anyhit_on_optix(ctxs):
    """optixRaycasting's default input is DuckHole.gltf (alphaMode MASK): __anyhit__texture_mask (optixRaycasting.cu:89-102) discards
    hits whose base-colour alpha is below the cutoff, so rays pass through the holes and hit what lies behind."""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    sc = common.duck_alpha_scene(1)
    b, o = host.Raycaster(bctx, sc), host.Raycaster(octx, sc)
    opaque = host.Raycaster(bctx, common.duck_scene())
    for r in (b, o, opaque):
        r.buffer_rays(1040)
    b.launch()
    o.launch(want_ext=False)
    opaque.launch()
    torch.cuda.synchronize()
    hb, ho, hq = b.hits.cpu().numpy(), o.hits.cpu().numpy(), opaque.hits.cpu().numpy()
    changed = (hq[:, 0] != ho[:, 0]).mean()
    assert changed > 0.02, f"the mask changes only {changed:.4f} of the rays: the test would not see a missing any-hit program"
    # texels right at the cutoff may fall either way (hardware bilinear filtering at 8-bit weights is identical, but t and the
    # barycentrics differ by ulps between the two triangle tests): a handful of rays, never a pattern
    differ = hb[:, 0].view(np.uint32) != ho[:, 0].view(np.uint32)
    assert differ.mean() < 2e-4, f"{differ.sum()} of {differ.size} rays disagree with the reference any-hit program"
    same = ~differ & (ho[:, 0] >= 0)
    assert np.abs(hb[same, 1:] - ho[same, 1:]).max() <= 1e-5
    for r in (b, o, opaque):
        r.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_cornell_image_matches_optix_at_the_same_seeds(ctxs, mode):
    """optixPathTracer (mode 0) / optixMultiGPU (mode 1) device programs on OptiX vs the wavefront restatement."""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    mg = (0, 1) if mode else None
    # BASELINE.json configs[0] at full size: 768 x 768, 16 samples per launch
    b, o = host.PathTracer(bctx, 768, 768, 16, multigpu=mg), host.PathTracer(octx, 768, 768, 16, multigpu=mg)
    for sub in range(2):
        b.launch_subframe(sub)
        o.launch_subframe(sub)
    torch.cuda.synchronize()
    ab = b.accum.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64)
    ao = o.accum.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64)
    assert ao.mean() > 0.05
    assert abs(ab.mean() - ao.mean()) / ao.mean() < 1e-3
    fb, fo = b.frame.cpu().numpy()[..., :3].astype(np.float64), o.frame.cpu().numpy()[..., :3].astype(np.float64)
    mse = np.mean((fb - fo) ** 2)
    psnr = 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))
    assert psnr > 55.0, f"PSNR {psnr:.1f} dB"


def test_synthetic_mesh_hits_match_optix(ctxs):
    """2 M-triangle instance of the BASELINE configs[4] scene: shared edges / tiny triangles; the few primitive
    differences must be ties (both engines report the same t to within 1e-6 relative)."""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    verts, mats = host.synthetic_mesh(bctx, 2_000_000, 0)
    b = host.PathTracer(bctx, 32, 32, 1, vertices=verts, mat_indices=mats, multigpu=(0, 1))
    o = host.PathTracer(octx, 32, 32, 1, vertices=verts, mat_indices=mats, multigpu=(0, 1))
    rays = common.random_rays(np.random.default_rng(9), 1 << 18, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
    got, ref, hit = _compare(bctx, octx, b.accel, o.accel, rays, False)
    diff = hit & (got["prim"] != ref["prim"])
    assert diff.sum() <= 1e-4 * hit.sum()
    if diff.any():
        rel = np.abs(got["t"][diff] - ref["t"][diff]) / np.abs(ref["t"][diff])
        assert rel.max() < 1e-6


@pytest.mark.parametrize("aperture", [0.0, 0.05])
def test_playground_matches_the_reference_program_on_optix(ctxs, aperture):
    """imgui_test's own optixTriangle.cu programs on OptiX vs the wavefront restatement, same Params / Camera / LightVariant bytes."""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    cam = host.playground_camera(eye=(0.3, 0.6, -1.2), up=(0.0, 1.0, 0.000073), lookat=(0.0, 0.1, 0.0), fov=50.0, aperture=aperture)
    b = host.Playground(bctx, 256, 192, spf=4, rows=24, camera=cam)
    o = host.Playground(octx, 256, 192, spf=4, rows=24, camera=cam)
    for dirty in (True, False):
        b.launch_frame(dirty=dirty)
        o.launch_frame(dirty=dirty)
    torch.cuda.synchronize()
    fb, fo = b.film.cpu().numpy().astype(np.float64), o.film.cpu().numpy().astype(np.float64)
    assert fo.mean() > 0.1
    assert abs(fb.mean() - fo.mean()) / fo.mean() < 1e-3
    ib, io = b.image.cpu().numpy()[..., :3].astype(np.float64), o.image.cpu().numpy()[..., :3].astype(np.float64)
    mse = np.mean((ib - io) ** 2)
    psnr = 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))
    assert psnr > 45.0, f"PSNR {psnr:.1f} dB"


@pytest.mark.parametrize("variant", ["duck_texture", "all_maps", "alpha_mask", "alpha_blend", "alpha_blend_double_sided"])
def test_whitted_matches_the_reference_programs_on_optix(ctxs, variant):
    """optixMeshViewer: SDK/cuda/whitted.cu on OptiX vs the wavefront restatement on the textured Duck — same LaunchParams, same
    SBT records (GeometryData + MaterialData with the same cudaTextureObject_t handles), hardware tex2D in both.  `all_maps` adds
    procedural metallic-roughness / emissive / normal textures so every sampleTexture path of the closest-hit program runs.  The
    alpha variants give the base-colour texture an alpha pattern (holes, solid, ramps) and set alphaMode MASK / BLEND: radiance rays go
    through __anyhit__radiance, shadow rays through __anyhit__occlusion (pending attenuation), BLEND hits continue behind themselves
    (whitted.cu:100-137,266-286); doubleSided lets the continuation hit the inside of the duck (deeper chains)."""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    sc = common.duck_scene()
    if variant.startswith("alpha"):
        sc = common.duck_alpha_scene(1 if variant == "alpha_mask" else 2)
        sc["materials"][0]["double_sided"] = variant.endswith("double_sided")
    if variant == "all_maps":
        rng = np.random.default_rng(3)
        yy, xx = np.mgrid[0:64, 0:64]
        mr = np.stack([np.full((64, 64), 255), 96 + 8 * ((xx // 8 + yy // 8) % 2) * 15, 255 * ((xx // 16) % 2), np.full((64, 64), 255)], -1).astype(np.uint8)
        em = np.stack([(xx * 2) % 64, (yy * 3) % 64, np.zeros((64, 64)), np.full((64, 64), 255)], -1).astype(np.uint8)
        nm = np.stack([128 + 40 * np.sin(xx / 5.0), 128 + 40 * np.cos(yy / 7.0), np.full((64, 64), 230), np.full((64, 64), 255)], -1).astype(np.uint8)
        sc["images"] += [np.ascontiguousarray(mr), np.ascontiguousarray(em), np.ascontiguousarray(nm)]
        sc["textures"] += [{"sampler": 0, "source": 1}, {"sampler": 0, "source": 2}, {"sampler": 0, "source": 3}]
        t = lambda i: {"index": i, "texcoord": 0, "offset": [0.1, 0.2], "rotation": 0.3, "scale": [2.0, 3.0]}
        sc["materials"][0].update({"metallic": 1.0, "roughness": 1.0, "metallic_roughness_tex": t(1), "emissive_tex": t(2), "normal_tex": t(3),
                                   "emissive_factor": [0.5, 0.5, 0.5]})
    w, h = 320, 240
    b, o = host.MeshViewer(bctx, sc, w, h), host.MeshViewer(octx, sc, w, h)
    for sub in range(3):
        b.launch_subframe(sub)
        o.launch_subframe(sub)
    torch.cuda.synchronize()
    ab, ao = b.accum.cpu().numpy()[..., :3].astype(np.float64), o.accum.cpu().numpy()[..., :3].astype(np.float64)
    covered = (np.abs(ao - 0.1) > 1e-6).any(axis=-1)
    assert covered.mean() > (0.015 if variant.startswith("alpha") else 0.03)
    assert abs(ab.mean() - ao.mean()) / ao.mean() < 2e-3
    fb, fo = b.frame.cpu().numpy()[..., :3].astype(np.float64), o.frame.cpu().numpy()[..., :3].astype(np.float64)
    mse = np.mean((fb - fo) ** 2)
    psnr = 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))
    assert psnr > 45.0, f"PSNR {psnr:.1f} dB"
    b.close()
    o.close()


def test_face_culling_and_instance_flags_match_optix(ctxs):
    """OptixRayFlags CULL_BACK / CULL_FRONT_FACING_TRIANGLES, OptixInstanceFlags DISABLE_TRIANGLE_FACE_CULLING / FLIP_TRIANGLE_FACING, the
    geometry flag DISABLE_TRIANGLE_FACE_CULLING and the visibility mask on the real OptiX runtime against b200rt: same hit / miss, same
    primitive, same instance.  A wrong facing convention or a wrong flip would change about half of the culled hits.  (The OptiX query
    programs trace with OPTIX_RAY_FLAG_DISABLE_ANYHIT, which OptiX declares mutually exclusive with the CULL_*_ANYHIT ray flags, and
    the two face-cull flags exclude each other too: those combinations are not put to OptiX.)"""
    from optix_raytracer_b200 import host
    bctx, octx = ctxs
    rng = np.random.default_rng(29)
    n = 1500
    c = rng.random((n, 1, 3), dtype=np.float32) * 2 - 1
    tris = (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * 0.4).astype(np.float32)
    sbt = (rng.random(n) < 0.3).astype(np.uint32)
    rec_flags = [1, 1 | 4]   # both records DISABLE_ANYHIT (the query pipeline has no any-hit program); record 1 is exempt from face culling

    def xf(angle, scale, t):
        ca, sa = np.cos(angle), np.sin(angle)
        return np.array([[ca * scale, 0, sa * scale, t[0]], [0, scale, 0, t[1]], [-sa * scale, 0, ca * scale, t[2]]], np.float32).reshape(12)
    xfs = [xf(0.0, 1.0, (0, 0, 0)), xf(0.9, 0.8, (1.2, 0.3, 0)), xf(-0.6, 1.1, (-1.2, -0.4, 0.8)), xf(0.1, 3.0, (0, 0, 0))]
    iflags, masks = [0, 1, 2, 0], [1, 1, 1, 0]
    accels = []
    for ctx in (bctx, octx):
        gas = ctx.build_accel([ctx.triangle_input(ctx.to_device(tris.reshape(-1, 3)), vertex_stride=12, sbt_index=ctx.to_device(sbt), num_sbt=2,
                                                  flags=rec_flags)])
        ias = ctx.build_accel([ctx.instance_input([(m, 0, gas, k, f) for m, f, k in zip(xfs, iflags, masks)])], compact=False)
        accels.append((gas, ias))
    rays = common.random_rays(rng, 100_000, [-2.5, -2, -2], [2.5, 2, 2])
    d = bctx.to_device(rays)
    counts = {}
    for rf in (0, 16, 32):
        for which, is_ias in ((0, False), (1, True)):
            got = host.ext_hits_to_numpy(bctx.trace_closest(accels[0][which], d, ray_flags=rf | 1))
            ref = host.ext_hits_to_numpy(octx.trace_closest(accels[1][which], d, ray_flags=rf, is_ias=is_ias))
            hb, ho = got["t"] >= 0, ref["t"] >= 0
            what = f"ray_flags {rf:#x} {'ias' if is_ias else 'gas'}"
            # silhouette edges of a triangle soup: a ray within an ulp of an edge may fall either way between two triangle tests
            assert (hb != ho).mean() < 2e-4, f"{what}: {(hb != ho).sum()} rays disagree on hit / miss"
            both = hb & ho
            assert (got["prim"][both] != ref["prim"][both]).mean() < 2e-4, f"{what}: primitive differs from OptiX"
            assert (got["inst"][both] != ref["inst"][both]).mean() < 2e-4, f"{what}: instance differs from OptiX"
            counts[(rf, is_ias)] = int(ho.sum())
    for is_ias in (False, True):
        assert 0 < counts[(16, is_ias)] < counts[(0, is_ias)] and 0 < counts[(32, is_ias)] < counts[(0, is_ias)]
    # optixTrace's 8-bit visibility mask against OptixInstance::visibilityMask (B200RT_RAY_VISIBILITY_MASK: bits 16-23, XOR 1)
    masks2 = [1, 2, 6, 255]
    ias2 = [ctx.build_accel([ctx.instance_input([(m, 0, acc[0], k, f) for m, f, k in zip(xfs, iflags, masks2)])], compact=False) for ctx, acc in zip((bctx, octx), accels)]
    for vis in (1, 2, 4, 255, 0):
        rf = ((vis ^ 1) & 0xff) << 16
        got = host.ext_hits_to_numpy(bctx.trace_closest(ias2[0], d, ray_flags=rf | 1))
        ref = host.ext_hits_to_numpy(octx.trace_closest(ias2[1], d, ray_flags=rf, is_ias=True))
        hb, ho = got["t"] >= 0, ref["t"] >= 0
        assert (hb != ho).mean() < 2e-4, f"visibility mask {vis}: {(hb != ho).sum()} rays disagree on hit / miss"
        both = hb & ho
        assert both.any() == bool(vis)
        if vis:
            assert (got["inst"][both] != ref["inst"][both]).mean() < 2e-4 and (got["prim"][both] != ref["prim"][both]).mean() < 2e-4
        assert set(np.unique(ref["inst"][ho]).tolist()) == {k for k, m in enumerate(masks2) if m & vis}, f"visibility mask {vis}: instances OptiX reports"
