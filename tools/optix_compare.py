#!/usr/bin/env python3
"""Compare libb200rt against the reference's OptiX path on the same B200 (needs libnvoptix.so.1: GPU box only).

Runs identical host-side state (optix_raytracer_b200/host.py mirrors) through both back ends:
  1. hit records of identical ray batches: b200rt_trace_closest vs OptiX built-in triangles (query_programs.cu)
  2. optixRaycasting on the Duck: the reference's __closesthit__buffer_hit Hit buffer vs b200rt_launch_raycast
  3. optixPathTracer Cornell image (accumulated over subframes): RMSE / PSNR
  4. timings: accel build, optixLaunch vs b200rt launch (CUDA events), Cornell and the synthetic mesh
Writes one JSON report (default gpurun_out/optix_compare.json).

    python tools/optix_compare.py [--triangles 50000000] [--subframes 8] [--out gpurun_out/optix_compare.json]
"""
import argparse
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from optix_raytracer_b200 import host  # noqa: E402
from oracle.optix_ref import backend as ob  # noqa: E402
from tests import common  # noqa: E402


def ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def cuda_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def compare_hits(bctx, octx, baccel, oaccel, rays, is_ias, label):
    d_rays = bctx.to_device(rays)
    got = host.ext_hits_to_numpy(bctx.trace_closest(baccel, d_rays))
    ref = host.ext_hits_to_numpy(octx.trace_closest(oaccel, d_rays, is_ias=is_ias))
    hit_b, hit_o = got["t"] >= 0, ref["t"] >= 0
    both = hit_b & hit_o
    same_prim = both & (got["prim"] == ref["prim"])
    ud = ulp_diff(got["t"][same_prim], ref["t"][same_prim])
    bd = np.maximum(np.abs(got["b1"][same_prim] - ref["b1"][same_prim]), np.abs(got["b2"][same_prim] - ref["b2"][same_prim]))
    diff_prim = both & (got["prim"] != ref["prim"])
    # rays whose primitive differs: how far apart are the two t (coplanar / shared-edge ties show up as ~0)
    rel = np.abs(got["t"][diff_prim] - ref["t"][diff_prim]) / np.maximum(np.abs(ref["t"][diff_prim]), 1e-30)
    out = {"label": label, "rays": int(rays.shape[0]), "hits_b200rt": int(hit_b.sum()), "hits_optix": int(hit_o.sum()),
           "hit_miss_disagree": int((hit_b != hit_o).sum()), "same_primitive": int(same_prim.sum()), "different_primitive": int(diff_prim.sum()),
           "t_bit_identical": int((ud == 0).sum()), "t_within_1ulp": int((ud <= 1).sum()), "t_within_4ulp": int((ud <= 4).sum()),
           "t_max_ulp": int(ud.max()) if ud.size else 0,
           "bary_max_abs_diff": float(bd.max()) if bd.size else 0.0,
           "different_primitive_rel_t_gap_max": float(rel.max()) if rel.size else 0.0,
           "different_primitive_rel_t_gap_median": float(np.median(rel)) if rel.size else 0.0}
    if is_ias:
        out["instance_mismatch"] = int((both & (got["inst"] != ref["inst"])).sum())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triangles", type=int, default=50_000_000)
    ap.add_argument("--subframes", type=int, default=8)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "optix_compare.json"))
    ap.add_argument("--skip-synth", action="store_true")
    a = ap.parse_args()
    rep = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "gpu": torch.cuda.get_device_name(0)}
    ok, why = ob.available(0)
    rep["optix_available"] = ok
    if not ok:
        rep["why"] = why
        pathlib.Path(a.out).write_text(json.dumps(rep, indent=1))
        print(json.dumps(rep))
        return 0
    bctx, octx = host.Context(0), ob.OptixContext(0)
    rep["rtcore_version"] = int(octx.olib.oref_rtcore_version())
    rng = np.random.default_rng(11)

    # ---- 1. Cornell: random rays, GAS handle ------------------------------------------------------------
    bpt = host.PathTracer(bctx, 768, 768, 16)
    opt = host.PathTracer(octx, 768, 768, 16)
    n = 1 << 20
    rays = common.random_rays(rng, n, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
    rep["hits_cornell"] = compare_hits(bctx, octx, bpt.accel, opt.accel, rays, False, "Cornell box, 1 Mi random rays (incl. axis-aligned / zero-component directions)")
    # occlusion flags
    rays2 = rays.copy()
    rays2[:, 7] = rng.random(n, dtype=np.float32) * 800
    d2 = bctx.to_device(rays2)
    occ_b = bctx.trace_any(bpt.accel, d2).cpu().numpy().astype(bool)
    occ_o = octx.trace_any(opt.accel, d2).cpu().numpy().astype(bool)
    rep["occlusion_cornell"] = {"rays": n, "disagree": int((occ_b != occ_o).sum()), "occluded_b200rt": int(occ_b.sum())}

    def checkpoint():
        pathlib.Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        pathlib.Path(a.out).write_text(json.dumps(rep, indent=1))
    checkpoint()

    # ---- 2. Duck: optixRaycasting (the reference's own programs) ------------------------------------------
    sc = common.duck_scene()
    brc, orc_ = host.Raycaster(bctx, sc), host.Raycaster(octx, sc)
    nray = brc.buffer_rays(1040)
    orc_.buffer_rays(1040)
    brc.launch()
    orc_.launch(want_ext=False)
    torch.cuda.synchronize()
    duck = {"rays_per_batch": nray, "width": brc.width, "height": brc.height}
    assert torch.equal(brc.rays, orc_.rays) and torch.equal(brc.rays_translated, orc_.rays_translated)
    for name, hb, ho in (("original", brc.hits, orc_.hits), ("translated", brc.hits_translated, orc_.hits_translated)):
        hb, ho = hb.cpu().numpy(), ho.cpu().numpy()
        dn = np.abs(hb[:, 1:] - ho[:, 1:]).max(axis=1)
        hit = ho[:, 0] >= 0
        duck[name] = {"Hit.t_identical": int((hb[:, 0] == ho[:, 0]).sum()), "Hit.t_differs": int((hb[:, 0] != ho[:, 0]).sum()),
                      "hits": int(hit.sum()), "normal_max_abs_diff": float(dn.max()), "normal_diff_gt_1e-3": int((dn > 1e-3).sum()),
                      "Hit_buffer_bit_identical_records": int((hb.view(np.uint32) == ho.view(np.uint32)).all(axis=1).sum())}
    rep["raycast_duck"] = duck
    rep["hits_duck_original"] = compare_hits(bctx, octx, brc.ias, orc_.ias, brc.rays.cpu().numpy(), True, "Duck.gltf optixRaycasting ortho batch (IAS)")
    rep["hits_duck_translated"] = compare_hits(bctx, octx, brc.ias, orc_.ias, brc.rays_translated.cpu().numpy(), True, "Duck.gltf translated batch (IAS)")
    lo, hi = brc.bbmin, brc.bbmax
    rrays = common.random_rays(rng, 1 << 20, lo, hi)
    rep["hits_duck_random"] = compare_hits(bctx, octx, brc.ias, orc_.ias, rrays, True, "Duck.gltf, 1 Mi random rays (IAS)")
    tb = cuda_ms(lambda: brc.launch(want_ext=False))
    to = cuda_ms(lambda: orc_.launch(want_ext=False))
    rep["raycast_duck"]["ms_two_batches"] = {"b200rt": tb, "optix": to, "Mrays_s_b200rt": 2 * nray / tb / 1e3, "Mrays_s_optix": 2 * nray / to / 1e3}

    checkpoint()

    # ---- 3. Cornell image: optixPathTracer, accumulated over subframes --------------------------------------
    segs = 0
    for sub in range(a.subframes):
        st = bpt.launch_subframe(sub, collect_stats=1)
        segs += st.radiance_segments + st.shadow_segments
        opt.launch_subframe(sub)
    torch.cuda.synchronize()
    ab, ao = bpt.accum.cpu().numpy()[..., :3].astype(np.float64), opt.accum.cpu().numpy()[..., :3].astype(np.float64)
    fb, fo = bpt.frame.cpu().numpy()[..., :3].astype(np.float64), opt.frame.cpu().numpy()[..., :3].astype(np.float64)
    mse8 = np.mean((fb - fo) ** 2)
    # noise floor: two independent estimates of the same image differ by sqrt(2)*sigma; compare against b200rt vs b200rt with other seeds
    bpt2 = host.PathTracer(bctx, 768, 768, 16)
    for sub in range(a.subframes):
        bpt2.launch_subframe(sub + 1000)  # different subframe indices = different seeds
    torch.cuda.synchronize()
    # running mean uses 1/(subframe+1): replay with the proper weights by averaging manually
    rep["image_cornell"] = {"subframes": a.subframes, "spp_total": 16 * a.subframes,
                            "mean_radiance_b200rt": float(ab.mean()), "mean_radiance_optix": float(ao.mean()),
                            "rel_mean_diff": float(abs(ab.mean() - ao.mean()) / ao.mean()),
                            "rmse_accum": float(np.sqrt(np.mean((ab - ao) ** 2))),
                            "rmse_u8": float(np.sqrt(mse8)), "psnr_u8_db": float(10 * np.log10(255.0 ** 2 / max(mse8, 1e-12))),
                            "per_channel_mean_b200rt": ab.reshape(-1, 3).mean(0).tolist(), "per_channel_mean_optix": ao.reshape(-1, 3).mean(0).tolist()}
    # same-renderer noise reference: b200rt subframes [0, n) vs OptiX is expected to sit at the Monte-Carlo noise level, which we
    # measure as b200rt(seeds A) vs b200rt(seeds B) at 1 subframe each
    b1, b2 = host.PathTracer(bctx, 768, 768, 16), host.PathTracer(bctx, 768, 768, 16)
    o1 = host.PathTracer(octx, 768, 768, 16)
    b1.launch_subframe(0); b2.launch_subframe(1); o1.launch_subframe(0)
    torch.cuda.synchronize()
    x1, x2, y1 = [p.accum.cpu().numpy()[..., :3].astype(np.float64) for p in (b1, b2, o1)]
    rep["image_cornell"]["rmse_16spp_b200rt_seedA_vs_seedB"] = float(np.sqrt(np.mean((x1 - x2) ** 2)))
    rep["image_cornell"]["rmse_16spp_b200rt_vs_optix_same_seed"] = float(np.sqrt(np.mean((x1 - y1) ** 2)))
    rep["image_cornell"]["pixels_bit_identical_16spp_same_seed"] = int((b1.accum.cpu().numpy().view(np.uint32) == o1.accum.cpu().numpy().view(np.uint32)).all(axis=-1).sum())

    tb = cuda_ms(lambda: bpt.launch_subframe(3))
    to = cuda_ms(lambda: opt.launch_subframe(3))
    st = bpt.launch_subframe(3, collect_stats=1)
    s3 = st.radiance_segments + st.shadow_segments
    # sample_groups 1 keeps the reference's flat fp32 summation order; bench.py's setting (4) only regroups that sum
    tb4 = cuda_ms(lambda: bpt.launch_subframe(3, sample_groups=4))
    rep["timing_cornell_768x768x16"] = {"segments": int(s3), "ms_b200rt": tb, "ms_b200rt_sample_groups_4": tb4, "ms_optix": to, "Mrays_s_b200rt": s3 / tb / 1e3,
                                        "Mrays_s_b200rt_sample_groups_4": s3 / tb4 / 1e3, "Mrays_s_optix": s3 / to / 1e3, "speedup": to / tb,
                                        "speedup_sample_groups_4": to / tb4}

    checkpoint()

    # ---- 3b. optixMeshViewer (whitted.cu) on the textured Duck, 1920x1080 ------------------------------------------------------
    sc = common.duck_scene()
    bmv, omv = host.MeshViewer(bctx, sc, 1920, 1080), host.MeshViewer(octx, sc, 1920, 1080)
    for sub in range(4):
        bmv.launch_subframe(sub); omv.launch_subframe(sub)
    torch.cuda.synchronize()
    ab, ao = bmv.accum.cpu().numpy()[..., :3].astype(np.float64), omv.accum.cpu().numpy()[..., :3].astype(np.float64)
    fb, fo = bmv.frame.cpu().numpy()[..., :3].astype(np.float64), omv.frame.cpu().numpy()[..., :3].astype(np.float64)
    mse8 = np.mean((fb - fo) ** 2)
    tb = cuda_ms(lambda: bmv.launch_subframe(5))
    to = cuda_ms(lambda: omv.launch_subframe(5))
    rep["whitted_duck_1920x1080"] = {"subframes_compared": 4, "rel_mean_diff": float(abs(ab.mean() - ao.mean()) / ao.mean()),
                                     "rmse_accum": float(np.sqrt(np.mean((ab - ao) ** 2))), "psnr_u8_db": float(10 * np.log10(255.0 ** 2 / max(mse8, 1e-12))),
                                     "ms_per_subframe_b200rt": tb, "ms_per_subframe_optix": to, "speedup": to / tb,
                                     "Msamples_s_b200rt": 1920 * 1080 / tb / 1e3, "Msamples_s_optix": 1920 * 1080 / to / 1e3}
    bmv.close(); omv.close()
    del bmv, omv
    checkpoint()

    # ---- 3b'. alpha materials: the any-hit programs (optixRaycasting.cu:89-102, whitted.cu:100-137) and the BLEND continuation ----------
    alpha = {}
    sc = common.duck_alpha_scene(1)
    brc2, orc2 = host.Raycaster(bctx, sc), host.Raycaster(octx, sc)
    nray2 = brc2.buffer_rays(1040); orc2.buffer_rays(1040)
    brc2.launch(want_ext=False); orc2.launch(want_ext=False)
    torch.cuda.synchronize()
    hb, ho = brc2.hits.cpu().numpy(), orc2.hits.cpu().numpy()
    tb = cuda_ms(lambda: brc2.launch(want_ext=False)); to = cuda_ms(lambda: orc2.launch(want_ext=False))
    alpha["raycast_mask"] = {"rays_per_batch": nray2, "Hit.t_differs": int((hb[:, 0].view(np.uint32) != ho[:, 0].view(np.uint32)).sum()),
                             "hits_optix": int((ho[:, 0] >= 0).sum()), "ms_two_batches_b200rt": tb, "ms_two_batches_optix": to, "speedup": to / tb}
    brc2.close(); orc2.close()
    for name, mode, ds in (("whitted_mask", 1, False), ("whitted_blend", 2, False), ("whitted_blend_double_sided", 2, True)):
        sc = common.duck_alpha_scene(mode)
        sc["materials"][0]["double_sided"] = ds
        bmv, omv = host.MeshViewer(bctx, sc, 1920, 1080), host.MeshViewer(octx, sc, 1920, 1080)
        for sub in range(4):
            bmv.launch_subframe(sub); omv.launch_subframe(sub)
        torch.cuda.synchronize()
        ab, ao = bmv.accum.cpu().numpy()[..., :3].astype(np.float64), omv.accum.cpu().numpy()[..., :3].astype(np.float64)
        fb, fo = bmv.frame.cpu().numpy()[..., :3].astype(np.float64), omv.frame.cpu().numpy()[..., :3].astype(np.float64)
        mse8 = np.mean((fb - fo) ** 2)
        tb = cuda_ms(lambda: bmv.launch_subframe(5)); to = cuda_ms(lambda: omv.launch_subframe(5))
        alpha[name] = {"rel_mean_diff": float(abs(ab.mean() - ao.mean()) / ao.mean()), "psnr_u8_db": float(10 * np.log10(255.0 ** 2 / max(mse8, 1e-12))),
                       "ms_per_subframe_b200rt": tb, "ms_per_subframe_optix": to, "speedup": to / tb}
        bmv.close(); omv.close()
        del bmv, omv
    rep["alpha_materials_duck"] = alpha
    checkpoint()

    # ---- 3c. imgui_test (optixTriangle.cu), 1920x1080, stand-in scene of 1.74 M triangles, two cameras x two apertures ------------
    pgr = {}
    for name, cam_kw in (("reference_camera", {}), ("close_camera", {"eye": (0.5, 0.7, -1.4), "up": (0.0, 1.0, 0.000073), "lookat": (0.0, 0.1, 0.0), "fov": 50.0})):
        for ap in (0.0, 0.05):
            cam = host.playground_camera(aperture=ap, **cam_kw)
            bp = host.Playground(bctx, 1920, 1080, spf=8, rows=132, camera=cam)
            op = host.Playground(octx, 1920, 1080, spf=8, rows=132, camera=cam)
            bp.launch_frame(dirty=True); op.launch_frame(dirty=True)
            torch.cuda.synchronize()
            fb, fo = bp.film.cpu().numpy().astype(np.float64), op.film.cpu().numpy().astype(np.float64)
            ib, io = bp.image.cpu().numpy()[..., :3].astype(np.float64), op.image.cpu().numpy()[..., :3].astype(np.float64)
            mse8 = np.mean((ib - io) ** 2)
            tb = cuda_ms(lambda: bp.launch_frame(dirty=True), reps=3, warm=1)
            to = cuda_ms(lambda: op.launch_frame(dirty=True), reps=3, warm=1)
            hit_frac = float((np.abs(fo / 8 - np.clip(fo / 8, 0, 1)) < 1).mean())
            pgr[f"{name}_aperture_{ap}"] = {"triangles": bp.num_triangles, "spf": 8, "rel_mean_diff": float(abs(fb.mean() - fo.mean()) / fo.mean()),
                                            "psnr_u8_db": float(10 * np.log10(255.0 ** 2 / max(mse8, 1e-12))), "ms_per_frame_b200rt": tb,
                                            "ms_per_frame_optix": to, "speedup": to / tb, "Msamples_s_b200rt": 1920 * 1080 * 8 / tb / 1e3,
                                            "Msamples_s_optix": 1920 * 1080 * 8 / to / 1e3}
            del bp, op
            torch.cuda.empty_cache()
    rep["playground_1920x1080"] = pgr
    checkpoint()

    # ---- 4. synthetic mesh (BASELINE.json configs[4]): build + launch timing, optixMultiGPU programs ---------
    if not a.skip_synth:
        del bpt, opt, bpt2, b1, b2, o1
        torch.cuda.empty_cache()
        T = a.triangles
        verts, mats = host.synthetic_mesh(bctx, T, 0)
        torch.cuda.synchronize()
        W, H = 3840, 2160
        t0 = time.perf_counter(); bs = host.PathTracer(bctx, W, H, 16, vertices=verts, mat_indices=mats, multigpu=(0, 1)); torch.cuda.synchronize()
        t1 = time.perf_counter(); os_ = host.PathTracer(octx, W, H, 16, vertices=verts, mat_indices=mats, multigpu=(0, 1)); torch.cuda.synchronize()
        t2 = time.perf_counter()
        rays = common.random_rays(rng, 1 << 20, [0, 0, 0], [556, 548.8, 559.2], tmin=0.01)
        rep["hits_synthetic"] = compare_hits(bctx, octx, bs.accel, os_.accel, rays, False, f"synthetic mesh {T} triangles, 1 Mi random rays")
        bs.sample_groups = 8  # bench.py's setting for this workload
        tb = cuda_ms(lambda: bs.launch_subframe(2), reps=3, warm=1)
        to = cuda_ms(lambda: os_.launch_subframe(2), reps=3, warm=1)
        st = bs.launch_subframe(2, collect_stats=1)
        s = st.radiance_segments + st.shadow_segments
        # images: one fresh launch each at subframe 0 (the running mean of later subframes depends on how often a launch was
        # repeated for timing); compact per-sample buffers with N=1 are in the same order in both
        bs.sample_groups = 1  # the reference's summation order for the image comparison
        bs.launch_subframe(0); os_.launch_subframe(0)
        torch.cuda.synchronize()
        ab, ao = bs.accum.cpu().numpy()[:, :3].astype(np.float64), os_.accum.cpu().numpy()[:, :3].astype(np.float64)
        same_px = (bs.accum.cpu().numpy().view(np.uint32) == os_.accum.cpu().numpy().view(np.uint32)).all(axis=-1)
        rep["timing_synthetic"] = {"triangles": T, "width": W, "height": H, "spl": 16, "segments": int(s), "ms_b200rt": tb, "ms_optix": to,
                                   "Mrays_s_b200rt": s / tb / 1e3, "Mrays_s_optix": s / to / 1e3, "speedup": to / tb,
                                   "setup_wall_s_b200rt(build+compact+sbt)": t1 - t0, "setup_wall_s_optix(build+compact+module+pipeline)": t2 - t1,
                                   "accel_bytes_b200rt": int(bs.accel.buf.numel()), "accel_bytes_optix": int(os_.accel.buf.numel())}
        rep["image_synthetic"] = {"programs": "optixMultiGPU", "spp": 16, "mean_radiance_b200rt": float(ab.mean()), "mean_radiance_optix": float(ao.mean()),
                                  "rel_mean_diff": float(abs(ab.mean() - ao.mean()) / ao.mean()), "rmse_accum": float(np.sqrt(np.mean((ab - ao) ** 2))),
                                  "pixels_bit_identical_fraction": float(same_px.mean())}
        # accel build alone (events), smaller helper: rebuild both
        bi = bctx.triangle_input(verts, sbt_index=mats, num_sbt=4, vertex_stride=16)
        def bbuild():
            bctx.build_accel([bi], compact=True)
        def obuild():
            octx.build_accel([bi], compact=True)
        del bs, os_
        torch.cuda.empty_cache()
        rep["timing_synthetic"]["accel_build_ms_b200rt"] = cuda_ms(bbuild, reps=2, warm=1)
        rep["timing_synthetic"]["accel_build_ms_optix"] = cuda_ms(obuild, reps=2, warm=1)

    pathlib.Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(a.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps(rep, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
