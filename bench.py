#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native path tracer (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W                  # b200rt arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own device programs on OptiX (libnvoptix.so.1)
                                                                    # on the same GPU; falls back to the CPU oracle port when OptiX
                                                                    # cannot be initialised (and says so in "reference_class")
    python bench.py --workload cornell|duck_raycast|whitted_duck|playground   # BASELINE.json configs[0..3]

Workload (default `synthetic`): BASELINE.json configs[4] — procedurally tessellated 50 M-triangle scene, 3840x2160,
samples_per_launch 16 (64 spp = 4 steps), optixMultiGPU device programs (depth cap 3), scene replicated per GPU, image split
with the reference's StaticWorkDistribution.  A "step" is one launch (one subframe) over the whole image by all ranks together.
Metric: Mrays/s = traced segments (radiance + shadow) of all ranks / max-over-ranks device time.  Both arms run the same host
mirror (optix_raytracer_b200/host.py: same build inputs, same Params bytes, same SBT records, same subframe indices) and divide
the same segment counts (counted once, untimed, by b200rt's instrumented launch: OptiX cannot count its own rays) by their
own CUDA-event times.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PT_WORKLOADS = ("synthetic", "cornell")


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200rt", choices=["b200rt", "reference"])
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "cornell", "duck_raycast", "whitted_duck", "playground"])
    ap.add_argument("--triangles", type=int, default=50_000_000)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spl", type=int, default=None, help="samples per launch (synthetic / cornell: 16; playground: samples per frame, 8)")
    ap.add_argument("--sample-groups", type=int, default=None, help="b200rt_pt_options.sample_groups: lanes per launch index (1 = reference summation order)")
    ap.add_argument("--ray-sort", type=int, default=None, help="b200rt_pt_options.ray_sort (default 0: measured slower end to end, profiles/r01_trace_kernel.md)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allgather"],
                    help="N > 1: how the frame is assembled — p2p: every rank's launch stores its pixels into rank 0's result buffer over NVLink "
                         "(the reference's single result buffer, optixMultiGPU.cpp:479-508); allgather: per-rank sample buffers all-gathered (NCCL) "
                         "and de-interleaved on every rank")
    ap.add_argument("--reference-class", default="auto", choices=["auto", "optix", "cpu"],
                    help="--impl reference: optix = the reference's device programs on libnvoptix (GPU), cpu = the oracle port on the host cores, "
                         "auto = optix when it can be initialised")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline sample")
    a = ap.parse_args(argv)
    if a.width is None:
        a.width, a.height = {"synthetic": (3840, 2160), "cornell": (768, 768), "duck_raycast": (1040, None), "whitted_duck": (1920, 1080),
                             "playground": (1920, 1080)}[a.workload]
    if a.spl is None:
        a.spl = {"synthetic": 16, "cornell": 16, "duck_raycast": 1, "whitted_duck": 1, "playground": 8}[a.workload]
    if a.ray_sort is None:
        a.ray_sort = 0
    if a.sample_groups is None:
        # lanes per launch index (fp32 summation order of a pixel's samples; the rays traced are the same for any value).  One GPU: 8 on the
        # bench scene (16 is 0.6 % faster for twice the lane state, 20 GB), 4 on the small ones.  With the image split over N > 1 GPUs a rank
        # holds 1/N of the launch indices, so 16 groups keep its launches as wide as one GPU's (and its lane state below one GPU's):
        # measured 8 GPUs 16026 -> 16538 Mrays/s, 4 GPUs 8272 -> 8459 (profiles/r02_scaling.md).
        world = int(os.environ.get("WORLD_SIZE", "1"))
        a.sample_groups = (8 if world == 1 else 16) if a.workload == "synthetic" else 4
    return a


def workload_name(a):
    return {
        "synthetic": f"synthetic tessellated mesh, {a.triangles} triangles, {a.width}x{a.height}, samples_per_launch {a.spl}, "
                     "optixMultiGPU programs (BASELINE.json configs[4])",
        "cornell": f"optixPathTracer Cornell box, {a.width}x{a.height}, samples_per_launch {a.spl} (BASELINE.json configs[0])",
        "duck_raycast": f"optixRaycasting on Duck.gltf, two batches of {a.width}-wide orthographic ray buffers (BASELINE.json configs[1])",
        "whitted_duck": f"optixMeshViewer (whitted.cu) on Duck.gltf, {a.width}x{a.height}, one subframe per step (BASELINE.json configs[2])",
        "playground": f"imgui_test stand-in scene (1.74 M triangles), {a.width}x{a.height}, {a.spl} samples per frame (BASELINE.json configs[3])",
    }[a.workload]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU legs: the scalar oracle on the host cores (cpu_baseline of the b200rt arm; the reference arm's fallback)
# ---------------------------------------------------------------------------------------------------
def oracle_scene_and_params(a):
    from oracle import pyoracle as orc
    from optix_raytracer_b200 import host
    sc = host.load_cornell()
    t0 = time.time()
    if a.workload == "synthetic":
        tris, mats = orc.synth_mesh(a.triangles, 0)
    else:
        tris, mats = sc["vertices"].reshape(-1, 3, 3), sc["mat_indices"]
    t1 = time.time()
    scene = orc.Scene(tris, mats)
    t2 = time.time()
    cam, lt = sc["camera"], sc["light"]
    U, V, W = orc.camera_uvw(cam["eye"], cam["lookat"], cam["up"], cam["fov_y"], a.width / float(a.height))
    p = orc.PTParams()
    p.subframe_index, p.width, p.height, p.samples_per_launch, p.nmat = 0, a.width, a.height, a.spl, 4
    p.mode = 1 if a.workload == "synthetic" else 0
    p.groups = a.sample_groups
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    p.eye, p.U, p.V, p.W = f3(cam["eye"]), f3(U), f3(V), f3(W)
    p.light_corner, p.light_v1, p.light_v2 = f3(lt["corner"]), f3(lt["v1"]), f3(lt["v2"])
    p.light_normal, p.light_emission, p.bg = f3(host.light_normal(lt["v1"], lt["v2"])), f3(lt["emission"]), f3([0, 0, 0])
    return orc, scene, p, sc, {"mesh_s": t1 - t0, "bvh_s": t2 - t1}


def cpu_strip(orc, scene, p, sc, a, y0, rows, accum):
    t = time.perf_counter()
    _, _, segs = scene.pathtrace(p, sc["emission_colors"], sc["diffuse_colors"], accum=accum, region=(0, y0, a.width, min(a.height, y0 + rows)),
                                 want_frame=False)
    return segs, time.perf_counter() - t


def run_cpu_baseline(a, seconds):
    """Bounded sample of the same workload on the host cores (kind 'port': the oracle restatement).  The path-tracing workloads only:
    the other workloads report the figure of the Cornell path tracer at their resolution (same traversal and shading code)."""
    if a.workload not in PT_WORKLOADS:
        b = argparse.Namespace(**vars(a))
        b.workload, b.width, b.height, b.spl, b.sample_groups = "cornell", 768, 768, 16, 1
        out = run_cpu_baseline(b, seconds)
        out["sample"] = "oracle port has no batch mode for this workload; sample = Cornell path tracer: " + out["sample"]
        return out
    orc, scene, p, sc, prep = oracle_scene_and_params(a)
    accum = np.zeros((a.height, a.width, 4), np.float32)
    rows = 8
    y = a.height // 2
    segs = 0
    spent = 0.0
    nstrips = 0
    while spent < seconds and y + rows <= a.height:
        s, dt = cpu_strip(orc, scene, p, sc, a, y, rows, accum)
        segs += s; spent += dt; y += rows; nstrips += 1
    return {"value": segs / spent / 1e6, "unit": "Mrays/s", "cores": orc.ncores(), "kind": "port",
            "sample": f"{nstrips} strips of {rows}x{a.width} pixels from row {a.height // 2}, {a.spl} spp, {segs} segments in {spent:.1f} s "
                      f"(oracle scene prep: mesh {prep['mesh_s']:.1f} s, BVH {prep['bvh_s']:.1f} s, untimed)"}


def run_reference_cpu(a, rank, why):
    """--impl reference without OptiX: the reference has no CPU renderer and its traversal lives in the closed libnvoptix, so the
    CPU arm is the oracle port (oracle/oracle.cpp) with all host threads; each step is a bounded strip of the image."""
    if rank != 0:
        return None
    if a.workload not in PT_WORKLOADS:
        return {"impl": "reference", "unavailable": f"OptiX unavailable ({why}) and the CPU port has no batch mode for workload {a.workload}"}
    orc, scene, p, sc, prep = oracle_scene_and_params(a)
    accum = np.zeros((a.height, a.width, 4), np.float32)
    rows = 8 if a.workload == "synthetic" else 64
    y = max(0, a.height // 2 - rows * (a.steps + a.warmup) // 2)
    for _ in range(a.warmup):
        cpu_strip(orc, scene, p, sc, a, y, rows, accum); y += rows
    segs, spent = 0, 0.0
    for _ in range(a.steps):
        s, dt = cpu_strip(orc, scene, p, sc, a, y, rows, accum)
        segs += s; spent += dt; y += rows
    val = segs / spent / 1e6
    sample = f"each step = one {rows}x{a.width}-pixel strip at {a.spl} spp ({segs // max(a.steps, 1)} segments/step)"
    return {"impl": "reference", "reference_class": "cpu", "reference_class_reason": why, "metric": "Mrays/s", "value": val, "unit": "Mrays/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": spent / max(a.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": orc.ncores(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


# ---------------------------------------------------------------------------------------------------
# GPU arms.  A Job is one workload set up on one context — host.Context (libb200rt.so) or oracle.optix_ref's OptixContext
# (libnvoptix.so.1): the host mirrors in host.py are back-end agnostic, so both arms execute the same Python below.
# ---------------------------------------------------------------------------------------------------
class Job:
    pt = None            # host.PathTracer of the path-tracing workloads
    frame = None         # device tensor(s) a step leaves behind (copied to pinned host memory in the e2e region)
    h2d_bytes = 0
    rays_note = ""

    def step(self, sub):
        raise NotImplementedError

    def rays(self, sub):
        raise NotImplementedError

    def close(self):
        pass


class PTJob(Job):
    def __init__(self, a, ctx, rank, world, host):
        hc = getattr(ctx, "helper", ctx)  # scene generation / fillSamples are plain CUDA kernels in the reference too
        if a.workload == "synthetic":
            verts, mats = host.synthetic_mesh(hc, a.triangles, 0)
            multigpu = (rank, world)
        else:
            verts = mats = None
            multigpu = (rank, world) if world > 1 else None
        self.multigpu = multigpu
        self.pt = host.PathTracer(ctx, a.width, a.height, a.spl, vertices=verts, mat_indices=mats, multigpu=multigpu)
        self.pt.sample_groups = a.sample_groups
        self.pt.ray_sort = a.ray_sort
        self.frame = self.pt.frame
        self.h2d_bytes = C.sizeof(self.pt.params)
        self.n_local = self.pt.num_samples if multigpu else a.width * a.height
        self.seg = {}
        self.rays_note = "radiance + shadow segments, counted once per subframe index by b200rt's instrumented launch (untimed)"

    def step(self, sub):
        self.pt.launch_subframe(sub)

    def rays(self, sub):
        return self.seg[sub]


class RaycastJob(Job):
    def __init__(self, a, ctx, host, scene):
        self.rc = host.Raycaster(ctx, scene)
        self.n = self.rc.buffer_rays(a.width)
        a.height = self.rc.height
        self.frame = (self.rc.hits, self.rc.hits_translated)
        self.h2d_bytes = 0
        self.rays_note = "buffered rays of both batches"

    def step(self, sub):
        self.rc.launch(want_ext=False)

    def rays(self, sub):
        return 2 * self.n

    def close(self):
        self.rc.close()


class WhittedJob(Job):
    def __init__(self, a, ctx, host, scene):
        self.mv = host.MeshViewer(ctx, scene, a.width, a.height)
        self.frame = self.mv.frame
        self.h2d_bytes = 128
        self.n = a.width * a.height
        self.rays_note = "launch indices (camera rays); shadow rays and BLEND continuation rays are not counted"

    def step(self, sub):
        self.mv.launch_subframe(sub)

    def rays(self, sub):
        return self.n

    def close(self):
        self.mv.close()


class PlaygroundJob(Job):
    def __init__(self, a, ctx, host):
        self.pg = host.Playground(ctx, a.width, a.height, spf=a.spl, rows=132)
        self.frame = self.pg.image
        self.h2d_bytes = 128
        self.n = a.width * a.height * a.spl
        self.rays_note = "camera samples (launch indices x samples per frame); light and bounce probes are not counted"

    def step(self, sub):
        self.pg.launch_frame(dirty=(sub == 0))

    def rays(self, sub):
        return self.n


def make_job(a, ctx, rank, world, host):
    if a.workload in PT_WORKLOADS:
        return PTJob(a, ctx, rank, world, host)
    if a.workload == "playground":
        return PlaygroundJob(a, ctx, host)
    from tests import common  # the Duck as a committed fixture (tests/golden/duck_mesh.npz): /root/reference does not exist on the GPU box
    scene = common.duck_scene()
    return RaycastJob(a, ctx, host, scene) if a.workload == "duck_raycast" else WhittedJob(a, ctx, host, scene)


def count_segments(a, job, subs, rank, world, host, L):
    """Segments of the path-tracing workloads per subframe index, by b200rt's instrumented launch.  Untimed.  The b200rt arm counts on
    its own path tracer; the reference arm borrows a b200rt path tracer over the same scene for the count and frees it again."""
    import torch
    if job.pt is None:
        return
    own = job.pt.ctx.lib is not None
    if own:
        pt = job.pt
    else:
        bctx = job.pt.ctx.helper
        pt = host.PathTracer(bctx, a.width, a.height, a.spl, vertices=job.pt.d_vertices if a.workload == "synthetic" else None,
                             mat_indices=job.pt.d_mat if a.workload == "synthetic" else None, multigpu=job.multigpu)
        pt.sample_groups = 1
    for sub in subs:
        pt.launch_subframe(sub, collect_stats=L.PT_STATS_SEGMENTS)
        job.seg[sub] = int(pt.stats.radiance_segments + pt.stats.shadow_segments)
    torch.cuda.synchronize()
    if not own:
        del pt
        torch.cuda.empty_cache()


def run_gpu(a, rank, world, local_rank, impl):
    import torch
    import torch.distributed as dist
    from optix_raytracer_b200 import host, _lib as L

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if impl == "reference":
        from oracle.optix_ref import backend as ob
        ctx = ob.OptixContext(local_rank)
    else:
        ctx = host.Context(local_rank)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- scene + acceleration structure (untimed setup; build time reported separately)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_setup = time.perf_counter()
    job = make_job(a, ctx, rank, world, host)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t_setup) * 1e3  # first build of the process: scene generation, allocations, build, compaction, SBT upload
    pt = job.pt
    is_pt = pt is not None
    accel_build_ms, info, scene_bytes = None, None, None
    if is_pt:
        # the metric's "BVH build ms": the accel-build call alone on preallocated buffers, second and third build of the process
        accel_build_ms = min(ctx.time_accel_build([pt.build_input], reps=2, warm=1)) if rank == 0 else None
        scene_bytes = int(pt.accel.buf.numel())
        if impl == "b200rt":
            info = pt.accel.info()
            scene_bytes = int(info.total_bytes)
    l2_bytes = torch.cuda.get_device_properties(local_rank).L2_cache_size
    flush = None
    if scene_bytes is None or scene_bytes < 2 * l2_bytes:
        flush = torch.empty(int(l2_bytes * 1.5) // 4, dtype=torch.float32, device=dev)

    exchange, shared = "none", None
    if world > 1 and is_pt and impl == "b200rt":
        exchange = a.exchange
        if exchange == "p2p":
            # one result buffer in rank 0's HBM, written by every rank's launch over NVLink; if any rank cannot map it (no peer path), all
            # ranks fall back to the gather together
            ok = 1
            try:
                shared = host.SharedResultBuffer(ctx, a.height, a.width, rank)
            except Exception as e:  # noqa: BLE001
                ok = 0
                print(f"[rank {rank}] shared result buffer unavailable ({e}); falling back to all-gather", file=sys.stderr)
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                if shared is not None:
                    shared.close()
                shared, exchange = None, "allgather"
            else:
                pt.params.result_buffer = shared.ptr
        if exchange == "allgather":
            gathered = torch.empty((world, job.n_local, 4), dtype=torch.float32, device=dev)
            full_accum = torch.empty((a.height, a.width, 4), dtype=torch.float32, device=dev)
            full_frame = torch.empty((a.height, a.width, 4), dtype=torch.uint8, device=dev)
    elif world > 1 and is_pt:
        # reference arm, N > 1: every rank runs the reference's programs on its own GPU over its share of the image (the reference's
        # one-thread device loop, optixMultiGPU.cpp:562-594, as one process per GPU) and keeps its pixels in a local result buffer
        exchange = "local result buffers (no gather)"
    islands = None
    if world > 1 and rank == 0 and impl == "b200rt":
        # optixNVLink's topology report (optixNVLink.cpp:1698-1825): which devices reach each other's memory over NVLink
        try:
            from optix_raytracer_b200 import topology
            islands = topology.format_islands(topology.compute_p2p_islands(topology.find_peers(world)))
            print(islands, file=sys.stderr)
        except Exception as e:  # noqa: BLE001
            islands = f"unavailable ({e})"

    frames = job.frame if isinstance(job.frame, tuple) else (job.frame,)
    h_frames = [torch.empty(f.shape, dtype=f.dtype).pin_memory() for f in frames]
    d2h_bytes = sum(f.numel() * f.element_size() for f in frames)

    def step(sub, want_host_frame=False):
        job.step(sub)
        out = frames
        if exchange == "allgather":
            dist.all_gather_into_tensor(gathered.view(-1), pt.accum.view(-1))
            ctx.check(ctx.lib.b200rt_deinterleave(ctx.h, ctx.stream, gathered.data_ptr(), world, job.n_local, a.width, a.height,
                                                  full_accum.data_ptr(), full_frame.data_ptr()), "deinterleave")
            out = (full_frame,)
        elif exchange == "p2p":
            out = (shared.tensor,)  # rank 0 only
            if want_host_frame:
                # the frame is complete when every rank's launch has finished: one barrier, then rank 0 reads it
                torch.cuda.synchronize()
                dist.barrier()
        if want_host_frame and (rank == 0 or exchange not in ("p2p",)):
            for h, f in zip(h_frames, out):
                h.copy_(f, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # subframe indices: warm-up 0 .. W-1, timed W .. W+K-1, e2e W+K .. W+2K-1 (the same in both arms)
    timed_subs = list(range(a.warmup, a.warmup + a.steps))
    e2e_subs = list(range(a.warmup + a.steps, a.warmup + 2 * a.steps))
    count_segments(a, job, timed_subs + e2e_subs, rank, world, host, L)
    for sub in range(a.warmup):
        step(sub)
    # ---- timed region: K steps, CUDA events around every step on the launching stream, no host synchronisation between steps
    # (L2 flushed between steps, outside the events, when the scene could be L2 resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    barrier()
    evs = []
    wall0 = time.perf_counter()
    for sub in timed_subs:
        if flush is not None:
            flush.fill_(1.0)
        s0, s1 = ev(), ev()
        s0.record()
        step(sub)
        s1.record()
        evs.append((s0, s1))
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    t_ms = sum(s0.elapsed_time(s1) for s0, s1 in evs)
    rays_total = sum(job.rays(sub) for sub in timed_subs)
    launches = ctx.kernel_launches - launches0 + (a.steps if exchange == "allgather" else 0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: same steps through the public call with host buffers (Params H2D from pinned memory is part of
    # launch_subframe; the frame comes back to pinned host memory every step)
    step(a.warmup + 2 * a.steps, want_host_frame=True)  # untimed: the first copy into the pinned frame buffer pays its one-time mapping cost
    barrier()
    evs = []
    for sub in e2e_subs:
        if flush is not None:
            flush.fill_(1.0)
        s0, s1 = ev(), ev()
        s0.record()
        step(sub, want_host_frame=True)
        s1.record()
        evs.append((s0, s1))
    barrier()
    e2e_ms = sum(s0.elapsed_time(s1) for s0, s1 in evs)
    if os.environ.get("B200RT_BENCH_DEBUG"):
        print("[bench debug] e2e per step (ms): " + " ".join(f"{s0.elapsed_time(s1):.3f}" for s0, s1 in evs), file=sys.stderr)
    e2e_rays = sum(job.rays(sub) for sub in e2e_subs)

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_ms, e2e_ms = allmax(t_ms), allmax(e2e_ms)
    rays_total, e2e_rays = allsum(float(rays_total)), allsum(float(e2e_rays))
    value = rays_total / (t_ms * 1e-3) / 1e6
    e2e_value = e2e_rays / (e2e_ms * 1e-3) / 1e6

    roofline = None
    if impl == "b200rt" and is_pt:
        roofline = pt_roofline(a, job, world, L, info, hbm_peak, peak_src, accel_build_ms)
    elif impl == "b200rt" and a.workload == "duck_raycast":
        roofline = raycast_roofline(job, ctx, t_ms / a.steps, hbm_peak, peak_src)

    out = None
    if rank == 0:
        scaling = "strong" if is_pt else "weak"
        par = (f"image split x{world} (StaticWorkDistribution 8x4 tiles), scene replicated" if is_pt else f"{world} independent replicas")
        par += {"none": "", "allgather": ", ncclAllGather of the sample buffers + de-interleave",
                "p2p": ", one result buffer in rank 0's HBM written by every rank's launch over NVLink (no collective)"}.get(exchange, ", " + exchange)
        cfg = {"workload": workload_name(a), "width": a.width, "height": a.height, "samples_per_launch": a.spl, "parallelism": par,
               "rays_counted": job.rays_note,
               "l2": ("inputs larger than L2: accel %.2f GB vs L2 %.0f MB" % (scene_bytes / 1e9, l2_bytes / 1e6)) if flush is None
                     else "L2 flushed between timed steps (1.5x L2 fill)",
               "scene_setup_cold_ms": build_ms, "segments_per_step": rays_total / a.steps}
        if is_pt:
            cfg.update({"triangles": int(info.num_triangles) if info else a.triangles, "accel_bytes": scene_bytes, "bvh_build_ms": accel_build_ms})
            if impl == "b200rt":
                cfg.update({"bvh8_nodes": int(info.num_nodes), "bvh8_node_bytes": int(info.reserved) or 80, "sample_groups": a.sample_groups,
                            "ray_sort": a.ray_sort})
        if islands:
            cfg["p2p_islands"] = islands
        out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": t_ms / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": cfg,
               "samples_per_sec_per_gpu": a.width * a.height * (a.spl or 1) * a.steps / (t_ms * 1e-3) / world,
               "wall_ms_per_step": wall_ms / a.steps,
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": job.h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                       "ms_per_step": e2e_ms / a.steps},
               "gpu_launches": int(launches), "clocks": clocks}
        if roofline is not None:
            out["roofline"] = roofline
        if impl == "reference":
            out.update({"impl": "reference", "reference_class": "optix",
                        "reference_note": "the reference's own device programs (oracle/_ref/*.ptx, compiled from /root/reference/SDK where they lie) on the "
                                          "driver's libnvoptix.so.1 through oracle/optix_ref: optixAccelBuild + optixLaunch with the same build inputs, Params "
                                          "bytes, SBT records and subframe indices as the b200rt arm; CUDA events around the same host calls",
                        "gpu_launches": a.steps * (2 if a.workload == "duck_raycast" else 1), "rtcore_version": int(ctx.olib.oref_rtcore_version())})
    keep = (ctx, job)  # keep alive until the CPU baseline is done
    return out, keep


def pt_roofline(a, job, world, L, info, hbm_peak, peak_src, accel_build_ms):
    """Roofline of the dominant kernel (trace stage), rank-local, two extra passes on one subframe: (1) CUDA events around every stage
    launch, (2) the instrumented kernel counting node / triangle fetches.  Plus the accel build against its streaming bytes."""
    pt = job.pt
    prof_sub = a.warmup + 2 * a.steps
    pt.launch_subframe(prof_sub, collect_stats=L.PT_STATS_SEGMENTS | L.PT_STATS_TIMING)
    trace_ms, shade_ms, iters = pt.stats.trace_ms, pt.stats.shade_ms, pt.stats.iterations
    rad, shd = pt.stats.radiance_segments, pt.stats.shadow_segments
    pt.launch_subframe(prof_sub, collect_stats=L.PT_STATS_SEGMENTS | L.PT_STATS_TRAVERSAL)
    nodes, tris = pt.stats.nodes_fetched, pt.stats.tris_tested
    assert (pt.stats.radiance_segments, pt.stats.shadow_segments) == (rad, shd)
    # algorithmic bytes of the trace stage: queue entry 4 + ray_d 16 + res 32 (RW) per active lane-iteration;
    # + ray_o 16 + hit 8 per radiance ray; + shd_o/shd_d/pend 48 per shadow ray; node_bytes per node and 48 B per triangle fetched
    lane_iters = rad  # every lane-iteration with an extension ray; the few shadow-only visits are counted via shd
    node_bytes = int(info.reserved) or 80  # 80: 8-bit boxes (large scenes), 224: fp32 boxes (cache-resident scenes)
    # a scene with fp32 node boxes (bvh_build.cu: at most 600 k triangles, < 100 MB) lives in L2: its node and triangle fetches never reach
    # HBM, and counting them against the HBM peak would give fractions above 1.  For such a scene the HBM bytes are the lane state only.
    cache_resident = node_bytes != 80
    scene_bytes = nodes * node_bytes + tris * 48
    algo_bytes = lane_iters * (4 + 16 + 32 + 16 + 8) + shd * 48 + (0 if cache_resident else scene_bytes)
    achieved = algo_bytes / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
    # measured DRAM bytes of the same launches (one ncu capture of this command, profiles/*_trace_dram.json), per launch like
    # `achieved`; null when no capture of this workload is committed
    traffic, traffic_src = None, None
    for f in sorted((ROOT / "profiles").glob("*_trace_dram.json"), reverse=True):
        try:
            cap = json.loads(f.read_text())
        except Exception:
            continue
        c = cap.get("config", {})
        if (c.get("workload"), c.get("triangles"), c.get("width"), c.get("height"), c.get("spl")) == (a.workload, a.triangles, a.width, a.height, a.spl) and world == 1:
            traffic = (cap["dram_read_bytes_per_step"] + cap["dram_write_bytes_per_step"]) / max(cap.get("trace_launches_per_step", iters), 1)
            traffic_src = (f"profiles/{f.name} (ncu dram__bytes_read.sum + dram__bytes_write.sum of the trace launches of one step of this command, "
                           f"captured with sample_groups {c.get('sample_groups')}; a committed capture, not re-measured in this run)")
            break
    roof = {"bound": "hbm", "kernel": "pt_trace_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes / max(iters, 1), "launches_per_step": iters,
            "avg_launch_ms": trace_ms / max(iters, 1), "trace_share_of_step": trace_ms / max(trace_ms + shade_ms, 1e-9),
            "nodes_per_segment": nodes / max(rad + shd, 1), "tris_per_segment": tris / max(rad + shd, 1),
            "visit_counts_source": "instrumented pt_trace_kernel on this repo's BVH8 (B200RT_PT_STATS_TRAVERSAL); the oracle-side counts on its own "
                                   "SAH tree are in profiles/ (tools/bvhlab)"}
    if cache_resident:
        roof["cache_resident"] = True
        roof["note"] = ("the scene lives in L2: `achieved` counts the lane state only (HBM traffic); the kernel is bound by instruction issue, not by HBM — "
                        f"node and triangle fetches served by the caches: {scene_bytes / max(iters, 1) / 1e6:.1f} MB per launch")
    if accel_build_ms:
        # B_build of SURVEY 8(d), for the variant built (bvh_build.cu), per triangle: gather 36 B read + 48 B written, Morton keys 48 + 12,
        # `passes` radix passes of 8 B histogram read + 12 B read + 12 B written, the radix tree with its boxes in one bottom-up pass (leaf:
        # 52 read + 32 written; internal node: sibling box 32 read, own box 32 written, 24 B of exchange words, 16 B of keys), triangle
        # records 8 + 48 read + 48 written; per wide node: the plan looks at ~14 boxes of 32 B and writes 40 B, the emit reads them (8 x 32 +
        # 40) and writes the node
        T = int(info.num_triangles)
        bits = 10 if T < (1 << 14) else (16 if T < (1 << 27) else 21)
        passes = (3 * bits + 7) // 8
        b_build = T * (36 + 48 + 48 + 12 + passes * 32 + 52 + 32 + 104 + 8 + 48 + 48) + int(info.num_nodes) * (14 * 32 + 40 + 8 * 32 + 40 + node_bytes)
        roof["build"] = {"bound": "hbm", "kernel": "b200rt_accel_build (all kernels of one build)", "algorithmic_bytes": b_build, "ms": accel_build_ms,
                         "achieved": b_build / (accel_build_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": b_build / (accel_build_ms * 1e-3) / 1e9 / hbm_peak, "radix_passes": passes}
    return roof


def raycast_roofline(job, ctx, ms_per_step, hbm_peak, peak_src):
    """optixRaycasting launch = one one-ray-per-thread kernel per batch: 32 B ray in, 16 B Hit out, node and triangle fetches as counted by
    b200rt_trace_stats on the same ray buffers.  The Duck's BVH (Node8F, < 1 MB) is cache resident, so this says how far the launch is
    from streaming its rays, not that HBM is what bounds it."""
    n1, t1 = ctx.trace_stats(job.rc.ias, job.rc.rays)
    n2, t2 = ctx.trace_stats(job.rc.ias, job.rc.rays_translated)
    algo = 2 * job.n * 48                              # what has to stream through HBM: the rays in, the Hit records out
    cached = (n1 + n2) * 224 + (t1 + t2) * 48          # node / triangle fetches: served by L1 / L2, never by HBM
    ach = algo / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "raycast_simple_kernel", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": algo / 2, "launches_per_step": 2, "cache_resident": True,
            "note": f"the scene lives in L2: `achieved` counts rays in + hits out only; node and triangle fetches served by the caches: {cached / 2 / 1e6:.1f} MB "
                    "per launch; the kernel is bound by instruction issue, not by HBM",
            "nodes_per_segment": (n1 + n2) / (2 * job.n), "tris_per_segment": (t1 + t2) / (2 * job.n)}


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    impl = a.impl
    if impl == "reference":
        why = "--reference-class cpu"
        use_optix = False
        if a.reference_class != "cpu":
            try:
                from oracle.optix_ref import backend as ob
                use_optix, why = ob.available(local_rank)
            except Exception as e:  # noqa: BLE001
                use_optix, why = False, f"{type(e).__name__}: {e}"
            if not use_optix and a.reference_class == "optix":
                if rank == 0:
                    print(json.dumps({"impl": "reference", "unavailable": f"OptiX: {why}"}), flush=True)
                return 0
        if not use_optix:
            out = run_reference_cpu(a, rank, why)
            if out is not None:
                print(json.dumps(out), flush=True)
            return 0
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out, keep = run_gpu(a, rank, world, local_rank, impl)
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = run_cpu_baseline(a, a.cpu_seconds)
        print(json.dumps(out), flush=True)
    if hasattr(keep[1], "close"):
        keep[1].close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
