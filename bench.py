#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native path tracer (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W                 # b200rt arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W # CPU arm: the oracle port on the host cores

Workload (default): BASELINE.json configs[4] — procedurally tessellated 50 M-triangle scene, 3840x2160,
samples_per_launch 16 (64 spp = 4 steps), optixMultiGPU device programs (depth cap 3), scene replicated per
GPU, image split with the reference's StaticWorkDistribution, per-rank sample buffers all-gathered over NCCL
and de-interleaved.  A "step" is one launch (one subframe) over the whole image by all ranks together.
Metric: Mrays/s = traced segments (radiance + shadow) of all ranks / max-over-ranks device time.
Prints ONE JSON line on rank 0.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200rt", choices=["b200rt", "reference"])
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "cornell"])
    ap.add_argument("--triangles", type=int, default=50_000_000)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spl", type=int, default=16)
    ap.add_argument("--sample-groups", type=int, default=None, help="b200rt_pt_options.sample_groups: lanes per launch index (1 = reference summation order)")
    ap.add_argument("--ray-sort", type=int, default=None, help="b200rt_pt_options.ray_sort (default 0: measured slower end to end, profiles/r01_trace_kernel.md)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "allgather"],
                    help="N > 1: how the frame is assembled — p2p: every rank's launch stores its pixels into rank 0's result buffer over NVLink "
                         "(the reference's single result buffer, optixMultiGPU.cpp:479-508); allgather: per-rank sample buffers all-gathered (NCCL) "
                         "and de-interleaved on every rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(a):
    if a.workload == "synthetic":
        return (f"synthetic tessellated mesh, {a.triangles} triangles, {a.width}x{a.height}, samples_per_launch {a.spl}, "
                "optixMultiGPU programs (BASELINE.json configs[4])")
    return f"optixPathTracer Cornell box, {a.width}x{a.height}, samples_per_launch {a.spl} (BASELINE.json configs[0])"


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
# CPU arm / cpu_baseline: the scalar oracle on the host cores
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
    """Bounded sample of the same workload on the host cores (kind 'port': the oracle restatement)."""
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


def run_reference(a, rank):
    """--impl reference: the reference has no CPU renderer and its traversal lives in the closed libnvoptix, so the
    CPU arm is the oracle port (oracle/oracle.cpp) with all host threads; each step is a bounded strip of the image."""
    if rank != 0:
        return None
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
    return {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": spent / max(a.steps, 1) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": orc.ncores(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


# ---------------------------------------------------------------------------------------------------
# b200rt arm
# ---------------------------------------------------------------------------------------------------
def run_b200rt(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from optix_raytracer_b200 import host, _lib as L

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = host.Context(local_rank)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- scene + acceleration structure (untimed setup; build time reported separately)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    if a.workload == "synthetic":
        verts, mats = host.synthetic_mesh(ctx, a.triangles, 0)
        multigpu = (rank, world)
    else:
        verts = mats = None
        multigpu = (rank, world) if world > 1 else None
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    pt = host.PathTracer(ctx, a.width, a.height, a.spl, vertices=verts, mat_indices=mats, multigpu=multigpu)
    pt.sample_groups = a.sample_groups
    pt.ray_sort = a.ray_sort
    e1.record()
    torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)  # first build of the process: allocations, build, compaction, SBT upload
    # the metric's "BVH build ms": the accel-build call alone on preallocated buffers, second and third build of the process
    accel_build_ms = min(ctx.time_accel_build([pt.build_input], reps=2, warm=1)) if rank == 0 else None
    info = pt.accel.info()
    scene_bytes = int(info.total_bytes)
    l2_bytes = torch.cuda.get_device_properties(local_rank).L2_cache_size
    flush = None
    if scene_bytes < 2 * l2_bytes:
        flush = torch.empty(int(l2_bytes * 1.5) // 4, dtype=torch.float32, device=dev)

    n_local = pt.num_samples if multigpu else a.width * a.height
    exchange, shared = "none", None
    if world > 1:
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
            gathered = torch.empty((world, n_local, 4), dtype=torch.float32, device=dev)
            full_accum = torch.empty((a.height, a.width, 4), dtype=torch.float32, device=dev)
            full_frame = torch.empty((a.height, a.width, 4), dtype=torch.uint8, device=dev)
    islands = None
    if world > 1 and rank == 0:
        # optixNVLink's topology report (optixNVLink.cpp:1698-1825): which devices reach each other's memory over NVLink
        try:
            from optix_raytracer_b200 import topology
            islands = topology.format_islands(topology.compute_p2p_islands(topology.find_peers(world)))
            print(islands, file=sys.stderr)
        except Exception as e:  # noqa: BLE001
            islands = f"unavailable ({e})"
    h_frame = torch.empty((a.height, a.width, 4), dtype=torch.uint8).pin_memory()
    stats_mask = L.PT_STATS_SEGMENTS

    def step(sub, want_host_frame=False, mask=stats_mask):
        pt.launch_subframe(sub, collect_stats=mask)
        segs = pt.stats.radiance_segments + pt.stats.shadow_segments
        frame = pt.frame
        if exchange == "allgather":
            dist.all_gather_into_tensor(gathered.view(-1), pt.accum.view(-1))
            ctx.check(ctx.lib.b200rt_deinterleave(ctx.h, ctx.stream, gathered.data_ptr(), world, n_local, a.width, a.height,
                                                  full_accum.data_ptr(), full_frame.data_ptr()), "deinterleave")
            frame = full_frame
        elif exchange == "p2p":
            frame = shared.tensor  # rank 0 only
            if want_host_frame:
                # the frame is complete when every rank's launch has finished: one barrier, then rank 0 reads it
                torch.cuda.synchronize()
                dist.barrier()
        if want_host_frame and rank == 0:
            h_frame.copy_(frame, non_blocking=True)
        return segs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sub = 0
    for _ in range(a.warmup):
        step(sub); sub += 1
    # ---- timed region: K steps, CUDA events per step (L2 flushed between steps when the scene could be L2 resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    barrier()
    t_ms, segs_total = 0.0, 0
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        if flush is not None:
            flush.fill_(1.0)
        s0, s1 = ev(), ev()
        s0.record()
        segs_total += step(sub); sub += 1
        s1.record()
        s1.synchronize()
        t_ms += s0.elapsed_time(s1)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    launches = ctx.kernel_launches - launches0 + (a.steps if exchange == "allgather" else 0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: same steps through the public call with host buffers (Params H2D from pinned memory is part of
    # launch_subframe; the frame comes back to pinned host memory every step)
    barrier()
    e2e_ms, e2e_segs = 0.0, 0
    for _ in range(a.steps):
        if flush is not None:
            flush.fill_(1.0)
        s0, s1 = ev(), ev()
        s0.record()
        e2e_segs += step(sub, want_host_frame=True); sub += 1
        s1.record()
        s1.synchronize()
        e2e_ms += s0.elapsed_time(s1)
    barrier()

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
    segs_total, e2e_segs = allsum(float(segs_total)), allsum(float(e2e_segs))
    value = segs_total / (t_ms * 1e-3) / 1e6
    e2e_value = e2e_segs / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (trace stage), rank-local, two extra passes on the same subframe:
    # (1) CUDA events around every stage launch, (2) the instrumented kernel counting node / triangle fetches
    prof_sub = sub
    pt.launch_subframe(prof_sub, collect_stats=L.PT_STATS_SEGMENTS | L.PT_STATS_TIMING)
    trace_ms, shade_ms, iters = pt.stats.trace_ms, pt.stats.shade_ms, pt.stats.iterations
    rad, shd = pt.stats.radiance_segments, pt.stats.shadow_segments
    pt.launch_subframe(prof_sub, collect_stats=L.PT_STATS_SEGMENTS | L.PT_STATS_TRAVERSAL)
    nodes, tris = pt.stats.nodes_fetched, pt.stats.tris_tested
    assert (pt.stats.radiance_segments, pt.stats.shadow_segments) == (rad, shd)
    # algorithmic bytes of the trace stage: queue entry 4 + ray_d 16 + res 32 (RW) per active lane-iteration;
    # + ray_o 16 + hit 8 per radiance ray; + shd_o/shd_d/pend 48 per shadow ray; 80 B per node and 48 B per triangle fetched
    lane_iters = rad  # every lane-iteration with an extension ray; the few shadow-only visits are counted via shd
    node_bytes = int(info.reserved) or 80  # 80: 8-bit boxes (large scenes), 224: fp32 boxes (cache-resident scenes)
    algo_bytes = lane_iters * (4 + 16 + 32 + 16 + 8) + shd * 48 + nodes * node_bytes + tris * 48
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
            traffic = (cap["dram_read_bytes_per_step"] + cap["dram_write_bytes_per_step"]) / max(iters, 1)
            traffic_src = f"profiles/{f.name} (ncu dram__bytes_read.sum + dram__bytes_write.sum of the trace launches of one step, captured with sample_groups {c.get('sample_groups')})"
            break
    roofline = {"bound": "hbm", "kernel": "pt_trace_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes / max(iters, 1), "launches_per_step": iters,
                "avg_launch_ms": trace_ms / max(iters, 1), "trace_share_of_step": trace_ms / max(trace_ms + shade_ms, 1e-9),
                "nodes_per_segment": nodes / max(rad + shd, 1), "tris_per_segment": tris / max(rad + shd, 1)}

    out = None
    if rank == 0:
        out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": t_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": workload_name(a), "triangles": int(info.num_triangles), "bvh8_nodes": int(info.num_nodes), "bvh8_node_bytes": int(info.reserved) or 80,
                          "accel_bytes": scene_bytes, "width": a.width, "height": a.height, "samples_per_launch": a.spl, "sample_groups": a.sample_groups, "ray_sort": a.ray_sort,
                          "parallelism": f"image split x{world} (StaticWorkDistribution 8x4 tiles), scene replicated"
                                         + {"none": "", "allgather": ", ncclAllGather of the sample buffers + de-interleave",
                                            "p2p": ", one result buffer in rank 0's HBM written by every rank's launch over NVLink (no collective)"}[exchange],
                          "l2": ("inputs larger than L2: accel %.2f GB vs L2 %.0f MB" % (scene_bytes / 1e9, l2_bytes / 1e6)) if flush is None
                                else "L2 flushed between timed steps (1.5x L2 fill)",
                          "bvh_build_ms": accel_build_ms, "scene_setup_cold_ms": build_ms, **({"p2p_islands": islands} if islands else {}), "segments_per_step": segs_total / a.steps},
               "samples_per_sec_per_gpu": a.width * a.height * a.spl * a.steps / (t_ms * 1e-3) / world,
               "wall_ms_per_step": wall_ms / a.steps,
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(pt.params), "d2h_bytes_per_step": a.width * a.height * 4,
                       "ms_per_step": e2e_ms / a.steps},
               "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}
    ctx_keep = (ctx, pt)  # keep alive until the CPU baseline is done
    return out, ctx_keep


def main():
    a = parse()
    if a.width is None:
        a.width, a.height = (3840, 2160) if a.workload == "synthetic" else (768, 768)
    if a.ray_sort is None:
        a.ray_sort = 0
    if a.sample_groups is None:
        a.sample_groups = 8 if a.workload == "synthetic" else 4   # measured best on one B200 (gpurun_out/sg_*.json); any value gives the same image on any N
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        out = run_reference(a, rank)
        if out is not None:
            print(json.dumps(out), flush=True)
        return 0
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out, keep = run_b200rt(a, rank, world, local_rank)
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = run_cpu_baseline(a, a.cpu_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
