"""Launches are asynchronous (SURVEY.md 8(b) "Threading": optixLaunch returns at once and the caller synchronises; the reference's
one-thread loop over devices, SDK/optixMultiGPU/optixMultiGPU.cpp:562-594, and the two streams of optixRaycasting.cpp:291-313 rely on
it).  These tests check that the b200rt launches return before their work is done, that a queue of launches gives the image of
synchronous launches bit for bit, that launches on two streams of one context do not corrupt each other's lane state, and — on a box
with two GPUs — that one host thread keeps two devices busy at the same time and that the N-GPU frame equals the 1-GPU frame."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _device_ms(fn, dev=0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


@pytest.mark.parametrize("multigpu", [None, (0, 1)])
def test_pathtracer_launch_returns_before_the_work_is_done(multigpu):
    from optix_raytracer_b200 import host
    ctx = host.Context(0)
    pt = host.PathTracer(ctx, 768, 768, 16, multigpu=multigpu)
    pt.sample_groups = 4
    for sub in range(3):   # warm-up: workspace allocation, graph instantiation
        pt.launch_subframe(sub)
    dev_ms = _device_ms(lambda: pt.launch_subframe(3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pt.launch_subframe(4)
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    assert dev_ms > 1.0
    assert host_ms < 0.5 * dev_ms, f"launch call took {host_ms:.3f} ms on the host, the work {dev_ms:.3f} ms on the device: the call waited"
    # the asynchronous loop (CUDA graph, conditional WHILE) computes what the instrumented, synchronising launch computes
    ref = host.PathTracer(ctx, 768, 768, 16, multigpu=multigpu)
    ref.sample_groups = 4
    for sub in range(5):
        ref.launch_subframe(sub, collect_stats=1)
    torch.cuda.synchronize()
    assert torch.equal(pt.accum.view(torch.int32), ref.accum.view(torch.int32))
    ctx.close()


def test_queued_launches_on_two_streams_of_one_context_keep_their_state_apart():
    from optix_raytracer_b200 import host
    ctx = host.Context(0)
    a = host.PathTracer(ctx, 256, 256, 8)
    b = host.PathTracer(ctx, 320, 200, 4, multigpu=(0, 1))
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for sub in range(4):   # interleaved, never synchronised in between: both use the context's one workspace
        with torch.cuda.stream(sa):
            a.launch_subframe(sub)
        with torch.cuda.stream(sb):
            b.launch_subframe(sub)
    torch.cuda.synchronize()
    ra = host.PathTracer(ctx, 256, 256, 8)
    rb = host.PathTracer(ctx, 320, 200, 4, multigpu=(0, 1))
    for sub in range(4):
        ra.launch_subframe(sub); torch.cuda.synchronize()
        rb.launch_subframe(sub); torch.cuda.synchronize()
    assert torch.equal(a.accum.view(torch.int32), ra.accum.view(torch.int32))
    assert torch.equal(b.accum.view(torch.int32), rb.accum.view(torch.int32))
    ctx.close()


def test_raycast_handle_swap_in_the_same_params_block():
    """the any-hit decision must follow the handle that is in Params NOW, not the one seen at the first launch."""
    from optix_raytracer_b200 import host
    from tests import common
    ctx = host.Context(0)
    opaque, masked = host.Raycaster(ctx, common.duck_scene()), host.Raycaster(ctx, common.duck_alpha_scene(1))
    opaque.buffer_rays(400); masked.buffer_rays(400)
    opaque.launch(); masked.launch()
    torch.cuda.synchronize()
    t_opaque, t_masked = opaque.hits.clone(), masked.hits.clone()
    assert not torch.equal(t_opaque, t_masked)  # the holes of the alpha mask let rays through
    # same Params block and the SAME SBT (the masked scene's records), first with the masked scene's traversable, then with the opaque
    # scene's: geometry flags live in the traversable, so the second launch must not run any-hit programs
    p = host.RaycastParams(opaque.ias.handle, masked.rays.data_ptr(), masked.hits.data_ptr())
    masked.d_params.copy_(ctx.to_device(np.frombuffer(bytes(p), np.uint8).copy()))
    ctx.launch_raycast(masked.programs, masked.d_params.data_ptr(), masked.sbt, masked.width, masked.height, 0)
    torch.cuda.synchronize()
    assert torch.equal(masked.hits, t_opaque)
    p = host.RaycastParams(masked.ias.handle, masked.rays.data_ptr(), masked.hits.data_ptr())
    masked.d_params.copy_(ctx.to_device(np.frombuffer(bytes(p), np.uint8).copy()))
    ctx.launch_raycast(masked.programs, masked.d_params.data_ptr(), masked.sbt, masked.width, masked.height, 0)
    torch.cuda.synchronize()
    assert torch.equal(masked.hits, t_masked)
    opaque.close(); masked.close(); ctx.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_thread_keeps_two_devices_busy_and_the_split_frame_equals_the_whole():
    from optix_raytracer_b200 import host
    W, H, SPL = 1024, 768, 16
    ctxs = [host.Context(d) for d in range(2)]
    pts = []
    for d, c in enumerate(ctxs):
        with torch.cuda.device(d):
            pts.append(host.PathTracer(c, W, H, SPL, multigpu=(d, 2)))
    for p in pts:
        p.sample_groups = 4
    def both(sub):
        for d, p in enumerate(pts):
            with torch.cuda.device(d):
                p.launch_subframe(sub)
    def sync():
        for d in range(2):
            torch.cuda.synchronize(d)
    for sub in range(2):
        both(sub)
    sync()
    # each device alone, then both from this one thread without synchronising in between (optixMultiGPU.cpp:562-594)
    alone = []
    for d, p in enumerate(pts):
        with torch.cuda.device(d):
            t0 = time.perf_counter(); p.launch_subframe(2); torch.cuda.synchronize(d); alone.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); both(3); sync(); together = time.perf_counter() - t0
    assert together < 0.75 * sum(alone), f"two devices took {together * 1e3:.2f} ms together, {alone[0] * 1e3:.2f} + {alone[1] * 1e3:.2f} ms alone: serialised"
    # N-GPU frame == 1-GPU frame (seeds depend on the pixel only): scatter the two compact sample buffers and compare with one device's
    one = host.PathTracer(ctxs[0], W, H, SPL, multigpu=(0, 1))
    one.sample_groups = 4
    for sub in range(4):
        one.launch_subframe(sub)
    sync()
    full = one.accum.cpu().numpy()   # N = 1: sample s is pixel wd_sample_pixel(s)
    img1 = np.zeros((H, W, 4), np.float32)
    n1 = one.num_samples
    xy = np.array([host.wd_sample_pixel(W, H, 1, 0, s) for s in range(0, n1, 997)])
    for k, s in enumerate(range(0, n1, 997)):
        img1[xy[k, 1], xy[k, 0]] = full[s]
    for d, p in enumerate(pts):
        part = p.accum.cpu().numpy()
        for s in range(0, p.num_samples, 997):
            x, y = host.wd_sample_pixel(W, H, 2, d, s)
            if x < W and y < H:
                s1 = None
                # find the N = 1 sample index of this pixel: tiles of 8x4, strips of 8 columns
                col, row = x // 8, y // 4
                s1 = (row * ((W + 7) // 8) + col) * 32 + (y % 4) * 8 + (x % 8)
                assert np.array_equal(part[s].view(np.uint32), full[s1].view(np.uint32)), (d, s, x, y)
    for c in ctxs:
        c.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_textures_shared_across_a_p2p_island_render_the_same_frames():
    """optixNVLink's texture sharing (loadTextures, SDK/optixNVLink/optixNVLink.cpp:1501-1590): one process drives two
    devices of one P2P island, the Duck's base-colour texture lives on ONE of them (the one with the least texture memory) and the other
    samples it over NVLink through its own texture object (b200rt_texture_view).  The viewer's frames on the device WITHOUT a copy equal,
    bit for bit, the frames it renders from a private copy; a second texture goes to the other device (least usage)."""
    from optix_raytracer_b200 import host, topology
    from tests import common
    if not (torch.cuda.can_device_access_peer(0, 1) and torch.cuda.can_device_access_peer(1, 0)):
        pytest.skip("devices 0 and 1 have no peer access")
    ctxs = [host.Context(d) for d in range(2)]
    ctxs[0].enable_peer_access(1)
    ctxs[1].enable_peer_access(0)
    islands = topology.compute_p2p_islands(topology.peers_from_cuda(2, torch.cuda.can_device_access_peer))
    assert islands == [0b11]
    sc = common.duck_scene()
    shared, usage = host.create_scene_textures_shared(ctxs, sc, islands)
    assert usage[0] == float(sc["images"][0].size) and usage[1] == 0.0, usage     # one copy in the island, on device 0
    assert shared[1][1][0][2] == 0 and shared[0][1][0][2] != 0                      # device 1 holds a view (no array of its own)
    W, H = 960, 540
    with torch.cuda.device(1):
        mv_shared = host.MeshViewer(ctxs[1], sc, W, H, textures=shared[1])
        mv_own = host.MeshViewer(ctxs[1], sc, W, H)
        mv_plain = host.MeshViewer(ctxs[1], common.duck_scene(textured=False), W, H)
        for sub in range(3):
            for mv in (mv_shared, mv_own, mv_plain):
                mv.launch_subframe(sub)
        torch.cuda.synchronize(1)
        assert torch.equal(mv_shared.frame, mv_own.frame) and torch.equal(mv_shared.accum.view(torch.int32), mv_own.accum.view(torch.int32))
        assert not torch.equal(mv_shared.frame, mv_plain.frame), "the texture does not show in the frame: the test would not see a broken view"
    with torch.cuda.device(0):
        mv0 = host.MeshViewer(ctxs[0], sc, W, H, textures=shared[0])
        for sub in range(3):
            mv0.launch_subframe(sub)
        torch.cuda.synchronize(0)
        assert torch.equal(mv0.frame.cpu(), mv_own.frame.cpu())
    # a scene with two textures: the second copy goes to the device that holds less
    sc2 = common.duck_scene()
    sc2["images"] = [sc2["images"][0], sc2["images"][0][::-1].copy()]
    sc2["textures"] = [{"sampler": 0, "source": 0}, {"sampler": 0, "source": 1}]
    shared2, usage2 = host.create_scene_textures_shared(ctxs, sc2, islands)
    assert usage2[0] == usage2[1] > 0
    for per_ctx in (shared[0], shared[1], shared2[0], shared2[1]):
        for c, tex, arr in per_ctx[1]:
            c.lib.b200rt_texture_destroy(c.h, tex, arr)
    for mv in (mv_own, mv_plain):
        mv.close()
    for c in ctxs:
        c.close()


def test_accel_build_and_compact_do_not_wait_for_the_device():
    """optixAccelBuild / optixAccelCompact are asynchronous: the level and clustering loops run on the device (CUDA graphs)."""
    from optix_raytracer_b200 import host, _lib as L
    import ctypes as C
    ctx = host.Context(0)
    verts, mats = host.synthetic_mesh(ctx, 4_000_000, 0)
    bi = ctx.triangle_input(verts, sbt_index=mats, num_sbt=4, vertex_stride=16)
    arr = (L.BuildInput * 1)(bi)
    opts = L.AccelBuildOptions(L.BUILD_FLAG_ALLOW_COMPACTION, L.BUILD_OPERATION_BUILD)
    sizes = L.AccelBufferSizes()
    ctx._accel_memory_usage(opts, arr, 1, sizes)
    temp, out = ctx.empty_bytes(sizes.tempSizeInBytes), ctx.empty_bytes(sizes.outputSizeInBytes)
    out2 = ctx.empty_bytes(sizes.outputSizeInBytes)
    csize = torch.zeros(1, dtype=torch.int64, device=ctx.torch_device)
    emit = L.AccelEmitDesc(csize.data_ptr(), L.PROPERTY_TYPE_COMPACTED_SIZE)
    def build():
        h, h2 = C.c_uint64(), C.c_uint64()
        ctx._accel_build(opts, arr, 1, temp.data_ptr(), sizes.tempSizeInBytes, out.data_ptr(), sizes.outputSizeInBytes, h, emit)
        ctx._accel_compact(h.value, out2.data_ptr(), sizes.outputSizeInBytes, h2)
        return h2.value
    build(); torch.cuda.synchronize()
    dev_ms = _device_ms(build)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); handle = build(); host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    assert host_ms < 0.6 * dev_ms, f"build + compact calls took {host_ms:.3f} ms on the host, the work {dev_ms:.3f} ms on the device"
    acc = host.Accel(ctx, out2, handle)
    info = acc.info()
    assert info.num_triangles == 4_000_000 and int(csize.item()) == (info.total_bytes + 127) // 128 * 128
    ctx.close()
