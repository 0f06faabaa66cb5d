"""Host-side mirror of the reference samples' launch plumbing, on top of the C ABI (include/b200rt.h).

This is the Python test/bench harness equivalent of what the reference's C++ `main()`s do around
`optixAccelBuild` / `optixLaunch` (SDK/optixPathTracer/optixPathTracer.cpp:424-511,576-898,
SDK/optixRaycasting/optixRaycasting.cpp:94-349, SDK/optixMultiGPU/optixMultiGPU.cpp:479-594): it owns
device buffers (torch tensors — plumbing only), fills the reference's Params / SBT record layouts
byte for byte and calls the library.  All ray tracing happens in libb200rt.so.
"""
import ctypes as C
import json
import math
import pathlib
import struct

import numpy as np
import torch

from . import _lib as L

DATA_DIR = pathlib.Path(__file__).resolve().parent / "data"


class B200RTError(RuntimeError):
    pass


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Context:
    """b200rt_context (replaces OptixDeviceContext; SDK/optixPathTracer/optixPathTracer.cpp:555-573)."""

    def __init__(self, device=0, log_level=0):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise B200RTError("no CUDA device: b200rt has no CPU path")
        self.device = device
        self.torch_device = torch.device("cuda", device)
        self._messages = []

        def _cb(level, tag, msg, _):
            self._messages.append((level, tag.decode(), msg.decode()))
            if log_level >= 4:
                print(f"[{level:2d}][{tag.decode():>12s}]: {msg.decode()}")

        self._cb = L.LOG_CB(_cb)
        h = C.c_void_p()
        rc = self.lib.b200rt_context_create(device, self._cb, None, log_level, C.byref(h))
        if rc:
            raise B200RTError(f"b200rt_context_create: {self.lib.b200rt_error_name(rc).decode()}")
        self.h = h

    def check(self, rc, what=""):
        if rc:
            raise B200RTError(f"{what}: {self.lib.b200rt_error_name(rc).decode()} "
                              f"({self.lib.b200rt_last_error_message(self.h).decode()})")

    def enable_peer_access(self, peer_device):
        """enablePeerAccess (SDK/optixNVLink/optixNVLink.cpp:1852-1866): this context's device may read and write `peer_device`'s memory."""
        self.check(self.lib.b200rt_enable_peer_access(self.h, int(peer_device)), "enable_peer_access")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rt_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    @property
    def kernel_launches(self):
        return int(self.lib.b200rt_context_kernel_launches(self.h))

    def to_device(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.torch_device)

    def empty_bytes(self, n, align=128):
        # torch's caching allocator returns >= 512-byte aligned blocks
        t = torch.empty(max(int(n), 1), dtype=torch.uint8, device=self.torch_device)
        assert t.data_ptr() % align == 0
        return t

    # ---- acceleration structures ---------------------------------------------------------------
    # the three acceleration-structure calls of the boundary; oracle/optix_ref/backend.py overrides them with OptiX's
    def _accel_memory_usage(self, opts, arr, n, sizes):
        self.check(self.lib.b200rt_accel_compute_memory_usage(self.h, C.byref(opts), arr, n, C.byref(sizes)), "compute_memory_usage")

    def _accel_build(self, opts, arr, n, temp, temp_bytes, out, out_bytes, handle, emit):
        self.check(self.lib.b200rt_accel_build(self.h, self.stream, C.byref(opts), arr, n, temp, temp_bytes, out, out_bytes, C.byref(handle),
                                               C.byref(emit) if emit is not None else None, 1 if emit is not None else 0), "accel_build")

    def _accel_compact(self, handle, out, out_bytes, new_handle):
        self.check(self.lib.b200rt_accel_compact(self.h, self.stream, handle, out, out_bytes, C.byref(new_handle)), "accel_compact")

    def build_accel(self, build_inputs, compact=True, keep=()):
        """optixAccelComputeMemoryUsage + optixAccelBuild (+ optixAccelCompact), the sequence of
        SDK/optixPathTracer/optixPathTracer.cpp:627-684.  Returns Accel."""
        n = len(build_inputs)
        arr = (L.BuildInput * n)(*build_inputs)
        opts = L.AccelBuildOptions(L.BUILD_FLAG_ALLOW_COMPACTION if compact else 0, L.BUILD_OPERATION_BUILD)
        sizes = L.AccelBufferSizes()
        self._accel_memory_usage(opts, arr, n, sizes)
        temp = self.empty_bytes(sizes.tempSizeInBytes)
        out = self.empty_bytes(sizes.outputSizeInBytes)
        csize = torch.zeros(1, dtype=torch.int64, device=self.torch_device)
        # the compacted size may only be queried when the build allows compaction (OptiX rejects it otherwise)
        emit = L.AccelEmitDesc(csize.data_ptr(), L.PROPERTY_TYPE_COMPACTED_SIZE) if compact else None
        handle = C.c_uint64()
        self._accel_build(opts, arr, n, temp.data_ptr(), sizes.tempSizeInBytes, out.data_ptr(), sizes.outputSizeInBytes, handle, emit)
        acc = Accel(self, out, handle.value, keep)
        acc.uncompacted_bytes = int(sizes.outputSizeInBytes)
        acc.temp_bytes = int(sizes.tempSizeInBytes)
        if compact:
            compacted = int(csize.item())
            if compacted < sizes.outputSizeInBytes:
                out2 = self.empty_bytes(compacted)
                h2 = C.c_uint64()
                self._accel_compact(handle.value, out2.data_ptr(), compacted, h2)
                torch.cuda.synchronize(self.torch_device)
                acc = Accel(self, out2, h2.value, keep)
                acc.uncompacted_bytes = int(sizes.outputSizeInBytes)
                acc.temp_bytes = int(sizes.tempSizeInBytes)
        del temp
        return acc

    def time_accel_build(self, build_inputs, reps=2, warm=1):
        """The "BVH build ms" figure: optixAccelBuild alone (vertices on the device -> traversable ready), CUDA events on the build's
        stream around the call, temp and output buffers allocated beforehand as the reference does (optixPathTracer.cpp:634-662).
        Returns the per-build times of the `reps` builds after `warm` untimed ones."""
        n = len(build_inputs)
        arr = (L.BuildInput * n)(*build_inputs)
        opts = L.AccelBuildOptions(L.BUILD_FLAG_ALLOW_COMPACTION, L.BUILD_OPERATION_BUILD)
        sizes = L.AccelBufferSizes()
        self._accel_memory_usage(opts, arr, n, sizes)
        temp = self.empty_bytes(sizes.tempSizeInBytes)
        out = self.empty_bytes(sizes.outputSizeInBytes)
        csize = torch.zeros(1, dtype=torch.int64, device=self.torch_device)
        emit = L.AccelEmitDesc(csize.data_ptr(), L.PROPERTY_TYPE_COMPACTED_SIZE)
        times = []
        for k in range(warm + reps):
            handle = C.c_uint64()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._accel_build(opts, arr, n, temp.data_ptr(), sizes.tempSizeInBytes, out.data_ptr(), sizes.outputSizeInBytes, handle, emit)
            e1.record()
            torch.cuda.synchronize(self.torch_device)
            if k >= warm:
                times.append(e0.elapsed_time(e1))
        del temp, out
        return times

    # ---- launches (the optixLaunch replacements) and SBT headers --------------------------------------
    def prepare_programs(self, kind):
        """kind: 'pathtracer' | 'multigpu' | 'raycast'.  Nothing to do here: the device programs are compiled into
        libb200rt.so.  (The OptiX backend creates module / program groups / pipeline at this point.)"""
        return None

    def sbt_header(self, programs, kind, index):
        """The 32-byte SBT record header optixSbtRecordPackHeader would write; b200rt ignores it (zeros)."""
        return bytes(32)

    def launch_pathtracer(self, programs, d_params, params_size, sbt, width, height, opts):
        self.check(self.lib.b200rt_launch_pathtracer(self.h, self.stream, d_params, C.byref(sbt), width, height, C.byref(opts)), "launch_pathtracer")

    def launch_multigpu(self, programs, d_params, params_size, sbt, num_samples, opts):
        self.check(self.lib.b200rt_launch_multigpu(self.h, self.stream, d_params, C.byref(sbt), num_samples, C.byref(opts)), "launch_multigpu")

    def launch_whitted(self, programs, d_params, params_size, sbt, width, height):
        self.check(self.lib.b200rt_launch_whitted(self.h, self.stream, d_params, C.byref(sbt), width, height), "launch_whitted")

    def launch_playground(self, programs, d_params, params_size, sbt, width, height, opts):
        self.check(self.lib.b200rt_launch_playground(self.h, self.stream, d_params, width, height, C.byref(opts)), "launch_playground")

    def launch_raycast(self, programs, d_params, sbt, width, height, ext):
        self.check(self.lib.b200rt_launch_raycast(self.h, self.stream, d_params, C.byref(sbt), width, height, ext), "launch_raycast")

    def triangle_input(self, vertices, indices=None, sbt_index=None, num_sbt=1, flags=None, vertex_stride=None, pre_transform=None,
                       prim_offset=0):
        """Make one OptixBuildInput-compatible triangle input from device tensors.
        vertices: float32 tensor (any shape, contiguous); stride defaults to 12 ((n,3)) or 16 ((n,4))."""
        bi = L.BuildInput()
        bi.type = L.BUILD_INPUT_TYPE_TRIANGLES
        ta = bi.triangleArray
        if vertex_stride is None:
            vertex_stride = vertices.shape[-1] * 4
        nverts = vertices.numel() * 4 // vertex_stride
        vb = (C.c_uint64 * 1)(vertices.data_ptr())
        ta.vertexBuffers = vb
        ta.numVertices = nverts
        ta.vertexFormat = L.VERTEX_FORMAT_FLOAT3
        ta.vertexStrideInBytes = vertex_stride
        keep = [vertices, vb]
        if indices is not None:
            ta.indexBuffer = indices.data_ptr()
            ta.numIndexTriplets = indices.numel() // 3
            if indices.dtype == torch.int16 or indices.dtype == torch.uint16:
                ta.indexFormat = L.INDICES_FORMAT_UNSIGNED_SHORT3
                ta.indexStrideInBytes = 6
            else:
                ta.indexFormat = L.INDICES_FORMAT_UNSIGNED_INT3
                ta.indexStrideInBytes = 12
            keep.append(indices)
        fl = (C.c_uint32 * num_sbt)(*(flags if flags is not None else [L.GEOMETRY_FLAG_DISABLE_ANYHIT] * num_sbt))
        ta.flags = fl
        ta.numSbtRecords = num_sbt
        keep.append(fl)
        if sbt_index is not None:
            ta.sbtIndexOffsetBuffer = sbt_index.data_ptr()
            ta.sbtIndexOffsetSizeInBytes = sbt_index.element_size()
            ta.sbtIndexOffsetStrideInBytes = sbt_index.element_size()
            keep.append(sbt_index)
        if pre_transform is not None:
            ta.preTransform = pre_transform.data_ptr()
            ta.transformFormat = L.TRANSFORM_FORMAT_MATRIX_FLOAT12
            keep.append(pre_transform)
        ta.primitiveIndexOffset = prim_offset
        bi._keep = keep
        return bi

    def instance_input(self, instances):
        """instances: list of (transform12, sbt_offset, Accel[, visibility_mask[, OptixInstanceFlags]]).  SDK/sutil/Scene.cpp:1134-1212."""
        n = len(instances)
        arr = (L.Instance * n)()
        keep = []
        for i, inst in enumerate(instances):
            xf, sbt_off, acc = inst[:3]
            arr[i].transform = (C.c_float * 12)(*[float(x) for x in np.asarray(xf, np.float32).reshape(12)])
            arr[i].instanceId = i
            arr[i].sbtOffset = sbt_off
            arr[i].visibilityMask = inst[3] if len(inst) > 3 else 1
            arr[i].flags = inst[4] if len(inst) > 4 else 0
            arr[i].traversableHandle = acc.handle
            keep.append(acc)
        dev = self.to_device(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
        bi = L.BuildInput()
        bi.type = L.BUILD_INPUT_TYPE_INSTANCES
        bi.instanceArray.instances = dev.data_ptr()
        bi.instanceArray.numInstances = n
        bi.instanceArray.instanceStride = 0
        bi._keep = keep + [dev]
        return bi

    # ---- queries ---------------------------------------------------------------------------------
    def trace_closest(self, accel, rays, ray_flags=0):
        n = rays.shape[0]
        ext = torch.empty((n, 5), dtype=torch.int32, device=self.torch_device)
        self.check(self.lib.b200rt_trace_closest(self.h, self.stream, accel.handle, rays.data_ptr(), n, ray_flags, ext.data_ptr()), "trace_closest")
        return ext

    def trace_any(self, accel, rays, ray_flags=L.RAY_FLAG_TERMINATE_ON_FIRST_HIT):
        n = rays.shape[0]
        occ = torch.empty(n, dtype=torch.int32, device=self.torch_device)
        self.check(self.lib.b200rt_trace_any(self.h, self.stream, accel.handle, rays.data_ptr(), n, ray_flags, occ.data_ptr()), "trace_any")
        return occ

    def trace_stats(self, accel, rays):
        a, b = C.c_uint64(), C.c_uint64()
        self.check(self.lib.b200rt_trace_stats(self.h, self.stream, accel.handle, rays.data_ptr(), rays.shape[0], C.byref(a), C.byref(b)), "trace_stats")
        return a.value, b.value


class Accel:
    def __init__(self, ctx, buf, handle, keep=()):
        self.ctx, self.buf, self.handle, self.keep = ctx, buf, handle, list(keep)

    def info(self):
        inf = L.AccelInfo()
        self.ctx.check(self.ctx.lib.b200rt_accel_get_info(self.ctx.h, self.handle, C.byref(inf)), "accel_get_info")
        return inf


def ext_hits_to_numpy(ext):
    a = ext.cpu().numpy()
    return {"t": a[:, 0].view(np.float32).copy(), "prim": a[:, 1].astype(np.uint32), "inst": a[:, 2].astype(np.uint32),
            "b1": a[:, 3].view(np.float32).copy(), "b2": a[:, 4].view(np.float32).copy()}


# ---- host sutil mirrors -----------------------------------------------------------------------------
def camera_uvw(eye, lookat, up, fov_y, aspect):
    lib = L.load()
    U, V, W = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    lib.b200rt_camera_uvw(_f3(eye), _f3(lookat), _f3(up), fov_y, aspect, U, V, W)
    return np.array(U, np.float32), np.array(V, np.float32), np.array(W, np.float32)


def wd_num_samples(w, h, ngpu):
    return L.load().b200rt_wd_num_samples(w, h, ngpu)


def wd_sample_pixel(w, h, ngpu, gpu, sample):
    xy = (C.c_int32 * 2)()
    L.load().b200rt_wd_sample_pixel(w, h, ngpu, gpu, sample, xy)
    return xy[0], xy[1]


# ---- reference launch structs -------------------------------------------------------------------------
class ParallelogramLight(C.Structure):
    _fields_ = [("corner", C.c_float * 3), ("v1", C.c_float * 3), ("v2", C.c_float * 3), ("normal", C.c_float * 3),
                ("emission", C.c_float * 3)]


class PTParams(C.Structure):  # SDK/optixPathTracer/optixPathTracer.h:91-107 (152 bytes)
    _fields_ = [("subframe_index", C.c_uint32), ("accum_buffer", C.c_uint64), ("frame_buffer", C.c_uint64), ("width", C.c_uint32),
                ("height", C.c_uint32), ("samples_per_launch", C.c_uint32), ("eye", C.c_float * 3), ("U", C.c_float * 3),
                ("V", C.c_float * 3), ("W", C.c_float * 3), ("light", ParallelogramLight), ("handle", C.c_uint64)]


class MGParams(C.Structure):  # SDK/optixMultiGPU/optixMultiGPU.h:46-64 (168 bytes)
    _fields_ = [("subframe_index", C.c_uint32), ("sample_index_buffer", C.c_uint64), ("sample_accum_buffer", C.c_uint64),
                ("result_buffer", C.c_uint64), ("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_launch", C.c_uint32),
                ("device_idx", C.c_uint32), ("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3),
                ("light", ParallelogramLight), ("handle", C.c_uint64)]


class RaycastParams(C.Structure):  # SDK/optixRaycasting/optixRaycasting.h:41-46 (24 bytes)
    _fields_ = [("handle", C.c_uint64), ("rays", C.c_uint64), ("hits", C.c_uint64)]


assert C.sizeof(PTParams) == 152 and C.sizeof(MGParams) == 168 and C.sizeof(RaycastParams) == 24


def load_cornell():
    d = json.loads((DATA_DIR / "cornell.json").read_text())
    d["vertices"] = np.array(d["vertices"], np.float32)
    d["mat_indices"] = np.array(d["mat_indices"], np.uint32)
    d["emission_colors"] = np.array(d["emission_colors"], np.float32)
    d["diffuse_colors"] = np.array(d["diffuse_colors"], np.float32)
    return d


def light_normal(v1, v2):
    """normalize(cross(v1, v2)) as the reference's host vec_math computes it (optixPathTracer.cpp:439)."""
    v1, v2 = np.asarray(v1, np.float32), np.asarray(v2, np.float32)
    c = np.array([v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]], np.float32)
    d = np.float32(c[0] * c[0]) + np.float32(c[1] * c[1]) + np.float32(c[2] * c[2])
    return (c * (np.float32(1.0) / np.sqrt(d, dtype=np.float32))).astype(np.float32)


class ParamsUploader:
    """Params -> device copy of a launch (optixPathTracer.cpp:491-495: cudaMemcpyAsync on the launch's stream).  The launches are
    asynchronous, so the host may be several subframes ahead of the device: the pinned staging block of a copy must not be rewritten
    before that copy has run.  A small ring of pinned blocks, each guarded by an event recorded behind its copy."""

    def __init__(self, nbytes, device, slots=4):
        self.h = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.ev = [None] * slots
        self.d = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.k = 0

    def upload(self, struct):
        k = self.k
        self.k = (k + 1) % len(self.h)
        if self.ev[k] is not None:
            self.ev[k].synchronize()
        self.h[k].numpy()[:] = np.frombuffer(bytes(struct), np.uint8)
        self.d.copy_(self.h[k], non_blocking=True)
        self.ev[k] = torch.cuda.Event()
        self.ev[k].record(torch.cuda.current_stream(self.d.device))
        return self.d


class PathTracer:
    """Mirror of optixPathTracer's state + launchSubframe (optixPathTracer.cpp:424-511,576-898), and of
    optixMultiGPU's per-device state when multigpu=(gpu_idx, num_gpus) is given (optixMultiGPU.cpp:479-594).

    vertices: (ntri*3, 4) float32 (the reference's Vertex{x,y,z,pad}); mat_indices: (ntri,) uint32."""

    def __init__(self, ctx, width, height, samples_per_launch=16, scene=None, vertices=None, mat_indices=None, multigpu=None,
                 compact=True):
        self.ctx = ctx
        sc = scene or load_cornell()
        self.scene = sc
        self.width, self.height, self.spl = width, height, samples_per_launch
        dev = ctx.torch_device
        if vertices is None:
            v4 = np.zeros((sc["vertices"].shape[0], 4), np.float32)
            v4[:, :3] = sc["vertices"]
            vertices = ctx.to_device(v4)
            mat_indices = ctx.to_device(sc["mat_indices"])
        self.d_vertices, self.d_mat = vertices, mat_indices
        nmat = len(sc["emission_colors"])
        # buildMeshAccel (optixPathTracer.cpp:576-684): stride-16 float3 vertices, per-primitive u32 SBT index
        bi = ctx.triangle_input(vertices, sbt_index=mat_indices, num_sbt=nmat, vertex_stride=16)
        self.build_input = bi
        self.accel = ctx.build_accel([bi], compact=compact)
        # createSBT (optixPathTracer.cpp:829-898 / optixMultiGPU.cpp:960-1018)
        self.multigpu = multigpu
        ray_types = 2 if multigpu else 1
        self.programs = ctx.prepare_programs("multigpu" if multigpu else "pathtracer")
        rec = np.zeros((nmat * ray_types, 64), np.uint8)
        for i in range(nmat):
            data = struct.pack("<3f3fQ", *sc["emission_colors"][i], *sc["diffuse_colors"][i], vertices.data_ptr())
            rec[i * ray_types, 0:32] = np.frombuffer(ctx.sbt_header(self.programs, 2, 0), np.uint8)
            rec[i * ray_types, 32:64] = np.frombuffer(data, np.uint8)
            if ray_types == 2:  # zeroed occlusion record (optixMultiGPU.cpp:1000-1006)
                rec[i * ray_types + 1, 0:32] = np.frombuffer(ctx.sbt_header(self.programs, 2, 1), np.uint8)
        self.d_hitgroup = ctx.to_device(rec)
        miss = np.zeros((ray_types, 48), np.uint8)  # MissData{bg_color = 0}
        for r in range(ray_types):
            miss[r, 0:32] = np.frombuffer(ctx.sbt_header(self.programs, 1, r), np.uint8)
        self.d_miss = ctx.to_device(miss)
        rg = np.zeros(32, np.uint8)
        rg[:] = np.frombuffer(ctx.sbt_header(self.programs, 0, 0), np.uint8)
        self.d_raygen = ctx.to_device(rg)
        self.sbt = L.ShaderBindingTable()
        self.sbt.raygenRecord = self.d_raygen.data_ptr()
        self.sbt.missRecordBase = self.d_miss.data_ptr()
        self.sbt.missRecordStrideInBytes = 48
        self.sbt.missRecordCount = ray_types
        self.sbt.hitgroupRecordBase = self.d_hitgroup.data_ptr()
        self.sbt.hitgroupRecordStrideInBytes = 64
        self.sbt.hitgroupRecordCount = nmat * ray_types
        # initLaunchParams + handleCameraUpdate (optixPathTracer.cpp:424-470)
        cam, lt = sc["camera"], sc["light"]
        U, V, W = camera_uvw(cam["eye"], cam["lookat"], cam["up"], cam["fov_y"], width / float(height))
        light = ParallelogramLight(_f3(lt["corner"]), _f3(lt["v1"]), _f3(lt["v2"]), _f3(light_normal(lt["v1"], lt["v2"])), _f3(lt["emission"]))
        self.frame = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev)
        if multigpu:
            gpu_idx, num_gpus = multigpu
            self.num_samples = wd_num_samples(width, height, num_gpus)
            self.sample_index = torch.zeros((self.num_samples, 2), dtype=torch.int32, device=dev)
            hc = getattr(ctx, "helper", ctx)  # fillSamplesCUDA is a plain CUDA kernel in the reference too (optixMultiGPU_kernels.cu)
            hc.check(hc.lib.b200rt_fill_samples(hc.h, hc.stream, gpu_idx, num_gpus, width, height, self.sample_index.data_ptr(),
                                                self.num_samples), "fill_samples")
            self.accum = torch.zeros((self.num_samples, 4), dtype=torch.float32, device=dev)
            self.params = MGParams(0, self.sample_index.data_ptr(), self.accum.data_ptr(), self.frame.data_ptr(), width, height,
                                   samples_per_launch, 3, _f3(cam["eye"]), _f3(U), _f3(V), _f3(W), light, self.accel.handle)
        else:
            self.accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
            self.params = PTParams(0, self.accum.data_ptr(), self.frame.data_ptr(), width, height, samples_per_launch, _f3(cam["eye"]),
                                   _f3(U), _f3(V), _f3(W), light, self.accel.handle)
        self.uploader = ParamsUploader(C.sizeof(self.params), dev)
        self.d_params = self.uploader.d
        self.stats = L.PTStats()
        self.sample_groups = 1
        self.ray_sort = 0  # b200rt_pt_options.ray_sort

    def launch_subframe(self, subframe_index=None, collect_stats=False, sample_groups=None):
        """launchSubframe (optixPathTracer.cpp:488-511): copy Params to the device, launch; asynchronous.
        sample_groups: b200rt_pt_options.sample_groups (None = the instance default self.sample_groups)."""
        if subframe_index is not None:
            self.params.subframe_index = subframe_index
        self.uploader.upload(self.params)
        groups = self.sample_groups if sample_groups is None else sample_groups
        opts = L.PTOptions(int(groups), int(collect_stats), C.pointer(self.stats), int(self.ray_sort), 0)  # collect_stats: bit mask of L.PT_STATS_*
        ctx = self.ctx
        if self.multigpu:
            ctx.launch_multigpu(self.programs, self.d_params.data_ptr(), C.sizeof(self.params), self.sbt, self.num_samples, opts)
        else:
            ctx.launch_pathtracer(self.programs, self.d_params.data_ptr(), C.sizeof(self.params), self.sbt, self.width, self.height, opts)
        return self.stats if collect_stats else None


# ---- glTF (the subset sutil::Scene loads: SDK/sutil/Scene.cpp:84-210,267-550) ---------------------------
_GLTF_COMP = {5123: (np.uint16, 2), 5125: (np.uint32, 4), 5126: (np.float32, 4)}  # bufferViewFromGLTF: u16, u32, f32 — anything else throws
_GLTF_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}
_F = np.float32


def _mat4_mul(a, b):
    """sutil::Matrix operator* (SDK/sutil/Matrix.h:339-355): fp32, sum = 0; sum += a[i][k] * b[k][j] for k = 0..3 — spelled out, because
    a BLAS product adds in another order and the instance transforms are compared bit for bit with sutil's."""
    a, b = np.asarray(a, _F), np.asarray(b, _F)
    out = np.zeros((4, 4), _F)
    for i in range(4):
        for j in range(4):
            acc = _F(0.0)
            for k in range(4):
                acc = _F(acc + _F(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def _mat4_vec(m, v):
    """Matrix4x4 * float4 (SDK/sutil/Matrix.h:467-487): ((m0 x + m1 y) + m2 z) + m3 w per row, fp32."""
    m, v = np.asarray(m, _F), np.asarray(v, _F)
    return np.array([_F(_F(_F(_F(m[r, 0] * v[0]) + _F(m[r, 1] * v[1])) + _F(m[r, 2] * v[2])) + _F(m[r, 3] * v[3])) for r in range(4)], _F)


def _aabb_transform(lo, hi, m):
    """Aabb::transform(Matrix4x4) (SDK/sutil/Aabb.h:389-409): the eight corners through the matrix, min / max."""
    pts = [_mat4_vec(m, [x, y, z, 1.0])[:3] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]
    return np.min(pts, axis=0).astype(_F), np.max(pts, axis=0).astype(_F)


def _node_matrices(node):
    """translation, rotation, scale, matrix of processGLTFNode (SDK/sutil/Scene.cpp:132-163); doubles of the file cast to fp32 first."""
    f = _F
    t = np.eye(4, dtype=f)
    if node.get("translation"):
        t[:3, 3] = np.array(node["translation"], f)
    r = np.eye(4, dtype=f)
    if node.get("rotation"):
        qx, qy, qz, qw = [f(v) for v in node["rotation"]]
        two = f(2.0)
        # sutil::Quaternion::rotationMatrix (SDK/sutil/Quaternion.h:239-267), left to right in fp32
        r[0, 0] = f(f(f(1.0) - f(f(two * qy) * qy)) - f(f(two * qz) * qz)); r[0, 1] = f(f(f(two * qx) * qy) - f(f(two * qz) * qw)); r[0, 2] = f(f(f(two * qx) * qz) + f(f(two * qy) * qw))
        r[1, 0] = f(f(f(two * qx) * qy) + f(f(two * qz) * qw)); r[1, 1] = f(f(f(1.0) - f(f(two * qx) * qx)) - f(f(two * qz) * qz)); r[1, 2] = f(f(f(two * qy) * qz) - f(f(two * qx) * qw))
        r[2, 0] = f(f(f(two * qx) * qz) - f(f(two * qy) * qw)); r[2, 1] = f(f(f(two * qy) * qz) + f(f(two * qx) * qw)); r[2, 2] = f(f(f(1.0) - f(f(two * qx) * qx)) - f(f(two * qy) * qy))
    sc = np.eye(4, dtype=f)
    if node.get("scale"):
        sc[0, 0], sc[1, 1], sc[2, 2] = [f(v) for v in node["scale"]]
    m = np.eye(4, dtype=f)
    if node.get("matrix"):
        m = np.array(node["matrix"], f).reshape(4, 4).T.copy()  # column-major in the file
    return m, t, r, sc


def load_gltf(path):
    """sutil::loadScene (SDK/sutil/Scene.cpp:267-550) for a .gltf with external buffers.  Returns a dict:
      buffers    the glTF buffers, whole (Scene::addBuffer uploads each once; every view points into them)
      meshes     [{primitives: [{positions, normals, texcoords[2], colors, indices (decoded copies, for CPU-side use),
                                 views: {name: (buffer, byte offset, count, byte stride, element size)} — what bufferViewFromGLTF
                                 (Scene.cpp:84-123) makes: nothing is repacked, interleaved attributes stay interleaved —, material}], aabb}]
      instances  [{transform (4x4 fp32), mesh, world_aabb}] in the order processGLTFNode meets them, starting from every node that is
                 nobody's child (Scene.cpp:535-549: not from scenes[scene].nodes)
      cameras, materials, images, textures, samplers."""
    path = pathlib.Path(path)
    g = json.loads(path.read_text())
    buffers = [np.fromfile(path.parent / b["uri"], dtype=np.uint8) for b in g["buffers"]]

    def view(idx):
        """bufferViewFromGLTF: (buffer, offset, count, stride, elmt) — elmt is the COMPONENT size, as in the reference"""
        a = g["accessors"][idx]
        bv = g["bufferViews"][a["bufferView"]]
        if a["componentType"] not in _GLTF_COMP:
            raise B200RTError("gltf accessor component type not supported")
        dt, sz = _GLTF_COMP[a["componentType"]]
        stride = bv.get("byteStride", 0) or sz * _GLTF_NCOMP.get(a["type"], 1)
        return (bv["buffer"], bv.get("byteOffset", 0) + a.get("byteOffset", 0), a["count"], stride, sz), dt, a

    def decode(v, dt, ncomp):
        buf, off, count, stride, sz = v
        raw = buffers[buf]
        out = np.zeros((count, ncomp), dt)
        for i in range(ncomp):
            # the last element of a view read wider than its accessor (COLOR_0 as Vec4f over a VEC3 accessor) may end past the buffer
            n_ok = min(count, max(0, (raw.size - off - i * sz - sz) // stride + 1)) if raw.size >= off + i * sz + sz else 0
            if n_ok:
                out[:n_ok, i] = np.ndarray((n_ok,), dt, raw.data, off + i * sz, (stride,))
        return out

    meshes = []
    for m in g["meshes"]:
        prims = []
        lo = np.full(3, np.inf, np.float32)
        hi = np.full(3, -np.inf, np.float32)
        for p in m["primitives"]:
            if p.get("mode", 4) != 4:  # TINYGLTF_MODE_TRIANGLES only
                continue
            views = {}
            att = p["attributes"]
            idx = None
            if "indices" in p:
                views["indices"], dt, _ = view(p["indices"])
                idx = decode(views["indices"], dt, 1).reshape(-1)
            views["positions"], dt, pa = view(att["POSITION"])
            pos = decode(views["positions"], dt, 3).astype(np.float32)
            if pa.get("min") and pa.get("max"):
                lo = np.minimum(lo, np.array(pa["min"], np.float32))
                hi = np.maximum(hi, np.array(pa["max"], np.float32))
            nrm = None
            if "NORMAL" in att:
                views["normals"], dt, _ = view(att["NORMAL"])
                nrm = decode(views["normals"], dt, 3).astype(np.float32)
            uvs = [None, None]
            for j in range(2):
                if f"TEXCOORD_{j}" in att:
                    views[f"texcoords{j}"], dt, _ = view(att[f"TEXCOORD_{j}"])
                    uvs[j] = decode(views[f"texcoords{j}"], dt, 2).astype(np.float32)
            col = None
            if "COLOR_0" in att:
                # BufferView<Vec4f> whatever the accessor's type (Scene.cpp:520-524): a VEC3 colour is read four floats at a time
                views["colors"], dt, _ = view(att["COLOR_0"])
                col = decode(views["colors"], dt, 4).astype(np.float32)
            prims.append({"positions": pos, "normals": nrm, "indices": idx, "material": p.get("material", -1), "texcoords": uvs, "colors": col,
                          "views": views})
        meshes.append({"name": m.get("name", ""), "primitives": prims, "aabb": (lo, hi)})
    instances = []
    cameras = []

    def walk(ni, parent):
        node = g["nodes"][ni]
        m, t, r, sc = _node_matrices(node)
        xf = _mat4_mul(_mat4_mul(_mat4_mul(_mat4_mul(parent, m), t), r), sc)  # parent * matrix * translation * rotation * scale
        if "camera" in node:
            # processGLTFNode (Scene.cpp:166-192): eye = M (0,0,0,1), up = M (0,1,0,0), fovY in degrees
            cam = g["cameras"][node["camera"]]
            if cam.get("type") != "perspective":
                return  # the reference returns here: the children of a non-perspective camera node are not visited
            yfov = np.float32(np.float32(cam["perspective"]["yfov"]) * np.float32(180.0)) / np.float32(math.pi)
            cameras.append({"eye": _mat4_vec(xf, [0, 0, 0, 1])[:3], "up": _mat4_vec(xf, [0, 1, 0, 0])[:3], "fov_y": float(yfov),
                            "aspect": float(np.float32(cam["perspective"].get("aspectRatio", 0.0)))})
        elif "mesh" in node:
            lo, hi = meshes[node["mesh"]]["aabb"]
            instances.append({"transform": xf, "mesh": node["mesh"], "world_aabb": _aabb_transform(lo, hi, xf)})
        for c in node.get("children", []):
            walk(c, xf)

    is_root = [True] * len(g.get("nodes", []))
    for node in g.get("nodes", []):
        for c in node.get("children", []):
            is_root[c] = False
    for ni, root in enumerate(is_root):
        if root:
            walk(ni, np.eye(4, dtype=np.float32))
    return {"buffers": buffers, "meshes": meshes, "instances": instances, "cameras": cameras, "materials": _gltf_materials(g),
            "images": _gltf_images(g, path.parent), "textures": g.get("textures", []), "samplers": g.get("samplers", [])}


def _gltf_materials(g):
    """The MaterialData fields sutil::loadScene fills (Scene.cpp:350-443); defaults as MaterialData() (MaterialData.h:44-52: metallic = roughness = 1)."""
    out = []
    for m in g.get("materials", []):
        pbr = m.get("pbrMetallicRoughness", {})

        def tex(info):
            if not info:
                return None
            xf = info.get("extensions", {}).get("KHR_texture_transform", {})
            return {"index": info["index"], "texcoord": xf.get("texCoord", info.get("texCoord", 0)), "offset": xf.get("offset", [0.0, 0.0]),
                    "rotation": xf.get("rotation", 0.0), "scale": xf.get("scale", [1.0, 1.0])}
        out.append({"base_color": pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0]), "metallic": pbr.get("metallicFactor", 1.0),
                    "roughness": pbr.get("roughnessFactor", 1.0), "base_color_tex": tex(pbr.get("baseColorTexture")),
                    "metallic_roughness_tex": tex(pbr.get("metallicRoughnessTexture")), "normal_tex": tex(m.get("normalTexture")),
                    "emissive_tex": tex(m.get("emissiveTexture")), "emissive_factor": m.get("emissiveFactor", [0.0, 0.0, 0.0]),
                    "alpha_mode": {"OPAQUE": 0, "MASK": 1, "BLEND": 2}[m.get("alphaMode", "OPAQUE")], "alpha_cutoff": m.get("alphaCutoff", 0.5),
                    "double_sided": bool(m.get("doubleSided", False))})
    return out


def _gltf_images(g, base):
    """8-bit RGBA pixels of every glTF image, 4 channels always as tinygltf/stb decode them (tiny_gltf.h:2368)."""
    out = []
    for im in g.get("images", []):
        try:
            from PIL import Image
            out.append(np.ascontiguousarray(np.array(Image.open(base / im["uri"]).convert("RGBA"), np.uint8)))
        except Exception:  # image library or file missing: the material falls back to its factors
            out.append(None)
    return out


# byte offsets inside whitted::HitGroupData (SDK/cuda/whitted.h:44-48, SDK/cuda/GeometryData.h:73-80,248-262)
HG_OFF_INDICES, HG_OFF_POSITIONS, HG_OFF_NORMALS, HG_OFF_MATERIAL, HG_SIZE = 16, 32, 48, 112, 352


def create_scene_textures(ctx, scene):
    """Scene::addImage / addSampler (Scene.cpp:576-652): one texture object per glTF texture (sampler wrap modes; linear unless NEAREST).
    Returns ({texture index: cudaTextureObject_t or None}, [(context, texture, array)] to destroy)."""
    hc = getattr(ctx, "helper", ctx)
    tex_objects, handles = {}, []
    wrap = {10497: 0, 33071: 1, 33648: 2}
    for ti, t in enumerate(scene.get("textures", [])):
        img = scene["images"][t["source"]] if t.get("source") is not None else None
        if img is None:
            tex_objects[ti] = None
            continue
        smp = scene["samplers"][t["sampler"]] if t.get("sampler") is not None and scene.get("samplers") else {}
        linear = 0 if smp.get("magFilter") == 9728 else 1
        tex, arr = C.c_uint64(), C.c_uint64()
        hc.check(hc.lib.b200rt_texture_create(hc.h, img.shape[1], img.shape[0], img.ctypes.data_as(C.c_void_p), wrap.get(smp.get("wrapS", 10497), 0),
                                              wrap.get(smp.get("wrapT", 10497), 0), linear, C.byref(tex), C.byref(arr)), "texture_create")
        tex_objects[ti] = tex.value
        handles.append((hc, tex.value, arr.value))
    return tex_objects, handles


def create_scene_textures_shared(ctxs, scene, islands, share=True):
    """optixNVLink's texture sharing for ONE process driving several devices (loadTextures, SDK/optixNVLink/optixNVLink.cpp:1501-1590):
    each P2P island keeps one CUDA array per texture, on its device with the least texture memory so far; the other devices of the island
    get their own texture object over that array (b200rt_texture_view) and sample it over NVLink.  `ctxs`: one Context per device, in
    device order, peer access enabled inside the islands (Context.enable_peer_access).  Returns a list, one entry per context, of what
    create_scene_textures returns: ({texture index: cudaTextureObject_t or None}, [(context, texture, array)] to destroy), plus the
    per-device texture bytes."""
    from . import topology
    wrap = {10497: 0, 33071: 1, 33648: 2}
    tex = scene.get("textures", [])
    imgs = [scene["images"][t["source"]] if t.get("source") is not None else None for t in tex]
    sizes = [0.0 if im is None else float(im.shape[0] * im.shape[1] * 4) for im in imgs]
    owners, usage = topology.plan_texture_sharing(islands, sizes, len(ctxs), share)
    out = [({}, []) for _ in ctxs]
    for ti, (t, img) in enumerate(zip(tex, imgs)):
        if img is None:
            for objs, _ in out:
                objs[ti] = None
            continue
        smp = scene["samplers"][t["sampler"]] if t.get("sampler") is not None and scene.get("samplers") else {}
        linear = 0 if smp.get("magFilter") == 9728 else 1
        ws, wt = wrap.get(smp.get("wrapS", 10497), 0), wrap.get(smp.get("wrapT", 10497), 0)
        arrays = {}
        for d in sorted(set(owners[ti])):   # the copies first
            c = ctxs[d]
            to, arr = C.c_uint64(), C.c_uint64()
            c.check(c.lib.b200rt_texture_create(c.h, img.shape[1], img.shape[0], img.ctypes.data_as(C.c_void_p), ws, wt, linear, C.byref(to), C.byref(arr)),
                    "texture_create")
            arrays[d] = arr.value
            out[d][0][ti] = to.value
            out[d][1].append((c, to.value, arr.value))
        for d, own in enumerate(owners[ti]):  # then the views of the other devices of the island
            if own == d:
                continue
            c = ctxs[d]
            to = C.c_uint64()
            c.check(c.lib.b200rt_texture_view(c.h, arrays[own], ws, wt, linear, C.byref(to)), "texture_view")
            out[d][0][ti] = to.value
            out[d][1].append((c, to.value, 0))
    return out, usage


def build_scene_meshes(ctx, scene, tex_objects):
    """Scene::buildMeshAccels (Scene.cpp:817-1132) + the whitted::HitGroupData payload of every primitive group: one GAS per mesh, one build
    input per primitive group with the geometry flags of its material (Scene.cpp:904-966: OPAQUE -> DISABLE_ANYHIT, MASK -> NONE, BLEND ->
    REQUIRE_SINGLE_ANYHIT_CALL; doubleSided adds DISABLE_TRIANGLE_FACE_CULLING).  Returns (accels, per-mesh record payloads, device buffers)."""
    materials = scene.get("materials", [])
    keep, mesh_accels, mesh_records = [], [], []
    # Scene::addBuffer (Scene.cpp:562-574): every glTF buffer goes to the device once, whole; all views of a loaded file point into these
    d_buffers = [ctx.to_device(b) for b in scene.get("buffers", [])]
    keep += d_buffers
    for m in scene["meshes"]:
        inputs, recs = [], []
        for p in m["primitives"]:
            mat = materials[p["material"]] if p.get("material", -1) >= 0 and p["material"] < len(materials) else None
            am = 0 if mat is None else mat["alpha_mode"]
            gf = {0: 1, 1: 0, 2: 2}[am] | (4 if (mat is not None and mat["double_sided"]) else 0)
            geo = bytearray(112)
            views = p.get("views") if d_buffers else None
            if views:
                # the reference's own memory layout: BufferView = {buffer base + byte offset, count, byte stride, COMPONENT size} (Scene.cpp:84-123)
                def bview(name):
                    if name not in views:
                        return struct.pack("<QIHH", 0, 0, 0, 0)
                    buf, off, count, stride, elmt = views[name]
                    return struct.pack("<QIHH", d_buffers[buf].data_ptr() + off, count, stride, elmt)
                bi = L.BuildInput()
                bi.type = L.BUILD_INPUT_TYPE_TRIANGLES
                ta = bi.triangleArray
                pbuf, poff, pcount, pstride, _ = views["positions"]
                vb = (C.c_uint64 * 1)(d_buffers[pbuf].data_ptr() + poff)
                ta.vertexBuffers, ta.numVertices, ta.vertexFormat, ta.vertexStrideInBytes = vb, pcount, L.VERTEX_FORMAT_FLOAT3, pstride
                if "indices" in views:
                    ibuf, ioff, icount, istride, ielmt = views["indices"]
                    ta.indexBuffer, ta.numIndexTriplets = d_buffers[ibuf].data_ptr() + ioff, icount // 3
                    ta.indexFormat = L.INDICES_FORMAT_UNSIGNED_SHORT3 if ielmt == 2 else L.INDICES_FORMAT_UNSIGNED_INT3
                    ta.indexStrideInBytes = istride * 3  # Scene.cpp:930
                fl = (C.c_uint32 * 1)(gf)
                ta.flags, ta.numSbtRecords = fl, 1
                bi._keep = [vb, fl]
                inputs.append(bi)
                geo[HG_OFF_INDICES:HG_OFF_INDICES + 16] = bview("indices")
                geo[HG_OFF_POSITIONS:HG_OFF_POSITIONS + 16] = bview("positions")
                geo[HG_OFF_NORMALS:HG_OFF_NORMALS + 16] = bview("normals")
                geo[64:80] = bview("texcoords0")
                geo[80:96] = bview("texcoords1")
                geo[96:112] = bview("colors")
                recs.append(bytes(geo) + pack_material(mat, tex_objects))
                continue
            d_pos = ctx.to_device(p["positions"])
            d_nrm = ctx.to_device(p["normals"]) if p.get("normals") is not None else None
            idx = p["indices"]
            d_idx = None
            if idx is not None:
                d_idx = ctx.to_device(idx.astype(np.uint16).view(np.int16) if idx.dtype == np.uint16 else idx.astype(np.uint32).view(np.int32))
            uvs = p.get("texcoords") or [None, None]
            d_uv = [ctx.to_device(u) if u is not None else None for u in uvs]
            inputs.append(ctx.triangle_input(d_pos, indices=d_idx, num_sbt=1, flags=[gf], vertex_stride=12))
            keep += [d_pos, d_nrm, d_idx] + d_uv
            # whitted::HitGroupData: GeometryData{type=TRIANGLE_MESH(0) @0, 16-byte aligned union @16 holding TriangleMesh{indices, positions,
            # normals, texcoords[2], colors}}; BufferView = {ptr, count, u16 stride, u16 elmt}.  Offsets pinned against the reference
            # headers: tests/golden/kat.json "hitgroup_layout" (tests/test_oracle_kat.py)

            def bview(t, elmt, stride):
                return struct.pack("<QIHH", t.data_ptr() if t is not None else 0, (t.shape[0] if t is not None else 0), stride if t is not None else 0,
                                   elmt if t is not None else 0)
            isz = 0 if idx is None else (2 if idx.dtype == np.uint16 else 4)
            geo[HG_OFF_INDICES:HG_OFF_INDICES + 16] = struct.pack("<QIHH", d_idx.data_ptr() if d_idx is not None else 0, 0 if idx is None else idx.shape[0], isz, isz)
            geo[HG_OFF_POSITIONS:HG_OFF_POSITIONS + 16] = bview(d_pos, 12, 12)
            geo[HG_OFF_NORMALS:HG_OFF_NORMALS + 16] = bview(d_nrm, 12, 12)
            geo[64:80] = bview(d_uv[0], 8, 8)
            geo[80:96] = bview(d_uv[1], 8, 8)
            recs.append(bytes(geo) + pack_material(mat, tex_objects))
        mesh_accels.append(ctx.build_accel(inputs, compact=True))
        mesh_records.append(recs)
    return mesh_accels, mesh_records, keep


class Raycaster:
    """Mirror of optixRaycasting's state (optixRaycasting.cpp:94-349) for a loaded glTF scene: one GAS per mesh
    (one build input per primitive group), an IAS over the mesh instances, whitted::HitGroupData SBT records."""

    def __init__(self, ctx, scene, compact=True):
        self.ctx = ctx
        self.scene = scene
        dev = ctx.torch_device
        self.programs = ctx.prepare_programs("raycast")
        # createSBT (optixRaycasting.cpp:219-237): one record per primitive group holding the mesh views and the scene's MaterialData —
        # __anyhit__texture_mask reads alpha_mode / alpha_cutoff / base_color_tex from it
        self.tex_objects, self._tex_handles = create_scene_textures(ctx, scene)
        self.mesh_accels, mesh_records, self.keep = build_scene_meshes(ctx, scene, self.tex_objects)
        records, self.mesh_sbt_base = [], []
        for recs in mesh_records:
            self.mesh_sbt_base.append(len(records))
            records += [ctx.sbt_header(self.programs, 2, 0) + data for data in recs]
        self.d_hitgroup = ctx.to_device(np.frombuffer(b"".join(records), np.uint8).copy())
        inst = []
        for i in scene["instances"]:
            inst.append((i["transform"][:3, :].reshape(12), self.mesh_sbt_base[i["mesh"]], self.mesh_accels[i["mesh"]]))
        self.ias = ctx.build_accel([ctx.instance_input(inst)], compact=False)
        self.sbt = L.ShaderBindingTable()
        self.d_miss = ctx.to_device(np.frombuffer(ctx.sbt_header(self.programs, 1, 0), np.uint8).copy())
        self.d_raygen = ctx.to_device(np.frombuffer(ctx.sbt_header(self.programs, 0, 0), np.uint8).copy())
        self.sbt.raygenRecord = self.d_raygen.data_ptr()
        self.sbt.missRecordBase = self.d_miss.data_ptr()
        self.sbt.missRecordStrideInBytes = 32
        self.sbt.missRecordCount = 1
        self.sbt.hitgroupRecordBase = self.d_hitgroup.data_ptr()
        self.sbt.hitgroupRecordStrideInBytes = 32 + 352
        self.sbt.hitgroupRecordCount = len(records)
        lo = np.min([i["world_aabb"][0] for i in scene["instances"]], axis=0).astype(np.float32)
        hi = np.max([i["world_aabb"][1] for i in scene["instances"]], axis=0).astype(np.float32)
        self.bbmin, self.bbmax = lo, hi

    def buffer_rays(self, width):
        """bufferRays (optixRaycasting.cpp:255-286)."""
        # the ray-generation / translate / shade helpers are plain CUDA kernels in the reference too (optixRaycastingKernels.cu);
        # a context of another back end (oracle/optix_ref) borrows b200rt's through its `helper` attribute
        ctx, dev = getattr(self.ctx, "helper", self.ctx), self.ctx.torch_device
        span = (self.bbmax - self.bbmin).astype(np.float32)
        self.width = width
        self.height = int(np.float32(width) * span[1] / span[0])
        n = self.width * self.height
        self.rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
        ctx.check(ctx.lib.b200rt_create_rays_ortho(ctx.h, ctx.stream, self.rays.data_ptr(), self.width, self.height, _f3(self.bbmin),
                                                   _f3(self.bbmax), 0.05), "create_rays_ortho")
        self.rays_translated = self.rays.clone()
        off = (span * np.array([0.2, 0, 0], np.float32)).astype(np.float32)
        ctx.check(ctx.lib.b200rt_translate_rays(ctx.h, ctx.stream, self.rays_translated.data_ptr(), n, _f3(off)), "translate_rays")
        self.hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.hits_translated = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.ext = torch.empty((n, 5), dtype=torch.int32, device=dev)
        self.ext_translated = torch.empty((n, 5), dtype=torch.int32, device=dev)
        self.translate_offset = off
        ctx = self.ctx
        p1 = RaycastParams(self.ias.handle, self.rays.data_ptr(), self.hits.data_ptr())
        p2 = RaycastParams(self.ias.handle, self.rays_translated.data_ptr(), self.hits_translated.data_ptr())
        self.d_params = ctx.to_device(np.frombuffer(bytes(p1), np.uint8).copy())
        self.d_params_translated = ctx.to_device(np.frombuffer(bytes(p2), np.uint8).copy())
        return n

    def launch(self, want_ext=True):
        """launch (optixRaycasting.cpp:289-317): the two batches go to two streams, as in the reference, so they may overlap; the
        caller's stream waits for both."""
        ctx, dev = self.ctx, self.ctx.torch_device
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_streams"):
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        batches = ((self.d_params, self.ext), (self.d_params_translated, self.ext_translated))
        for st, (params, ext) in zip(self._streams, batches):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                ctx.launch_raycast(self.programs, params.data_ptr(), self.sbt, self.width, self.height, ext.data_ptr() if want_ext else 0)
        for st in self._streams:
            cur.wait_stream(st)

    def close(self):
        for hc, tex, arr in self._tex_handles:
            hc.lib.b200rt_texture_destroy(hc.h, tex, arr)
        self._tex_handles = []

    def shade(self, hits):
        ctx = getattr(self.ctx, "helper", self.ctx)
        n = hits.shape[0]
        img = torch.empty((n, 3), dtype=torch.float32, device=ctx.torch_device)
        ctx.check(ctx.lib.b200rt_shade_hits(ctx.h, ctx.stream, img.data_ptr(), n, hits.data_ptr()), "shade_hits")
        return img


def synthetic_mesh(ctx, num_triangles, seed=0):
    """Procedural tessellated scene of BASELINE.json configs[4] generated on the device."""
    dev = ctx.torch_device
    verts = torch.empty((num_triangles * 3, 4), dtype=torch.float32, device=dev)
    mats = torch.empty(num_triangles, dtype=torch.int32, device=dev)
    b = (C.c_float * 6)()
    ctx.check(ctx.lib.b200rt_generate_synthetic_mesh(ctx.h, ctx.stream, num_triangles, seed, verts.data_ptr(), mats.data_ptr(), b), "synthetic_mesh")
    return verts, mats


# ---- imgui_test ("playground"): SDK/imgui_test/main.cpp:236-279, tracer_window.cpp:64-128 -----------------------------------------
class PGParams(C.Structure):  # SDK/imgui_test/optixTriangle.h:42-108 (128 bytes; offsets pinned in tests/golden/kat.json "playground_layout")
    _fields_ = [("image_width", C.c_uint32), ("image_height", C.c_uint32), ("samples_per_frame", C.c_uint32), ("camera", C.c_uint64),
                ("dt", C.c_uint32), ("dirty", C.c_uint8), ("image", C.c_uint64), ("film", C.c_uint64), ("tfactor", C.c_float), ("handle", C.c_uint64),
                ("normals", C.c_uint64), ("vertices", C.c_uint64), ("mat_indices", C.c_uint64), ("nmat_indices", C.c_int32), ("lights", C.c_uint64),
                ("nlights", C.c_int32), ("materials", C.c_uint64), ("nmaterials", C.c_int32)]


assert C.sizeof(PGParams) == 128 and PGParams.lights.offset == 96 and PGParams.dirty.offset == 28


def playground_light(kind, a, lumi, scalar=0.0):
    """44 raw bytes of LightVariant (light.h:42-50): kind 0 PointLight(position, lumi), 1 DirectionalLight(direction, lumi, jitter),
    2 VolumetricLight(position, radius, lumi); each light ends with float3 m_dark = 0."""
    a, lumi = [float(x) for x in a], [float(x) for x in lumi]
    if kind == 0:
        return struct.pack("<10fi", *a, *lumi, 0.0, 0.0, 0.0, 0.0, 0)
    if kind == 1:
        return struct.pack("<10fi", *a, *lumi, float(scalar), 0.0, 0.0, 0.0, 1)
    return struct.pack("<10fi", *a, float(scalar), *lumi, 0.0, 0.0, 0.0, 2)


def playground_default_lights():
    """main.cpp:245-249"""
    return (playground_light(2, (0.0, 2.0, 0.0), (0.1, 0.08, 0.08), 0.1) + playground_light(2, (2.0, 2.0, 0.0), (0.1, 0.08, 0.08), 0.1)
            + playground_light(2, (2.0, 2.0, 2.0), (0.1, 0.08, 0.08), 0.1) + playground_light(1, (-1.0, 1.0, -1.0), (0.1, 0.1, 0.1), 0.05))


def playground_default_materials():
    """main.cpp:251-261: 28 DiffuseMaterial colours"""
    m = [(0.5, 0.5, 0.5)]
    for i in range(5):
        for j in range(5):
            m.append((np.float32(i) / np.float32(5.0), np.float32(j) / np.float32(5.0), 0.2))
    m += [(0.5, 0.5, 0.5), (0.5, 0.5, 0.5)]
    return np.array(m, np.float32)


def playground_camera(eye=(0.0, 1.0, -10.0), up=(0.0, 0.0000073, 1.0), lookat=(0.0, 0.1, 0.0), aperture=0.0, fd=1.0, fov=45.0, ortho=False):
    """92 raw bytes of the Camera main.cpp:236-243 sets up (defaults = the reference's values)."""
    buf = (C.c_uint8 * 92)()
    L.load().b200rt_playground_camera(_f3(eye), _f3(up), _f3(lookat), aperture, fd, fov, int(bool(ortho)), buf)
    return bytes(buf)


class Playground:
    """Mirror of imgui_test's state (main.cpp:64-279, triangle_gas.cpp:170-241) and of TracerWindow::run's per-frame work
    (tracer_window.cpp:88-105) without the window.  Geometry: the procedural stand-in (rows -> 25 blobs of 4*rows^2 triangles + floor),
    or caller-supplied (T*3,3) vertices / normals and (T,) int32 material indices (device tensors)."""

    def __init__(self, ctx, width, height, spf=5, rows=132, seed=0, camera=None, lights=None, materials=None, vertices=None, normals=None,
                 mat_indices=None):
        self.ctx = ctx
        hc = getattr(ctx, "helper", ctx)
        dev = ctx.torch_device
        self.width, self.height, self.spf = width, height, spf
        if vertices is None:
            n = C.c_uint64()
            hc.check(hc.lib.b200rt_generate_playground_scene(hc.h, hc.stream, rows, seed, 0, 0, 0, C.byref(n)), "playground_scene(size)")
            T = int(n.value)
            vertices = torch.empty((T * 3, 3), dtype=torch.float32, device=dev)
            normals = torch.empty((T * 3, 3), dtype=torch.float32, device=dev)
            mat_indices = torch.empty(T, dtype=torch.int32, device=dev)
            hc.check(hc.lib.b200rt_generate_playground_scene(hc.h, hc.stream, rows, seed, vertices.data_ptr(), normals.data_ptr(),
                                                             mat_indices.data_ptr(), C.byref(n)), "playground_scene")
        self.vertices, self.normals, self.mat_indices = vertices, normals, mat_indices
        self.num_triangles = mat_indices.numel()
        # TriangleGAS (triangle_gas.cpp:176-234): unindexed float3 vertices, OPTIX_GEOMETRY_FLAG_NONE, BUILD_FLAG_NONE, no compaction
        bi = ctx.triangle_input(vertices, num_sbt=1, flags=[0], vertex_stride=12)
        self.accel = ctx.build_accel([bi], compact=False)
        self.programs = ctx.prepare_programs("playground")
        self.camera_bytes = camera or playground_camera()
        self.lights_bytes = lights if lights is not None else playground_default_lights()
        self.materials = materials if materials is not None else playground_default_materials()
        self.d_camera = ctx.to_device(np.frombuffer(self.camera_bytes, np.uint8).copy())
        self.d_lights = ctx.to_device(np.frombuffer(self.lights_bytes, np.uint8).copy())
        self.d_materials = ctx.to_device(self.materials)
        self.film = torch.zeros((height, width, 3), dtype=torch.float32, device=dev)
        self.image = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev)
        # SBT (main.cpp:190-227): raygen, miss {bg .3,.1,.2 — never read by __miss__ms}, one empty hit-group record
        rec = np.zeros((3, 48), np.uint8)
        for k in range(3):
            rec[k, 0:32] = np.frombuffer(ctx.sbt_header(self.programs, k, 0), np.uint8)
        rec[1, 32:44] = np.frombuffer(struct.pack("<3f", 0.3, 0.1, 0.2), np.uint8)
        self.d_sbt = ctx.to_device(rec)
        self.sbt = L.ShaderBindingTable()
        self.sbt.raygenRecord = self.d_sbt.data_ptr()
        self.sbt.missRecordBase = self.d_sbt.data_ptr() + 48
        self.sbt.missRecordStrideInBytes, self.sbt.missRecordCount = 48, 1
        self.sbt.hitgroupRecordBase = self.d_sbt.data_ptr() + 96
        self.sbt.hitgroupRecordStrideInBytes, self.sbt.hitgroupRecordCount = 48, 1
        self.params = PGParams()
        p = self.params
        p.image_width, p.image_height, p.samples_per_frame = width, height, spf
        p.camera, p.dt, p.dirty = self.d_camera.data_ptr(), 0, 1
        p.image, p.film, p.tfactor, p.handle = self.image.data_ptr(), self.film.data_ptr(), 0.5, self.accel.handle
        p.normals, p.vertices, p.mat_indices, p.nmat_indices = normals.data_ptr(), vertices.data_ptr(), mat_indices.data_ptr(), self.num_triangles
        p.lights, p.nlights = self.d_lights.data_ptr(), len(self.lights_bytes) // 44
        p.materials, p.nmaterials = self.d_materials.data_ptr(), self.materials.shape[0]
        self.uploader = ParamsUploader(128, dev)
        self.d_params = self.uploader.d
        self.stats = L.PTStats()

    def launch_frame(self, dirty=False, collect_stats=False):
        """One iteration of TracerWindow::run's loop (tracer_window.cpp:88-105): dirty resets dt, frame_step(), Params upload, launch."""
        p = self.params
        p.dirty = 1 if dirty else 0
        if dirty:
            p.dt = 0
        p.dt += p.samples_per_frame
        self.uploader.upload(p)
        opts = L.PTOptions(0, int(collect_stats), C.pointer(self.stats), 0, 0)
        self.ctx.launch_playground(self.programs, self.d_params.data_ptr(), 128, self.sbt, self.width, self.height, opts)
        return self.stats if collect_stats else None


# ---- optixMeshViewer: SDK/optixMeshViewer/optixMeshViewer.cpp:190-308 + sutil::Scene::finalize (Scene.cpp:673-689,817-1212,1405-1433) -----
class WLaunchParams(C.Structure):  # whitted::LaunchParams (SDK/cuda/whitted.h:59-77), 128 bytes
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("subframe_index", C.c_uint32), ("accum_buffer", C.c_uint64),
                ("frame_buffer", C.c_uint64), ("max_depth", C.c_int32), ("scene_epsilon", C.c_float), ("eye", C.c_float * 3), ("U", C.c_float * 3),
                ("V", C.c_float * 3), ("W", C.c_float * 3), ("lights_data", C.c_uint64), ("lights_count", C.c_uint32), ("lights_stride", C.c_uint16),
                ("lights_elmt", C.c_uint16), ("miss_color", C.c_float * 3), ("handle", C.c_uint64)]


assert C.sizeof(WLaunchParams) == 128 and WLaunchParams.lights_data.offset == 88 and WLaunchParams.handle.offset == 120


def pack_texture(t, tex_objects):
    """MaterialData::Texture (40 B): texcoord, cudaTextureObject_t, offset, rotation (sin, cos), scale (Scene.cpp:214-265)."""
    if t is None or tex_objects.get(t["index"]) is None:
        return struct.pack("<iiQ6f", 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    rot = np.float32(t["rotation"])
    return struct.pack("<iiQ6f", int(t["texcoord"]), 0, tex_objects[t["index"]], float(t["offset"][0]), float(t["offset"][1]),
                       float(np.sin(rot, dtype=np.float32)), float(np.cos(rot, dtype=np.float32)), float(t["scale"][0]), float(t["scale"][1]))


def pack_material(m, tex_objects):
    """MaterialData (240 B, SDK/cuda/MaterialData.h:34-140; offsets pinned in tests/golden/kat.json "whitted_layout")."""
    b = bytearray(240)
    if m is None:
        m = {"base_color": [1, 1, 1, 1], "metallic": 1.0, "roughness": 1.0, "base_color_tex": None, "metallic_roughness_tex": None,
             "normal_tex": None, "emissive_tex": None, "emissive_factor": [0, 0, 0], "alpha_mode": 0, "alpha_cutoff": 0.0, "double_sided": False}
    b[8:48] = pack_texture(m["normal_tex"], tex_objects)
    b[48:56] = struct.pack("<if", m["alpha_mode"], float(m["alpha_cutoff"]) if m["alpha_mode"] == 1 else 0.0)
    b[56:68] = struct.pack("<3f", *[float(x) for x in m["emissive_factor"]])
    b[72:112] = pack_texture(m["emissive_tex"], tex_objects)
    b[112] = 1 if m["double_sided"] else 0
    b[128:152] = struct.pack("<6f", *[float(x) for x in m["base_color"]], float(m["metallic"]), float(m["roughness"]))
    b[152:192] = pack_texture(m["base_color_tex"], tex_objects)
    b[192:232] = pack_texture(m["metallic_roughness_tex"], tex_objects)
    return bytes(b)


class MeshViewer:
    """Mirror of optixMeshViewer's state for a scene dict as load_gltf returns it (textures optional: `images` entries may be None)."""

    def __init__(self, ctx, scene, width, height, textures=None):
        """textures: (texture objects, handles) made elsewhere — create_scene_textures_shared's entry for this context — instead of a
        private copy of every texture."""
        self.ctx, self.scene, self.width, self.height = ctx, scene, width, height
        dev = ctx.torch_device
        self.programs = ctx.prepare_programs("whitted")
        self.tex_objects, self._tex_handles = textures if textures is not None else create_scene_textures(ctx, scene)
        self.mesh_accels, mesh_records, self.keep = build_scene_meshes(ctx, scene, self.tex_objects)
        # createSBT (Scene.cpp:1405-1433): per INSTANCE, per primitive group: radiance record + occlusion record (same data)
        records, inst = [], []
        for i in scene["instances"]:
            inst.append((i["transform"][:3, :].reshape(12), len(records), self.mesh_accels[i["mesh"]]))
            for data in mesh_records[i["mesh"]]:
                records.append(ctx.sbt_header(self.programs, 2, 0) + data)
                records.append(ctx.sbt_header(self.programs, 2, 1) + data)
        self.d_hitgroup = ctx.to_device(np.frombuffer(b"".join(records), np.uint8).copy())
        self.ias = ctx.build_accel([ctx.instance_input(inst)], compact=False)
        miss = np.zeros((2, 32), np.uint8)
        for r in range(2):
            miss[r] = np.frombuffer(ctx.sbt_header(self.programs, 1, r), np.uint8)
        self.d_miss = ctx.to_device(miss)
        self.d_raygen = ctx.to_device(np.frombuffer(ctx.sbt_header(self.programs, 0, 0), np.uint8).copy())
        self.sbt = L.ShaderBindingTable()
        self.sbt.raygenRecord = self.d_raygen.data_ptr()
        self.sbt.missRecordBase, self.sbt.missRecordStrideInBytes, self.sbt.missRecordCount = self.d_miss.data_ptr(), 32, 2
        self.sbt.hitgroupRecordBase, self.sbt.hitgroupRecordStrideInBytes, self.sbt.hitgroupRecordCount = self.d_hitgroup.data_ptr(), 32 + 352, len(records)
        # scene AABB, camera (Scene.cpp:683-688,777-791), lights (optixMeshViewer.cpp:199-212)
        lo = np.min([i["world_aabb"][0] for i in scene["instances"]], axis=0).astype(np.float32)
        hi = np.max([i["world_aabb"][1] for i in scene["instances"]], axis=0).astype(np.float32)
        center = ((lo + hi) * np.float32(0.5)).astype(np.float32)
        ext = (hi - lo).astype(np.float32)
        max_ext = ext[int(np.argmax(ext))]
        if scene.get("cameras"):
            cam = scene["cameras"][0]
            eye, up, fov = cam["eye"], cam["up"], cam["fov_y"]
        else:
            eye, up, fov = (center + np.array([0.0, 0.0, 1.5 * max_ext], np.float32)).astype(np.float32), np.array([0, 1, 0], np.float32), 45.0
        self.eye = np.asarray(eye, np.float32)
        U, V, W = camera_uvw(self.eye, center, up, fov, np.float32(width) / np.float32(height))
        lights = b"".join([
            struct.pack("<i3ff3fi", 0, 1.0, 1.0, 0.8, 5.0, *[float(x) for x in (center + max_ext)], 2),
            struct.pack("<i3ff3fi", 0, 0.8, 0.8, 1.0, 3.0, *[float(x) for x in (center + np.array([-max_ext, np.float32(0.5) * max_ext, np.float32(-0.5) * max_ext], np.float32))], 2)])
        self.d_lights = ctx.to_device(np.frombuffer(lights, np.uint8).copy())
        self.accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
        self.frame = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev)
        self.params = WLaunchParams(width, height, 0, self.accum.data_ptr(), self.frame.data_ptr(), 0, 0.0, _f3(self.eye), _f3(U), _f3(V), _f3(W),
                                    self.d_lights.data_ptr(), 2, 0, 0, _f3([0.1, 0.1, 0.1]), self.ias.handle)
        self.uploader = ParamsUploader(128, dev)
        self.d_params = self.uploader.d

    def launch_subframe(self, subframe_index=None):
        """launchSubframe (optixMeshViewer.cpp:283-308)."""
        if subframe_index is not None:
            self.params.subframe_index = subframe_index
        # the reference's cudaMemcpyAsync(d_params, &params, ...) (optixMeshViewer.cpp:289-293) from a ring of pinned staging blocks: the host
        # may change `params` for the next subframe while this launch is still queued, and nothing here waits for the device (a torch copy
        # from pageable memory synchronises the stream)
        self.uploader.upload(self.params)
        self.ctx.launch_whitted(self.programs, self.d_params.data_ptr(), 128, self.sbt, self.width, self.height)

    def close(self):
        for hc, tex, arr in self._tex_handles:
            hc.lib.b200rt_texture_destroy(hc.h, tex, arr)
        self._tex_handles = []

# ---- one result buffer for all GPUs (optixMultiGPU / optixNVLink) ------------------------------------------------------------------------
u64_t = C.c_uint64


class SharedResultBuffer:
    """The reference's optixMultiGPU lets every device write its pixels into ONE result buffer (zero-copy host memory,
    SDK/sutil/CUDAOutputBuffer.h:203-216; device memory reached by peer access in optixNVLink.cpp:1975-1992).  With one process per GPU the
    buffer lives in the owner rank's HBM (b200rt_shared_buffer_create), its 64-byte handle goes to the other ranks through
    torch.distributed, and they open it with their own context: `ptr` is what each rank puts into Params::result_buffer, and the final
    stores of its launch travel over NVLink — no gather, 4 bytes per pixel.  `tensor` (owner only) views the buffer as (h, w, 4) uint8.
    Collective: construct on every rank."""

    def __init__(self, ctx, height, width, rank, owner=0):
        import torch.distributed as dist
        self.ctx, self.owner, self.rank, self.shape = ctx, owner, rank, (height, width, 4)
        ptr, payload, err = u64_t(), [None], None
        if rank == owner:
            handle = C.create_string_buffer(64)
            try:
                ctx.check(ctx.lib.b200rt_shared_buffer_create(ctx.h, height * width * 4, C.byref(ptr), handle), "shared_buffer_create")
                payload[0] = handle.raw
            except B200RTError as e:  # the others are waiting in the broadcast: tell them
                err = e
        dist.broadcast_object_list(payload, src=owner)
        if payload[0] is None:
            raise err or B200RTError("the owner rank could not create the shared result buffer")
        if rank != owner:
            ctx.check(ctx.lib.b200rt_shared_buffer_open(ctx.h, payload[0], C.byref(ptr)), "shared_buffer_open")
        self.ptr = ptr.value
        self.tensor = None
        if rank == owner:
            self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "|u1", "data": (self.ptr, False), "version": 2}
            self.tensor = torch.as_tensor(self, device=ctx.torch_device)

    def close(self):
        if self.ptr:
            fn = self.ctx.lib.b200rt_shared_buffer_destroy if self.rank == self.owner else self.ctx.lib.b200rt_shared_buffer_close
            fn(self.ctx.h, self.ptr)
            self.ptr = 0


# ---- output and model ingest either side of the path ---------------------------------------------------------
from .scene_io import load_exr, load_nbt, read_nbt, save_exr, save_nbt  # noqa: E402,F401  (NBT models, EXR output: no GPU involved)


def _to_srgb_u8(f):
    """toSRGB + the 256-scale quantisation of sutil::saveImage's float branches (SDK/sutil/sutil.cpp:585-618, SDK/cuda/helpers.h:36-48)."""
    f = np.asarray(f, np.float32)
    inv = np.float32(1.0 / 2.4)
    with np.errstate(invalid="ignore", divide="ignore"):
        srgb = np.where(f < np.float32(0.0031308), np.float32(12.92) * f, np.float32(1.055) * np.power(f, inv, dtype=np.float32) - np.float32(0.055))
    v = (np.float32(256.0) * srgb.astype(np.float32)).astype(np.int64)
    return np.clip(v, 0, 255).astype(np.uint8)


def save_image(path, image, disable_srgb_conversion=False):
    """sutil::saveImage (SDK/sutil/sutil.cpp:542-709) for .ppm / .png / .exr (scene_io.save_exr): `image` is a (h, w, 4) uint8 frame (written as is), or a (h, w, 3|4)
    float32 buffer (sRGB-converted unless disabled, 256-scaled, clamped).  Rows are flipped: the launch index (0, 0) is the bottom-left
    pixel of the picture.  PPM is the binary P6 the reference writes; PNG keeps the alpha of a uchar4 frame like stbi_write_png(…, 4, …)."""
    if hasattr(image, "cpu"):
        image = image.cpu().numpy()
    a = np.asarray(image)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError("sutil::saveImage(): Unrecognized image buffer pixel format.")
    path = str(path)
    if len(path) < 5:
        raise ValueError("sutil::saveImage(): Failed to determine filename extension")
    ext = path[-3:].lower()
    if ext == "exr":
        # SDK/sutil/sutil.cpp:660-702: float3 / float4 buffers go to tinyexr as they are (linear, fp16, buffer row order)
        if a.dtype == np.uint8:
            raise ValueError("sutil::saveImage(): saving of uchar4 images to EXR not implemented yet")
        if a.dtype != np.float32:
            raise ValueError("sutil::saveImage: Unrecognized image buffer pixel format.")
        save_exr(path, a)
        return
    if a.dtype == np.uint8:
        if a.shape[2] != 4:
            raise ValueError("sutil::saveImage(): Unrecognized image buffer pixel format.")
        pix = a
    elif a.dtype == np.float32:
        rgb = a[..., :3] if disable_srgb_conversion else None
        q = np.clip((np.float32(256.0) * rgb).astype(np.int64), 0, 255).astype(np.uint8) if disable_srgb_conversion else _to_srgb_u8(a[..., :3])
        pix = np.concatenate([q, np.full(q.shape[:2] + (1,), 255, np.uint8)], axis=-1)
    else:
        raise ValueError("sutil::saveImage(): Unrecognized image buffer pixel format.")
    pix = pix[::-1]  # flipped vertically as it is written
    if ext == "ppm":
        h, w = pix.shape[:2]
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h))
            f.write(np.ascontiguousarray(pix[..., :3]).tobytes())
    elif ext == "png":
        from PIL import Image
        Image.fromarray(np.ascontiguousarray(pix if a.dtype == np.uint8 else pix[..., :3])).save(path)
    else:
        raise ValueError(f"sutil::saveImage(): Invalid extension '{ext}'")


def load_obj_like_assimp(path, floor_material=26):
    """imgui_test's load_assimp (SDK/imgui_test/triangle_gas.cpp:78-168) for a Wavefront OBJ: every face becomes three unindexed vertices
    (faces with more corners are fanned, as assimp's triangulation of convex polygons does), per-vertex normals where the file has them,
    material index 0, and the sample's floor — 20 x 20 cells of 0.1 at the lowest y of the model, two triangles each, normal (0, 1, 0),
    material 26 — appended.  Returns (vertices (3T, 3) f32, normals (3T', 3) f32, mat_indices (T,) i32): what TriangleGAS keeps and
    host.Playground takes."""
    v, vn, verts, norms = [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                vn.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                corners = []
                for c in t[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else None
                    corners.append((vi - 1 if vi > 0 else len(v) + vi, None if ni is None else (ni - 1 if ni > 0 else len(vn) + ni)))
                for k in range(1, len(corners) - 1):
                    for vi, ni in (corners[0], corners[k], corners[k + 1]):
                        verts.append(v[vi])
                        if ni is not None:
                            norms.append(vn[ni])
    vertices = np.asarray(verts, np.float32).reshape(-1, 3)
    normals = np.asarray(norms, np.float32).reshape(-1, 3)
    mats = [0] * (vertices.shape[0] // 3)
    floor = np.float32(vertices[:, 1].min()) if vertices.size else np.float32(1.0e12)
    fv = []
    for i in range(-10, 10):
        for j in range(-10, 10):
            x0, x1 = np.float32(i) * np.float32(0.1), np.float32(i + 1) * np.float32(0.1)
            z0, z1 = np.float32(j) * np.float32(0.1), np.float32(j + 1) * np.float32(0.1)
            fv += [[x0, floor, z0], [x0, floor, z1], [x1, floor, z0], [x1, floor, z0], [x0, floor, z1], [x1, floor, z1]]
            mats += [floor_material, floor_material]
    vertices = np.concatenate([vertices, np.asarray(fv, np.float32)], axis=0)
    normals = np.concatenate([normals, np.tile(np.asarray([[0.0, 1.0, 0.0]], np.float32), (len(fv), 1))], axis=0)
    return vertices, normals, np.asarray(mats, np.int32)
