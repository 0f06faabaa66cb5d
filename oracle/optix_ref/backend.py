"""TEST INFRASTRUCTURE — the reference's OptiX path as a second oracle and as the GPU baseline.

`OptixContext` is a drop-in for `optix_raytracer_b200.host.Context` in the host mirrors (`PathTracer`,
`Raycaster`): same build inputs, same Params bytes, same SBT payloads, but the acceleration structure is
built by `optixAccelBuild` and the launch is `optixLaunch` of the REFERENCE'S OWN device programs
(oracle/_ref/*.ptx, compiled by oracle/Makefile from /root/reference/SDK/optix*/...cu where they lie) on
the closed runtime `libnvoptix.so.1`, through the harness oracle/optix_ref/optix_harness.cpp.

Only tests/ and tools/optix_compare.py (the same-GPU comparison report) import this.  libnvoptix.so.1 is present
on the GPU box only (probed: SURVEY.md section 8c / DESIGN.md section 6); `available()` says whether it
can be used in this process.
"""
import ctypes as C
import os
import pathlib

import numpy as np
import torch

from optix_raytracer_b200 import _lib as L
from optix_raytracer_b200 import host

_REF = pathlib.Path(__file__).resolve().parents[1] / "_ref"
_LIB = None
_WHY = None
_SHIM = None


def shim_active():
    """True when this process drives the harness through optix_raytracer_b200/optix_shim/libnvoptix.so.1 (B200RT_OPTIX_SHIM=1)."""
    return _SHIM is not None

# OptixPayloadSemantics (reference include/optix_types.h:2003-2029)
_CALLER_R, _CALLER_W, _CALLER_RW = 1 << 0, 2 << 0, 3 << 0
_CH_R, _CH_W, _CH_RW = 1 << 2, 2 << 2, 3 << 2
_MS_W = 2 << 4
# radiancePayloadSemantics (reference SDK/optixPathTracer/optixPathTracer.h:51-80)
RADIANCE_PAYLOAD_SEMANTICS = ([_CALLER_RW | _CH_RW] * 5 + [_CALLER_R | _CH_W | _MS_W] * 6 + [_CALLER_R | _CH_W] * 6
                              + [_CALLER_R | _CH_W | _MS_W])
GRAPH_SINGLE_GAS, GRAPH_SINGLE_LEVEL_INSTANCING = 1 << 0, 1 << 1  # OptixTraversableGraphFlags

PROGRAMS = {
    # kind: (ptx, raygen, miss csv, closest-hit csv, any-hit csv, numPayloadValues, typed payload, graph flags, maxTraceDepth, maxTraversableDepth)
    "pathtracer": ("optixPathTracer.ptx", "__raygen__rg", "__miss__radiance", "__closesthit__radiance", "-", 0, RADIANCE_PAYLOAD_SEMANTICS,
                   GRAPH_SINGLE_GAS, 2, 1),                                      # optixPathTracer.cpp:686-826
    "multigpu": ("optixMultiGPU.ptx", "__raygen__rg", "__miss__radiance,-", "__closesthit__radiance,__closesthit__occlusion", "-,-", 2, None,
                 GRAPH_SINGLE_GAS, 2, 1),                                        # optixMultiGPU.cpp:786-950
    "raycast": ("optixRaycasting.ptx", "__raygen__from_buffer", "__miss__buffer_miss", "__closesthit__buffer_hit", "__anyhit__texture_mask", 4,
                None, GRAPH_SINGLE_LEVEL_INSTANCING, 1, 2),                      # optixRaycasting.cpp:94-196
    "whitted": ("whitted.ptx", "__raygen__pinhole", "__miss__constant_radiance,__miss__occlusion", "__closesthit__radiance,-",
                "__anyhit__radiance,__anyhit__occlusion", 4, None, GRAPH_SINGLE_LEVEL_INSTANCING, 8, 2),  # Scene.cpp:1215-1403
    "playground": ("optixTriangle.ptx", "__raygen__rg", "__miss__ms", "__closesthit__ch", "-", 3, None, GRAPH_SINGLE_GAS, 2, 1),  # imgui_test/main.cpp:71-188
    # (the sample links with maxTraceDepth 1 although its closest-hit program traverses again; 2 here only enlarges the stack)
    "query_gas": ("query_programs.ptx", "__raygen__query", "__miss__query", "__closesthit__query", "-", 5, None, GRAPH_SINGLE_GAS, 1, 1),
    "query_ias": ("query_programs.ptx", "__raygen__query", "__miss__query", "__closesthit__query", "-", 5, None, GRAPH_SINGLE_LEVEL_INSTANCING, 1, 2),
}


def _load():
    global _LIB, _WHY
    if _LIB is not None or _WHY is not None:
        return _LIB
    if os.environ.get("B200RT_OPTIX_SHIM") == "1":
        # Run the harness — an OptiX host program like the samples — on the product's own optixQueryFunctionTable instead of the driver's:
        # a library with the soname libnvoptix.so.1 that is already loaded is what optixInit()'s dlopen("libnvoptix.so.1") returns.
        global _SHIM
        shim = pathlib.Path(L.__file__).resolve().parent / "optix_shim" / "libnvoptix.so.1"
        _SHIM = C.CDLL(str(shim), mode=C.RTLD_GLOBAL)
    so = _REF / "liboptixref.so"
    if not so.exists():
        _WHY = f"{so} not built (make -C oracle optix needs /root/reference)"
        return None
    try:
        lib = C.CDLL(str(so))
    except OSError as e:
        _WHY = f"cannot load {so}: {e}"
        return None
    u32, i32, u64, vp, sz = C.c_uint32, C.c_int32, C.c_uint64, C.c_void_p, C.c_size_t
    lib.oref_log.restype = C.c_char_p
    lib.oref_init.argtypes = [i32, i32]
    lib.oref_rtcore_version.restype = u32
    lib.oref_accel_compute_memory_usage.argtypes = [vp, vp, u32, vp]
    lib.oref_accel_build.argtypes = [vp, vp, vp, u32, u64, sz, u64, sz, C.POINTER(u64), vp, u32]
    lib.oref_accel_compact.argtypes = [vp, u64, u64, sz, C.POINTER(u64)]
    lib.oref_pipeline_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, i32, C.POINTER(u32), i32, u32, u32, u32,
                                         C.POINTER(i32)]
    lib.oref_pack_header.argtypes = [i32, i32, i32, vp]
    lib.oref_launch.argtypes = [i32, vp, u64, sz, vp, u32, u32, u32]
    _LIB = lib
    return lib


def available(device=0):
    """(True, '') when OptiX can be initialised on `device` in this process, else (False, reason)."""
    lib = _load()
    if lib is None:
        return False, _WHY
    if not torch.cuda.is_available():
        return False, "no CUDA device"
    rc = lib.oref_init(device, 0)
    if rc:
        return False, f"optixInit / optixDeviceContextCreate failed ({rc}): {lib.oref_log().decode(errors='replace')[-300:]}"
    return True, ""


class OptixError(RuntimeError):
    pass


class OptixContext(host.Context):
    """host.Context with OptiX behind it."""

    def __init__(self, device=0, log_level=0):
        ok, why = available(device)
        if not ok:
            raise OptixError(why)
        self.lib = None           # no b200rt calls from this context
        self.olib = _load()
        self.h = None
        self.device = device
        self.torch_device = torch.device("cuda", device)
        self._pipelines = {}
        self.helper = host.Context(device)  # plain-CUDA helper kernels (ray generation, shading of hit buffers) are not OptiX calls

    def close(self):
        pass

    def ocheck(self, rc, what):
        if rc:
            raise OptixError(f"{what}: {rc}\n{self.olib.oref_log().decode(errors='replace')[-2000:]}")

    @property
    def kernel_launches(self):
        return 0

    def _accel_memory_usage(self, opts, arr, n, sizes):
        self.ocheck(self.olib.oref_accel_compute_memory_usage(C.byref(opts), arr, n, C.byref(sizes)), "optixAccelComputeMemoryUsage")

    def _accel_build(self, opts, arr, n, temp, temp_bytes, out, out_bytes, handle, emit):
        self.ocheck(self.olib.oref_accel_build(self.stream, C.byref(opts), arr, n, temp, temp_bytes, out, out_bytes, C.byref(handle),
                                               C.byref(emit) if emit is not None else None, 1 if emit is not None else 0), "optixAccelBuild")

    def _accel_compact(self, handle, out, out_bytes, new_handle):
        self.ocheck(self.olib.oref_accel_compact(self.stream, handle, out, out_bytes, C.byref(new_handle)), "optixAccelCompact")

    def prepare_programs(self, kind):
        if kind not in self._pipelines:
            ptx, rg, ms, ch, ah, npay, sem, graph, depth, tdepth = PROGRAMS[kind]
            pid = C.c_int32(-1)
            sem_arr = (C.c_uint32 * len(sem))(*sem) if sem else None
            self.ocheck(self.olib.oref_pipeline_create(str(_REF / ptx).encode(), rg.encode(), ms.encode(), ch.encode(), ah.encode(), npay, sem_arr,
                                                       len(sem) if sem else 0, graph, depth, tdepth, C.byref(pid)), f"pipeline {kind}")
            self._pipelines[kind] = pid.value
        return self._pipelines[kind]

    def sbt_header(self, programs, kind, index):
        buf = (C.c_uint8 * 32)()
        self.ocheck(self.olib.oref_pack_header(programs, kind, index, buf), "optixSbtRecordPackHeader")
        return bytes(buf)

    def launch_pathtracer(self, programs, d_params, params_size, sbt, width, height, opts):
        self.ocheck(self.olib.oref_launch(programs, self.stream, d_params, params_size, C.byref(sbt), width, height, 1), "optixLaunch")

    def launch_multigpu(self, programs, d_params, params_size, sbt, num_samples, opts):
        self.ocheck(self.olib.oref_launch(programs, self.stream, d_params, params_size, C.byref(sbt), num_samples, 1, 1), "optixLaunch")

    def launch_whitted(self, programs, d_params, params_size, sbt, width, height):
        self.ocheck(self.olib.oref_launch(programs, self.stream, d_params, params_size, C.byref(sbt), width, height, 1), "optixLaunch")

    def launch_playground(self, programs, d_params, params_size, sbt, width, height, opts):
        self.ocheck(self.olib.oref_launch(programs, self.stream, d_params, params_size, C.byref(sbt), width, height, 1), "optixLaunch")

    def launch_raycast(self, programs, d_params, sbt, width, height, ext):
        self.ocheck(self.olib.oref_launch(programs, self.stream, d_params, 24, C.byref(sbt), width, height, 1), "optixLaunch")

    # ---- ray-buffer queries with OptiX's built-in triangle test (query_programs.cu) ---------------------------
    def _query(self, accel, rays, ray_flags, any_hit, is_ias):
        pid = self.prepare_programs("query_ias" if is_ias else "query_gas")
        n = rays.shape[0]
        out = torch.empty(n if any_hit else (n, 5), dtype=torch.int32, device=self.torch_device)
        params = np.zeros(40, np.uint8)
        params[0:8] = np.frombuffer(np.uint64(accel.handle).tobytes(), np.uint8)
        params[8:16] = np.frombuffer(np.uint64(rays.data_ptr()).tobytes(), np.uint8)
        params[16:24] = np.frombuffer(np.uint64(out.data_ptr()).tobytes(), np.uint8)
        params[24:28] = np.frombuffer(np.uint32(ray_flags).tobytes(), np.uint8)
        params[28:32] = np.frombuffer(np.uint32(1 if any_hit else 0).tobytes(), np.uint8)
        params[32:40] = np.frombuffer(np.uint64(n).tobytes(), np.uint8)
        d_params = self.to_device(params)
        nhit = 256  # trace stride 0: the hit record index is the instance's sbtOffset; cover any offset the test scenes use
        rec = np.zeros((2 + nhit, 32), np.uint8)
        rec[0] = np.frombuffer(self.sbt_header(pid, 0, 0), np.uint8)
        rec[1] = np.frombuffer(self.sbt_header(pid, 1, 0), np.uint8)
        rec[2:] = np.frombuffer(self.sbt_header(pid, 2, 0), np.uint8)
        d_rec = self.to_device(rec)
        sbt = L.ShaderBindingTable()
        sbt.raygenRecord = d_rec.data_ptr()
        sbt.missRecordBase = d_rec.data_ptr() + 32
        sbt.missRecordStrideInBytes, sbt.missRecordCount = 32, 1
        sbt.hitgroupRecordBase = d_rec.data_ptr() + 64
        sbt.hitgroupRecordStrideInBytes, sbt.hitgroupRecordCount = 32, nhit
        w = min(n, 1 << 15)
        h = (n + w - 1) // w
        self.ocheck(self.olib.oref_launch(pid, self.stream, d_params.data_ptr(), 40, C.byref(sbt), w, h, 1), "optixLaunch(query)")
        torch.cuda.synchronize(self.torch_device)
        return out

    def trace_closest(self, accel, rays, ray_flags=0, is_ias=False):
        return self._query(accel, rays, ray_flags, False, is_ias)

    def trace_any(self, accel, rays, ray_flags=0, is_ias=False):
        return self._query(accel, rays, ray_flags, True, is_ias)
