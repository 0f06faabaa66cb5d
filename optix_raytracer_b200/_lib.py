"""ctypes binding of libb200rt.so (include/b200rt.h).  The library is the product: if it is missing
or cannot be loaded this module raises — there is no fallback path."""
import ctypes as C
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
import os

# B200RT_LIB_PATH: development override used for A/B runs of kernel variants (tools/ab_variants.sh); the product path is the default
LIB_PATH = pathlib.Path(os.environ.get("B200RT_LIB_PATH", _HERE / "libb200rt.so"))

u32, i32, u64, f32, vp, sz = C.c_uint32, C.c_int32, C.c_uint64, C.c_float, C.c_void_p, C.c_size_t

# ---- constants (include/b200rt.h) -------------------------------------------------------------
BUILD_INPUT_TYPE_TRIANGLES = 0x2141
BUILD_INPUT_TYPE_INSTANCES = 0x2143
VERTEX_FORMAT_FLOAT3 = 0x2121
INDICES_FORMAT_NONE = 0
INDICES_FORMAT_UNSIGNED_SHORT3 = 0x2102
INDICES_FORMAT_UNSIGNED_INT3 = 0x2103
TRANSFORM_FORMAT_MATRIX_FLOAT12 = 0x21E1
BUILD_OPERATION_BUILD = 0x2161
PROPERTY_TYPE_COMPACTED_SIZE = 0x2181
BUILD_FLAG_ALLOW_COMPACTION = 1 << 1
GEOMETRY_FLAG_DISABLE_ANYHIT = 1
GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING = 4
RAY_FLAG_TERMINATE_ON_FIRST_HIT = 1 << 2
RAY_FLAG_CULL_BACK_FACING_TRIANGLES = 1 << 4
RAY_FLAG_CULL_FRONT_FACING_TRIANGLES = 1 << 5
SBT_RECORD_HEADER_SIZE = 32
PT_STATS_SEGMENTS, PT_STATS_TIMING, PT_STATS_TRAVERSAL = 1, 2, 4


class TriangleArray(C.Structure):
    _fields_ = [("vertexBuffers", C.POINTER(u64)), ("numVertices", u32), ("vertexFormat", u32),
                ("vertexStrideInBytes", u32), ("indexBuffer", u64), ("numIndexTriplets", u32), ("indexFormat", u32),
                ("indexStrideInBytes", u32), ("preTransform", u64), ("flags", C.POINTER(u32)), ("numSbtRecords", u32),
                ("sbtIndexOffsetBuffer", u64), ("sbtIndexOffsetSizeInBytes", u32), ("sbtIndexOffsetStrideInBytes", u32),
                ("primitiveIndexOffset", u32), ("transformFormat", u32), ("opaque_micromaps", C.c_char * 144)]


class InstanceArray(C.Structure):
    _fields_ = [("instances", u64), ("numInstances", u32), ("instanceStride", u32)]


class _BuildInputUnion(C.Union):
    _fields_ = [("triangleArray", TriangleArray), ("instanceArray", InstanceArray), ("pad", C.c_char * 1024)]


class BuildInput(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("type", u32), ("u", _BuildInputUnion)]


class Instance(C.Structure):
    _fields_ = [("transform", f32 * 12), ("instanceId", u32), ("sbtOffset", u32), ("visibilityMask", u32), ("flags", u32),
                ("traversableHandle", u64), ("pad", u32 * 2)]


class MotionOptions(C.Structure):
    _fields_ = [("numKeys", C.c_uint16), ("flags", C.c_uint16), ("timeBegin", f32), ("timeEnd", f32)]


class AccelBuildOptions(C.Structure):
    _fields_ = [("buildFlags", u32), ("operation", u32), ("motionOptions", MotionOptions)]


class AccelBufferSizes(C.Structure):
    _fields_ = [("outputSizeInBytes", sz), ("tempSizeInBytes", sz), ("tempUpdateSizeInBytes", sz)]


class AccelEmitDesc(C.Structure):
    _fields_ = [("result", u64), ("type", u32)]


class ShaderBindingTable(C.Structure):
    _fields_ = [("raygenRecord", u64), ("exceptionRecord", u64), ("missRecordBase", u64), ("missRecordStrideInBytes", u32),
                ("missRecordCount", u32), ("hitgroupRecordBase", u64), ("hitgroupRecordStrideInBytes", u32),
                ("hitgroupRecordCount", u32), ("callablesRecordBase", u64), ("callablesRecordStrideInBytes", u32),
                ("callablesRecordCount", u32)]


class AccelInfo(C.Structure):
    _fields_ = [("kind", u32), ("num_triangles", u32), ("num_nodes", u32), ("num_instances", u32), ("total_bytes", u64),
                ("bounds", f32 * 6), ("depth", u32), ("reserved", u32)]


class PTStats(C.Structure):
    _fields_ = [("radiance_segments", u64), ("shadow_segments", u64), ("iterations", u32), ("kernel_launches", u32),
                ("nodes_fetched", u64), ("tris_tested", u64), ("trace_ms", f32), ("shade_ms", f32), ("trace_launches", u32),
                ("reserved", u32)]


class PTOptions(C.Structure):
    _fields_ = [("sample_groups", u32), ("collect_stats", u32), ("stats", C.POINTER(PTStats)), ("ray_sort", u32), ("reserved", u32)]


assert C.sizeof(BuildInput) == 1032 and C.sizeof(TriangleArray) == 240 and C.sizeof(Instance) == 80
assert C.sizeof(AccelBuildOptions) == 20 and C.sizeof(ShaderBindingTable) == 64

LOG_CB = C.CFUNCTYPE(None, C.c_uint, C.c_char_p, C.c_char_p, vp)

# name -> (restype, argtypes); every symbol include/b200rt.h declares
SYMBOLS = {
    "b200rt_context_create": (i32, [i32, LOG_CB, vp, i32, C.POINTER(vp)]),
    "b200rt_context_destroy": (i32, [vp]),
    "b200rt_error_string": (C.c_char_p, [i32]),
    "b200rt_error_name": (C.c_char_p, [i32]),
    "b200rt_last_error_message": (C.c_char_p, [vp]),
    "b200rt_version": (C.c_char_p, []),
    "b200rt_context_kernel_launches": (u64, [vp]),
    "b200rt_enable_peer_access": (i32, [vp, i32]),
    "b200rt_shared_buffer_create": (i32, [vp, C.c_size_t, C.POINTER(u64), C.c_char_p]),
    "b200rt_shared_buffer_open": (i32, [vp, C.c_char_p, C.POINTER(u64)]),
    "b200rt_shared_buffer_close": (i32, [vp, u64]),
    "b200rt_shared_buffer_destroy": (i32, [vp, u64]),
    "b200rt_accel_compute_memory_usage": (i32, [vp, C.POINTER(AccelBuildOptions), C.POINTER(BuildInput), u32,
                                                C.POINTER(AccelBufferSizes)]),
    "b200rt_accel_build": (i32, [vp, vp, C.POINTER(AccelBuildOptions), C.POINTER(BuildInput), u32, u64, sz, u64, sz,
                                 C.POINTER(u64), C.POINTER(AccelEmitDesc), u32]),
    "b200rt_accel_compact": (i32, [vp, vp, u64, u64, sz, C.POINTER(u64)]),
    "b200rt_accel_emit_property": (i32, [vp, vp, u64, vp, u32]),
    "b200rt_accel_get_info": (i32, [vp, u64, C.POINTER(AccelInfo)]),
    "b200rt_launch_pathtracer": (i32, [vp, vp, u64, C.POINTER(ShaderBindingTable), u32, u32, C.POINTER(PTOptions)]),
    "b200rt_launch_multigpu": (i32, [vp, vp, u64, C.POINTER(ShaderBindingTable), u32, C.POINTER(PTOptions)]),
    "b200rt_fill_samples": (i32, [vp, vp, i32, i32, i32, i32, u64, i32]),
    "b200rt_deinterleave": (i32, [vp, vp, u64, i32, i32, i32, i32, u64, u64]),
    "b200rt_launch_raycast": (i32, [vp, vp, u64, C.POINTER(ShaderBindingTable), u32, u32, u64]),
    "b200rt_create_rays_ortho": (i32, [vp, vp, u64, i32, i32, C.POINTER(f32), C.POINTER(f32), f32]),
    "b200rt_translate_rays": (i32, [vp, vp, u64, i32, C.POINTER(f32)]),
    "b200rt_shade_hits": (i32, [vp, vp, u64, i32, u64]),
    "b200rt_trace_closest": (i32, [vp, vp, u64, u64, u64, u32, u64]),
    "b200rt_trace_any": (i32, [vp, vp, u64, u64, u64, u32, u64]),
    "b200rt_trace_stats": (i32, [vp, vp, u64, u64, u64, C.POINTER(u64), C.POINTER(u64)]),
    "b200rt_camera_uvw": (None, [C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), f32, f32, C.POINTER(f32), C.POINTER(f32),
                                 C.POINTER(f32)]),
    "b200rt_playground_camera": (None, [C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), f32, f32, f32, i32, vp]),
    "b200rt_launch_whitted": (i32, [vp, vp, u64, C.POINTER(ShaderBindingTable), u32, u32]),
    "b200rt_texture_create": (i32, [vp, i32, i32, vp, i32, i32, i32, C.POINTER(u64), C.POINTER(u64)]),
    "b200rt_texture_destroy": (i32, [vp, u64, u64]),
    "b200rt_texture_view": (i32, [vp, u64, i32, i32, i32, C.POINTER(u64)]),
    "b200rt_launch_playground": (i32, [vp, vp, u64, u32, u32, C.POINTER(PTOptions)]),
    "b200rt_generate_playground_scene": (i32, [vp, vp, u32, u32, u64, u64, u64, C.POINTER(u64)]),
    "b200rt_triangle_flag_word": (u32, [u32, u32]),
    "b200rt_wd_num_samples": (i32, [i32, i32, i32]),
    "b200rt_wd_sample_pixel": (None, [i32, i32, i32, i32, i32, C.POINTER(i32)]),
    "b200rt_generate_synthetic_mesh": (i32, [vp, vp, u64, u32, u64, u64, C.POINTER(f32)]),
}

_lib = None


def load():
    """Load libb200rt.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C optix_raytracer_b200/csrc` — b200rt has no CPU fallback")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
