/*
 * b200rt.h — C ABI of libb200rt.so, the B200-native (sm_100a) replacement for the OptiX launch
 * behind the optixPathTracer / optixMultiGPU / optixRaycasting samples of awegsche/OptiX_Raytracer.
 *
 * Every entry point replaces one call the reference makes through the OptiX function table
 * (reference include/optix_stubs.h:198-229, include/optix_function_table.h:46-343); the reference
 * call site is cited at each declaration.  Conventions follow the reference boundary:
 *   - plain pointers and sizes only; device addresses are passed as uint64_t (CUdeviceptr);
 *   - every function returns an int that is an OptixResult-compatible code (0 = success,
 *     7001 invalid value, 7003 invalid operation, 7050 launch failure, 7800 not supported,
 *     7900 CUDA error); no exception crosses the ABI (reference: OPTIX_CHECK, SDK/sutil/Exception.h:82-112);
 *   - the caller owns every device buffer it passes in (vertices, temp, output, SBT, params, frame /
 *     accum buffers); the library owns only its context and the per-context wavefront workspace;
 *   - launches and builds are asynchronous on the given CUDA stream unless stated otherwise; the
 *     caller synchronises (reference: CUDA_SYNC_CHECK after optixLaunch).
 * There is no CPU path: every compute entry point fails with 7900 if no CUDA device is usable.
 *
 * Struct layouts marked "layout == Optix..." are field-for-field identical to the OptiX 8.0 type
 * named (sizes asserted in csrc/api.cpp), so a reference host can pass its own structs by cast.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_SUCCESS 0
#define B200RT_ERROR_INVALID_VALUE 7001
#define B200RT_ERROR_HOST_OUT_OF_MEMORY 7002
#define B200RT_ERROR_INVALID_OPERATION 7003
#define B200RT_ERROR_LAUNCH_FAILURE 7050
#define B200RT_ERROR_INVALID_DEVICE_CONTEXT 7051
#define B200RT_ERROR_NOT_SUPPORTED 7800
#define B200RT_ERROR_CUDA_ERROR 7900

/* constants of reference include/optix_types.h:79-88 */
#define B200RT_SBT_RECORD_HEADER_SIZE 32
#define B200RT_SBT_RECORD_ALIGNMENT 16
#define B200RT_ACCEL_BUFFER_BYTE_ALIGNMENT 128
#define B200RT_INSTANCE_BYTE_ALIGNMENT 16

/* enum values of reference include/optix_types.h:345-368,1014-1024,1293-1295,1376-1379 */
#define B200RT_BUILD_INPUT_TYPE_TRIANGLES 0x2141
#define B200RT_BUILD_INPUT_TYPE_INSTANCES 0x2143
#define B200RT_VERTEX_FORMAT_FLOAT3 0x2121
#define B200RT_INDICES_FORMAT_NONE 0
#define B200RT_INDICES_FORMAT_UNSIGNED_SHORT3 0x2102
#define B200RT_INDICES_FORMAT_UNSIGNED_INT3 0x2103
#define B200RT_TRANSFORM_FORMAT_NONE 0
#define B200RT_TRANSFORM_FORMAT_MATRIX_FLOAT12 0x21E1
#define B200RT_BUILD_OPERATION_BUILD 0x2161
#define B200RT_PROPERTY_TYPE_COMPACTED_SIZE 0x2181
#define B200RT_PROPERTY_TYPE_AABBS 0x2182
#define B200RT_BUILD_FLAG_ALLOW_COMPACTION (1u << 1)
/* ray flags, reference include/optix_types.h:1794-1840 (same values).  Precedence of the any-hit state of a triangle, as documented
 * there: ray flags (DISABLE / ENFORCE_ANYHIT) over instance flags over the geometry flag of its SBT record.  OptiX declares some
 * combinations mutually exclusive (DISABLE / ENFORCE_ANYHIT with each other and with the CULL_*_ANYHIT flags; the two face-cull flags
 * with each other); this library gives them the obvious meaning instead of leaving them undefined: DISABLE wins over ENFORCE, the
 * CULL_*_ANYHIT flags see the any-hit state after the ray's override, and both face-cull flags together cull every triangle that takes
 * part in face culling. */
#define B200RT_RAY_FLAG_NONE 0u
#define B200RT_RAY_FLAG_DISABLE_ANYHIT (1u << 0)
#define B200RT_RAY_FLAG_ENFORCE_ANYHIT (1u << 1)
#define B200RT_RAY_FLAG_TERMINATE_ON_FIRST_HIT (1u << 2)
#define B200RT_RAY_FLAG_DISABLE_CLOSESTHIT (1u << 3)
#define B200RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES (1u << 4)
#define B200RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES (1u << 5)
#define B200RT_RAY_FLAG_CULL_DISABLED_ANYHIT (1u << 6)
#define B200RT_RAY_FLAG_CULL_ENFORCED_ANYHIT (1u << 7)
/* optixTrace's visibilityMask argument (8 bits, include/optix_types.h OptixVisibilityMask; SDK/imgui_test/optixTriangle.cu traces with 255,
 * optixPathTracer.cu:181,209 and optixRaycasting.cu with 1): OR it into a ray_flags word with this macro.  An instance is traversed when
 * (OptixInstance::visibilityMask & ray mask) != 0.  The field is stored XOR 1, so a flags word without it means mask 1. */
#define B200RT_RAY_VISIBILITY_MASK(m) (((((unsigned int)(m)) ^ 1u) & 0xffu) << 16)
/* instance flags (b200rt_instance.flags), reference include/optix_types.h:1088-1115 */
#define B200RT_INSTANCE_FLAG_NONE 0u
#define B200RT_INSTANCE_FLAG_DISABLE_TRIANGLE_FACE_CULLING (1u << 0)
#define B200RT_INSTANCE_FLAG_FLIP_TRIANGLE_FACING (1u << 1)
#define B200RT_INSTANCE_FLAG_DISABLE_ANYHIT (1u << 2)
#define B200RT_INSTANCE_FLAG_ENFORCE_ANYHIT (1u << 3)
/* geometry flags (b200rt triangle input, one per SBT record), reference include/optix_types.h:311-325 */
#define B200RT_GEOMETRY_FLAG_NONE 0u
#define B200RT_GEOMETRY_FLAG_DISABLE_ANYHIT (1u << 0)
#define B200RT_GEOMETRY_FLAG_REQUIRE_SINGLE_ANYHIT_CALL (1u << 1)
#define B200RT_GEOMETRY_FLAG_DISABLE_TRIANGLE_FACE_CULLING (1u << 2)

typedef struct b200rt_context_t* b200rt_context;
typedef uint64_t b200rt_deviceptr;         /* CUdeviceptr */
typedef uint64_t b200rt_traversable;       /* OptixTraversableHandle: device address of the accel blob */
typedef void* b200rt_stream;               /* cudaStream_t / CUstream */
typedef void (*b200rt_log_cb)(unsigned int level, const char* tag, const char* message, void* cbdata);

/* layout == OptixBuildInputTriangleArray (reference include/optix_types.h:632-705); the two micromap
 * members at the tail are opaque padding here (unused on this path). */
typedef struct b200rt_build_input_triangle_array {
    const b200rt_deviceptr* vertexBuffers; /* host array of device pointers; [0] is used (no motion) */
    unsigned int numVertices;
    unsigned int vertexFormat;             /* B200RT_VERTEX_FORMAT_FLOAT3 */
    unsigned int vertexStrideInBytes;      /* 0 => 12 */
    b200rt_deviceptr indexBuffer;          /* 0 => unindexed */
    unsigned int numIndexTriplets;
    unsigned int indexFormat;
    unsigned int indexStrideInBytes;       /* 0 => 3 * index size */
    b200rt_deviceptr preTransform;         /* device 3x4 row-major float matrix or 0 */
    const unsigned int* flags;             /* host array, one per SBT record (geometry flags) */
    unsigned int numSbtRecords;
    b200rt_deviceptr sbtIndexOffsetBuffer; /* per-primitive local SBT index, or 0 */
    unsigned int sbtIndexOffsetSizeInBytes;   /* 1, 2 or 4 */
    unsigned int sbtIndexOffsetStrideInBytes; /* 0 => size */
    unsigned int primitiveIndexOffset;
    unsigned int transformFormat;
    char opaque_micromaps[240 - 96];
} b200rt_build_input_triangle_array;

/* layout == OptixBuildInputInstanceArray (reference include/optix_types.h:970-988) */
typedef struct b200rt_build_input_instance_array {
    b200rt_deviceptr instances;            /* device array of b200rt_instance (== OptixInstance, 80 B) */
    unsigned int numInstances;
    unsigned int instanceStride;           /* 0 => 80 */
} b200rt_build_input_instance_array;

/* layout == OptixBuildInput (reference include/optix_types.h:1032-1051), 1032 bytes */
typedef struct b200rt_build_input {
    unsigned int type;
    union {
        b200rt_build_input_triangle_array triangleArray;
        b200rt_build_input_instance_array instanceArray;
        char pad[1024];
    };
} b200rt_build_input;

/* layout == OptixInstance (reference include/optix_types.h:1122-1147), 80 bytes */
typedef struct b200rt_instance {
    float transform[12];                   /* object->world, 3x4 row major */
    unsigned int instanceId;
    unsigned int sbtOffset;
    unsigned int visibilityMask;
    unsigned int flags;
    b200rt_traversable traversableHandle;
    unsigned int pad[2];
} b200rt_instance;

/* layout == OptixAccelBuildOptions (reference include/optix_types.h:1331-1346), 20 bytes */
typedef struct b200rt_accel_build_options {
    unsigned int buildFlags;
    unsigned int operation;
    struct { unsigned short numKeys, flags; float timeBegin, timeEnd; } motionOptions;
} b200rt_accel_build_options;

/* layout == OptixAccelBufferSizes (reference include/optix_types.h:1353-1368) */
typedef struct b200rt_accel_buffer_sizes {
    size_t outputSizeInBytes;
    size_t tempSizeInBytes;
    size_t tempUpdateSizeInBytes;
} b200rt_accel_buffer_sizes;

/* layout == OptixAccelEmitDesc (reference include/optix_types.h:1385-1392) */
typedef struct b200rt_accel_emit_desc {
    b200rt_deviceptr result;               /* device address that receives a size_t */
    unsigned int type;                     /* B200RT_PROPERTY_TYPE_COMPACTED_SIZE */
} b200rt_accel_emit_desc;

/* layout == OptixShaderBindingTable (reference include/optix_types.h:2293-2328), 64 bytes */
typedef struct b200rt_shader_binding_table {
    b200rt_deviceptr raygenRecord;
    b200rt_deviceptr exceptionRecord;
    b200rt_deviceptr missRecordBase;
    unsigned int missRecordStrideInBytes;
    unsigned int missRecordCount;
    b200rt_deviceptr hitgroupRecordBase;
    unsigned int hitgroupRecordStrideInBytes;
    unsigned int hitgroupRecordCount;
    b200rt_deviceptr callablesRecordBase;
    unsigned int callablesRecordStrideInBytes;
    unsigned int callablesRecordCount;
} b200rt_shader_binding_table;

/* ---------------------------------------------------------------------------------------------
 * Context.  Replaces cudaFree(0) + optixInit + optixDeviceContextCreate/Destroy
 * (reference SDK/optixPathTracer/optixPathTracer.cpp:555-573, SDK/sutil/Scene.cpp:800-815).
 * log level / callback signature as OptixDeviceContextOptions (level 0 = off ... 4 = print).
 * ------------------------------------------------------------------------------------------- */
int b200rt_context_create(int cuda_device, b200rt_log_cb cb, void* cbdata, int level, b200rt_context* out);
int b200rt_context_destroy(b200rt_context ctx);
const char* b200rt_error_string(int code);     /* optixGetErrorString */
const char* b200rt_error_name(int code);       /* optixGetErrorName */
const char* b200rt_last_error_message(b200rt_context ctx); /* detail text of the last failure on this context */
const char* b200rt_version(void);
uint64_t b200rt_context_kernel_launches(b200rt_context ctx); /* kernels launched so far through this context */
/* enablePeerAccess of the reference's optixNVLink (SDK/optixNVLink/optixNVLink.cpp:1617-1635): lets launches of this context store into
 * memory of `peer_device` — one result buffer for all GPUs (Params::result_buffer of optixMultiGPU pointing into another GPU's HBM, the
 * pixels travel over NVLink).  B200RT_ERROR_NOT_SUPPORTED when the two GPUs have no peer path. */
int b200rt_enable_peer_access(b200rt_context ctx, int peer_device);
/* One result buffer for several processes (one process per GPU): the owner creates it in its HBM (zero-filled) and hands the 64-byte
 * handle (a cudaIpcMemHandle_t) to the others, which open it with THEIR context: the pointer they get may be used as
 * Params::result_buffer of their launches, whose stores then go to the owner's memory over NVLink.  The reference's counterpart is the
 * single result buffer all devices of optixMultiGPU write (zero-copy: SDK/sutil/CUDAOutputBuffer.h:203-216; peer memory:
 * SDK/optixNVLink/optixNVLink.cpp:1975-1992) — there inside one process. */
int b200rt_shared_buffer_create(b200rt_context ctx, size_t bytes, b200rt_deviceptr* ptr, unsigned char handle64[64]);
int b200rt_shared_buffer_open(b200rt_context ctx, const unsigned char handle64[64], b200rt_deviceptr* ptr);
int b200rt_shared_buffer_close(b200rt_context ctx, b200rt_deviceptr ptr);    /* pointer from _open */
int b200rt_shared_buffer_destroy(b200rt_context ctx, b200rt_deviceptr ptr);  /* pointer from _create */

/* ---------------------------------------------------------------------------------------------
 * Acceleration structures.  Replace optixAccelComputeMemoryUsage / optixAccelBuild /
 * optixAccelCompact (reference SDK/optixPathTracer/optixPathTracer.cpp:627-684,
 * SDK/sutil/Scene.cpp:970,1043-1054,1106,1195, SDK/imgui_test/triangle_gas.cpp:213-234).
 * Triangle inputs give a GAS (all inputs of one call in one structure, SBT offsets accumulate over
 * inputs like OptiX); one instance input gives an IAS over previously built GAS handles.
 * outputBuffer must be 128-byte aligned; the returned handle is valid as long as outputBuffer is,
 * and does not reference tempBuffer or the vertex/index buffers after the build has run on `stream`.
 * The build records the exact size; emitted COMPACTED_SIZE properties are written on `stream`.
 * Builds and compactions are ASYNCHRONOUS like optixAccelBuild / optixAccelCompact: no call waits for the device (the loops of the
 * build whose trip counts only the device knows run as CUDA graphs with conditional nodes).  Any input builds (no depth limit is
 * imposed on the caller: the builder keeps every tree within the traversal stack).  What can only be known when the device has run —
 * an output buffer too small for the compaction, a source that is no traversable — is left in the header of the result and reported by
 * b200rt_accel_get_info.
 * b200rt_accel_emit_property replaces optixAccelEmitProperty (reference include/optix_stubs.h:520): COMPACTED_SIZE (size_t) or AABBS
 * (one OptixAabb = 6 floats) of a finished structure, written to device memory on `stream`.
 * ------------------------------------------------------------------------------------------- */
int b200rt_accel_compute_memory_usage(b200rt_context ctx, const b200rt_accel_build_options* options,
                                      const b200rt_build_input* inputs, unsigned int num_inputs,
                                      b200rt_accel_buffer_sizes* sizes);
int b200rt_accel_build(b200rt_context ctx, b200rt_stream stream, const b200rt_accel_build_options* options,
                       const b200rt_build_input* inputs, unsigned int num_inputs, b200rt_deviceptr temp_buffer,
                       size_t temp_bytes, b200rt_deviceptr output_buffer, size_t output_bytes,
                       b200rt_traversable* handle, const b200rt_accel_emit_desc* emitted, unsigned int num_emitted);
int b200rt_accel_compact(b200rt_context ctx, b200rt_stream stream, b200rt_traversable input,
                         b200rt_deviceptr output_buffer, size_t output_bytes, b200rt_traversable* handle);
int b200rt_accel_emit_property(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle,
                               const b200rt_accel_emit_desc* emitted, unsigned int num_emitted);

/* Introspection used by tests / bench (no OptiX equivalent): synchronous. */
typedef struct b200rt_accel_info {
    uint32_t kind;            /* 1 = GAS, 2 = IAS */
    uint32_t num_triangles;
    uint32_t num_nodes;       /* 8-wide nodes */
    uint32_t num_instances;
    uint64_t total_bytes;     /* exact (compacted) size */
    float bounds[6];          /* object-space (GAS) / world-space (IAS) AABB */
    uint32_t depth;           /* levels of the wide tree */
    uint32_t reserved;
} b200rt_accel_info;
int b200rt_accel_get_info(b200rt_context ctx, b200rt_traversable handle, b200rt_accel_info* info);

/* ---------------------------------------------------------------------------------------------
 * optixPathTracer launch.  Replaces
 *   optixLaunch(pipeline, stream, d_params, sizeof(Params), &sbt, width, height, 1)
 * (reference SDK/optixPathTracer/optixPathTracer.cpp:488-511) together with the device programs
 * __raygen__rg / __miss__radiance / __closesthit__radiance (SDK/optixPathTracer/optixPathTracer.cu:249-413).
 * d_params: device copy of the reference's 152-byte Params (SDK/optixPathTracer/optixPathTracer.h:82-107);
 * sbt: host struct whose records live on the device — miss record data = MissData{float4 bg_color},
 * hit-group record data = HitGroupData{float3 emission_color; float3 diffuse_color; float4* vertices}
 * (optixPathTracer.h:115-126), one record per material, selected by the per-primitive SBT index given to
 * the accel build.  width/height are the launch dimensions.
 * ------------------------------------------------------------------------------------------- */
#define B200RT_PT_TERMINATE_RUSSIAN_ROULETTE 0 /* optixPathTracer.cu:294-297 */
#define B200RT_PT_TERMINATE_DEPTH_CAP 1        /* optixMultiGPU.cu:271 */
typedef struct b200rt_pt_stats {
    uint64_t radiance_segments;  /* closest-hit rays traced */
    uint64_t shadow_segments;    /* occlusion rays traced */
    uint32_t iterations;         /* wavefront iterations (trace+shade pairs) */
    uint32_t kernel_launches;    /* kernels launched by this call */
    uint64_t nodes_fetched;      /* B200RT_PT_STATS_TRAVERSAL: 8-wide nodes fetched by the trace stage */
    uint64_t tris_tested;        /* B200RT_PT_STATS_TRAVERSAL: triangle records tested by the trace stage */
    float trace_ms;              /* B200RT_PT_STATS_TIMING: sum of CUDA-event durations of the trace kernels */
    float shade_ms;              /* B200RT_PT_STATS_TIMING: same for the shade kernels */
    uint32_t trace_launches;     /* launches that had at least one active lane */
    uint32_t reserved;
} b200rt_pt_stats;
#define B200RT_PT_STATS_SEGMENTS 1u   /* fill segment / iteration / launch counts */
#define B200RT_PT_STATS_TIMING 2u     /* bracket every stage kernel with CUDA events on `stream` (profiling pass) */
#define B200RT_PT_STATS_TRAVERSAL 4u  /* run the instrumented trace kernel that counts nodes / triangles (not for timing) */
typedef struct b200rt_pt_options {
    uint32_t sample_groups;      /* 0 / 1: one lane per launch index runs its samples_per_launch samples back to back — the reference's
                                  * flat fp32 summation order.  G > 1: G lanes per launch index run samples [g*spl/G, (g+1)*spl/G) in parallel
                                  * (same RNG streams) and the pixel value is ((s_0 + s_1) + ...) + s_{G-1} of their sums: G x fewer, G x
                                  * wider wavefront iterations; differs from G = 1 only in that summation order (~1 ulp). */
    uint32_t collect_stats;      /* bit mask of B200RT_PT_STATS_*; non-zero fills *stats */
    b200rt_pt_stats* stats;      /* host pointer or NULL */
    uint32_t ray_sort;           /* 1: before every trace stage but the first, order the rays by (origin cell, direction) with a radix sort
                                  * — for scenes whose BVH lives in HBM, where incoherent bounce rays are bound by node-fetch latency.
                                  * Costs one host synchronisation per wavefront iteration.  Results are unaffected (a lane owns its state). */
    uint32_t reserved;
} b200rt_pt_options;
int b200rt_launch_pathtracer(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params,
                             const b200rt_shader_binding_table* sbt, unsigned int width, unsigned int height,
                             const b200rt_pt_options* options);

/* optixMultiGPU launch: optixLaunch(pipeline, stream, d_params, sizeof(Params)=168, &sbt, num_samples, 1, 1)
 * (reference SDK/optixMultiGPU/optixMultiGPU.cpp:562-594) + programs of SDK/optixMultiGPU/optixMultiGPU.cu:214-385.
 * d_params: device copy of the 168-byte Params (SDK/optixMultiGPU/optixMultiGPU.h:46-64); the pixel of sample i is
 * params.sample_index_buffer[i] (int2).  SBT: 2 miss records, 2 hit-group records per material (even = radiance). */
int b200rt_launch_multigpu(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params,
                           const b200rt_shader_binding_table* sbt, unsigned int num_samples,
                           const b200rt_pt_options* options);

/* fillSamplesCUDA (reference SDK/optixMultiGPU/optixMultiGPU_kernels.cu:31-55): sample_indices[i] =
 * StaticWorkDistribution(width,height,num_gpus).getSamplePixel(gpu_idx, i) as int2. */
int b200rt_fill_samples(b200rt_context ctx, b200rt_stream stream, int gpu_idx, int num_gpus, int width, int height,
                        b200rt_deviceptr sample_indices_int2, int num_samples);

/* De-interleave after a gather (new; replaces the zero-copy writes of SDK/sutil/CUDAOutputBuffer.h:203-216):
 * gathered = num_gpus consecutive blocks of num_samples float4 (rank-major, as ncclAllGather lays them out);
 * writes accum (float4, width*height, may be 0) and frame (uchar4, make_color) in pixel order. */
int b200rt_deinterleave(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr gathered_float4, int num_gpus,
                        int num_samples, int width, int height, b200rt_deviceptr accum_float4,
                        b200rt_deviceptr frame_uchar4);

/* optixMeshViewer launch: optixLaunch(scene.pipeline(), 0, d_params, sizeof(whitted::LaunchParams)=128, scene.sbt(), width, height, 1)
 * (reference SDK/optixMeshViewer/optixMeshViewer.cpp:283-308) + the programs of SDK/cuda/whitted.cu:44-98,139-289 and
 * getLocalGeometry / sampleTexture (SDK/cuda/LocalGeometry.h:59-176, SDK/cuda/LocalShading.h:37-53).
 * d_params: device copy of whitted::LaunchParams (SDK/cuda/whitted.h:59-77); its `lights` BufferView points to Light records
 * (SDK/cuda/Light.h:31-71).  SBT: as sutil::Scene::createSBT builds it (SDK/sutil/Scene.cpp:1405-1433) — per primitive group a
 * radiance and an occlusion record, each carrying whitted::HitGroupData {GeometryData, MaterialData}; MaterialData textures are
 * cudaTextureObject_t handles and are sampled with tex2D like the reference does.
 * All three alpha modes: MASK / BLEND geometry (geometry flags without DISABLE_ANYHIT, SDK/sutil/Scene.cpp:904-966) runs
 * __anyhit__radiance / __anyhit__occlusion (whitted.cu:100-137: alpha cut-outs, pending / committed occlusion attenuation) inside
 * traversal, and ALPHA_MODE_BLEND hits continue behind themselves up to MAX_TRACE_DEPTH 8 (whitted.cu:266-286).  A scene whose
 * hit-group records hold a BLEND material costs one host synchronisation per level of continuation; other scenes none.
 * Errors found by the kernels — LaunchParams grew more lights than the launch's workspace was sized for, a BLEND material
 * appeared in records that had none when they were first seen — are asynchronous like CUDA's: the NEXT launch on the context
 * returns B200RT_ERROR_INVALID_OPERATION (and looks at the scene again).
 * The first launch with a given (d_params, hit-group records) after an accel build reads the light count, the records' alpha modes
 * and the traversable's any-hit flag back once (synchronises `stream`); later launches are fully asynchronous. */
int b200rt_launch_whitted(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params,
                          const b200rt_shader_binding_table* sbt, unsigned int width, unsigned int height);

/* sutil::Scene::addImage + addSampler (reference SDK/sutil/Scene.cpp:576-652): 8-bit RGBA image (as tinygltf decodes every glTF
 * image, SDK/support/tinygltf/tiny_gltf.h:2368) -> CUDA array + texture object (normalised coordinates, normalised-float reads).
 * address_s / address_t: cudaTextureAddressMode (0 wrap, 1 clamp, 2 mirror); linear_filter: 0 point, 1 linear. */
int b200rt_texture_create(b200rt_context ctx, int width, int height, const void* rgba8, int address_s, int address_t,
                          int linear_filter, uint64_t* texture_object, uint64_t* cuda_array);
int b200rt_texture_destroy(b200rt_context ctx, uint64_t texture_object, uint64_t cuda_array);
/* A texture object of ctx's device over an EXISTING CUDA array with the sampler state of b200rt_texture_create.  The array may live on
 * another device that ctx's device has peer access to (b200rt_enable_peer_access): optixNVLink's texture sharing keeps one copy of a
 * texture per P2P island and defines a sampler over it on every device of the island (defineTextureOnDevice / loadTexture, reference
 * SDK/optixNVLink/optixNVLink.cpp:1446-1468,1522-1561).  Destroy with b200rt_texture_destroy(ctx, texture_object, 0): the array
 * belongs to the context that created it.  One process drives the devices (the reference's model); CUDA arrays cannot cross
 * processes, so with one process per GPU every rank loads its own copy. */
int b200rt_texture_view(b200rt_context ctx, uint64_t cuda_array, int address_s, int address_t, int linear_filter,
                        uint64_t* texture_object);

/* imgui_test ("playground") launch: optixLaunch(pipeline, stream, d_param, sizeof(Params)=128, &sbt, buf_width, buf_height, 1)
 * (reference SDK/imgui_test/tracer_window.cpp:96-105) + the programs of SDK/imgui_test/optixTriangle.cu:103-268.
 * d_params: device copy of the sample's Params (SDK/imgui_test/optixTriangle.h:42-108) — its camera / lights / materials / normals /
 * mat_indices / film / image pointers are device pointers to the sample's own objects (Camera 92 B, LightVariant 44 B,
 * DiffuseMaterial 12 B).  No SBT is needed: the sample's single hit group reads everything from Params.  Params::dt must already
 * hold the frame's value (frame_step()).  Reads the 128-byte Params back once (synchronises `stream`), then runs asynchronously.
 * options->stats (optional): radiance_segments = primary rays, iterations = samples, kernel_launches. */
int b200rt_launch_playground(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, unsigned int width,
                             unsigned int height, const b200rt_pt_options* options);

/* Procedural stand-in for the model behind the reference's cover.png (not in the reference repo): 5 x 5
 * displaced lat-long blobs of 4*rows^2 triangles each over the sample's own 800-triangle floor (triangle_gas.cpp:147-166), as the
 * unindexed float3 vertex / per-vertex normal arrays and int material indices TriangleGAS keeps.  vertices_float3 == 0: only
 * *num_triangles is written (size query). */
int b200rt_generate_playground_scene(b200rt_context ctx, b200rt_stream stream, uint32_t rows, uint32_t seed,
                                     b200rt_deviceptr vertices_float3, b200rt_deviceptr normals_float3,
                                     b200rt_deviceptr mat_indices_i32, uint64_t* num_triangles);

/* ---------------------------------------------------------------------------------------------
 * optixRaycasting.  b200rt_launch_raycast replaces
 *   optixLaunch(pipeline, stream, d_params, sizeof(Params)=24, &sbt, width, height, 1)
 * (reference SDK/optixRaycasting/optixRaycasting.cpp:289-317) + __raygen__from_buffer /
 * __miss__buffer_miss / __closesthit__buffer_hit (SDK/optixRaycasting/optixRaycasting.cu:45-86).
 * d_params: device copy of {handle, Ray* rays, Hit* hits}; Ray = {float3 origin; float tmin; float3 dir;
 * float tmax} (32 B), Hit = {float t; float3 geom_normal} (16 B) (optixRaycastingKernels.h:35-47).
 * SBT hit-group record data = whitted::HitGroupData (352 B, SDK/cuda/whitted.h:44-48): the triangle-mesh
 * BufferViews of its GeometryData are read for the shading normal (SDK/cuda/LocalGeometry.h:59-176).
 * ext_hits (optional, 0 = none): n x {float t; uint32 prim; uint32 inst; float b1; float b2} with the
 * un-truncated t, build-input-local primitive index, instance index, and OptiX barycentrics; miss: t=-1.
 * ------------------------------------------------------------------------------------------- */
int b200rt_launch_raycast(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params,
                          const b200rt_shader_binding_table* sbt, unsigned int width, unsigned int height,
                          b200rt_deviceptr ext_hits);
/* createRaysOrthoOnDevice / translateRaysOnDevice / shadeHitsOnDevice
 * (reference SDK/optixRaycasting/optixRaycastingKernels.cu:42-115); stream-ordered on `stream`. */
int b200rt_create_rays_ortho(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr rays, int width, int height,
                             const float bbmin[3], const float bbmax[3], float padding);
int b200rt_translate_rays(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr rays, int count,
                          const float offset[3]);
int b200rt_shade_hits(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr image_float3, int count,
                      b200rt_deviceptr hits);

/* Generic ray queries over a traversable (the optixTrace semantics the launches are built from; used by the
 * parity tests).  rays: n x Ray (32 B).  closest: ext hit records (20 B, as above).  any: n x uint32 (1 = occluded). */
int b200rt_trace_closest(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays,
                         uint64_t n, unsigned int ray_flags, b200rt_deviceptr ext_hits);
int b200rt_trace_any(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays,
                     uint64_t n, unsigned int ray_flags, b200rt_deviceptr occluded_u32);
/* Instrumented closest-hit pass (not for timing): totals of 8-wide nodes fetched and triangles tested. */
int b200rt_trace_stats(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays,
                       uint64_t n, uint64_t* nodes_fetched, uint64_t* tris_tested);

/* ---------------------------------------------------------------------------------------------
 * Host-side sutil mirrors (pure host arithmetic, same results as the reference's host code).
 * ------------------------------------------------------------------------------------------- */
/* sutil::Camera::UVWFrame (reference SDK/sutil/Camera.cpp:34-46) */
void b200rt_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fov_y_deg,
                       float aspect, float U[3], float V[3], float W[3]);
/* imgui_test Camera (reference SDK/imgui_test/camera.h:19-117) set up as main.cpp:236-243 does — setters then compute_uvw();
 * writes the 92-byte object Params::camera points to. */
void b200rt_playground_camera(const float eye[3], const float up[3], const float lookat[3], float aperture, float fd,
                              float fov_deg, int ortho, void* camera92);
/* How OptixRayFlags and OptixInstanceFlags combine for the triangles of one instance (host function, no GPU; reference
 * include/optix_types.h:1088-1108, 1794-1839): bits 4-7 = the ray's CULL_* flags after DISABLE_TRIANGLE_FACE_CULLING (face bits dropped)
 * or FLIP_TRIANGLE_FACING (face bits swapped); bits 0-1 = any-hit override, 1 = off for every triangle, 2 = on, 0 = the geometry flag
 * decides; ray flags take precedence over instance flags.  Exposed so that the CPU oracle and the device code can be compared. */
unsigned int b200rt_triangle_flag_word(unsigned int ray_flags, unsigned int instance_flags);
/* StaticWorkDistribution (reference SDK/sutil/WorkDistribution.h:50-81) */
int b200rt_wd_num_samples(int width, int height, int num_gpus);
void b200rt_wd_sample_pixel(int width, int height, int num_gpus, int gpu_idx, int sample_idx, int xy[2]);

/* Procedural tessellated mesh for the synthetic multi-GPU configuration (BASELINE.json configs[4]; the
 * reference has no such asset — SURVEY.md §8(d) C5).  Writes num_triangles*3 float4 vertices (w = 0), the
 * layout optixPathTracer's HitGroupData::vertices uses, and num_triangles uint32 material indices. */
int b200rt_generate_synthetic_mesh(b200rt_context ctx, b200rt_stream stream, uint64_t num_triangles, uint32_t seed,
                                   b200rt_deviceptr vertices_float4, b200rt_deviceptr mat_indices_u32,
                                   float bounds_out[6]);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
