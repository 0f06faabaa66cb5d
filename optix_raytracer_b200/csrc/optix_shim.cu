// optix_shim.cu — optixQueryFunctionTable: the OptiX 8.0 (ABI 87) host entry point, answered by this library.
//
// The reference never links OptiX: optixInit() dlopens "libnvoptix.so.1", looks up optixQueryFunctionTable and lets it fill a table of
// 48 function pointers (reference include/optix_stubs.h:198-229, include/optix_function_table.h:46-343); every optixFoo() of the host
// code is an inline stub that calls through that table.  A build of this library with the soname libnvoptix.so.1
// (optix_raytracer_b200/optix_shim/libnvoptix.so.1, csrc/Makefile) put in front of the driver's on LD_LIBRARY_PATH therefore runs the
// reference's UNMODIFIED host code — optixDeviceContextCreate, optixModuleCreate, optixProgramGroupCreate, optixPipelineCreate,
// optixAccelBuild, optixSbtRecordPackHeader, optixLaunch (SDK/optixPathTracer/optixPathTracer.cpp:555-898,
// SDK/optixRaycasting/optixRaycasting.cpp:94-252, SDK/optixMultiGPU/optixMultiGPU.cpp:641-1018, SDK/sutil/Scene.cpp:800-1433,
// SDK/imgui_test/main.cpp:71-188) — on the B200-native launches of this repo (SURVEY.md 8(b) layer 2).
//
// What is honoured and what is not:
//   * the module input (PTX / OptiX-IR of the sample's device programs) is NOT compiled: the device programs are this library's
//     restatements.  A pipeline is recognised by the entry-function names of its program groups:
//       __raygen__rg + __miss__radiance       optixPathTracer (Params 152 B) or optixMultiGPU (168 B), told apart at optixLaunch
//       __raygen__from_buffer                 optixRaycasting
//       __raygen__pinhole                     optixMeshViewer (cuda/whitted.cu)
//       __raygen__rg + __miss__ms             imgui_test
//     any other set of programs fails optixPipelineCreate with OPTIX_ERROR_NOT_SUPPORTED (and a log line saying so);
//   * acceleration-structure calls pass straight through (the build structs are layout-identical, include/b200rt.h);
//   * stack sizes, caches, tasks, relocation, micromaps, the denoiser: accepted as no-ops where the samples call them
//     unconditionally (stack sizes, cache settings), OPTIX_ERROR_NOT_SUPPORTED otherwise.
// Struct layouts of the OptiX API used here are mirrored below (no OptiX header is needed to build the product) and pinned against the
// reference's own headers in tests/golden/kat.json "optix_api_layout" (oracle/ref_shim.cpp, tests/test_abi_cpu.py).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/b200rt.h"

namespace {

constexpr int OPTIX_SUCCESS_ = 0;
constexpr int OPTIX_ERROR_INVALID_VALUE_ = 7001;
constexpr int OPTIX_ERROR_HOST_OUT_OF_MEMORY_ = 7002;
constexpr int OPTIX_ERROR_NOT_SUPPORTED_ = 7800;
constexpr int OPTIX_ERROR_UNSUPPORTED_ABI_VERSION_ = 7801;
constexpr int OPTIX_ERROR_FUNCTION_TABLE_SIZE_MISMATCH_ = 7802;
constexpr int OPTIX_ERROR_INVALID_DEVICE_CONTEXT_ = 7051;
constexpr unsigned SHIM_ABI_VERSION = 87;      // include/optix_function_table.h:29
constexpr size_t SHIM_TABLE_ENTRIES = 48;      // include/optix_function_table.h:46-343
constexpr unsigned KIND_RAYGEN = 0x2421, KIND_MISS = 0x2422, KIND_EXCEPTION = 0x2423, KIND_HITGROUP = 0x2424, KIND_CALLABLES = 0x2425;
constexpr unsigned PROP_MAX_TRACE_DEPTH = 0x2001, PROP_RTCORE_VERSION = 0x2005;

typedef void (*LogCb)(unsigned int level, const char* tag, const char* message, void* cbdata);

// OptixDeviceContextOptions (include/optix_types.h), 24 bytes
struct ContextOptions { LogCb logCallbackFunction; void* logCallbackData; int logCallbackLevel; int validationMode; };
static_assert(sizeof(ContextOptions) == 24, "OptixDeviceContextOptions");
// OptixProgramGroupDesc, 56 bytes: kind, flags, union { single {module, entryFunctionName}; hitgroup {CH, AH, IS pairs} }
struct ProgramGroupDesc {
    unsigned kind, flags;
    union {
        struct { void* module; const char* name; } single;
        struct { void* moduleCH; const char* nameCH; void* moduleAH; const char* nameAH; void* moduleIS; const char* nameIS; } hitgroup;
    };
};
static_assert(sizeof(ProgramGroupDesc) == 56, "OptixProgramGroupDesc");
// OptixPipelineCompileOptions, 40 bytes
struct PipelineCompileOptions {
    int usesMotionBlur; unsigned traversableGraphFlags; int numPayloadValues, numAttributeValues; unsigned exceptionFlags; unsigned pad;
    const char* pipelineLaunchParamsVariableName; unsigned usesPrimitiveTypeFlags; int allowOpacityMicromaps;
};
static_assert(sizeof(PipelineCompileOptions) == 40 && offsetof(PipelineCompileOptions, pipelineLaunchParamsVariableName) == 24, "OptixPipelineCompileOptions");
struct StackSizes { unsigned v[7]; };  // OptixStackSizes, 28 bytes

enum Which { W_NONE = 0, W_PATHTRACER_FAMILY, W_RAYCAST, W_WHITTED, W_PLAYGROUND };

unsigned env_sample_groups()
{
    // 1 by default: a drop-in keeps the reference's arithmetic, i.e. the flat fp32 summation order of a pixel's samples.
    // B200RT_SAMPLE_GROUPS=<n> opts in to n parallel lanes per pixel whose sums are added in lane order (the per-pixel sum is regrouped,
    // ~1 ulp; the oracle restates the grouped order too): 1.77x instead of 1.50x OptiX on the Cornell launch.
    const char* e = getenv("B200RT_SAMPLE_GROUPS");
    const long v = e ? strtol(e, nullptr, 10) : 1;
    return v >= 1 && v <= 64 ? (unsigned)v : 1u;
}

struct ShimContext { b200rt_context ctx = nullptr; LogCb cb = nullptr; void* cbdata = nullptr; int level = 0; };
struct ShimModule { ShimContext* c; };
struct ShimProgramGroup { ShimContext* c; unsigned kind; std::string name, name_ah, name_is; };
struct ShimPipeline { ShimContext* c; Which which; };

void say(ShimContext* c, unsigned level, const char* tag, const std::string& msg)
{
    if (c && c->cb && (int)level <= c->level) c->cb(level, tag, msg.c_str(), c->cbdata);
}
void put_log(char* log, size_t* log_size, const std::string& msg)
{
    if (!log_size) return;
    if (log && *log_size) {
        const size_t n = msg.size() < *log_size - 1 ? msg.size() : *log_size - 1;
        memcpy(log, msg.data(), n);
        log[n] = 0;
    }
    *log_size = msg.size() + 1;
}

// ---- error strings -------------------------------------------------------------------------------------------------------------------
// optixGetErrorName: the codes are OptixResult values (include/b200rt.h), so the names are OptiX's with their own prefix
const char* s_error_name(int r)
{
    switch (r) {
        case 0: return "OPTIX_SUCCESS";
        case 7001: return "OPTIX_ERROR_INVALID_VALUE";
        case 7002: return "OPTIX_ERROR_HOST_OUT_OF_MEMORY";
        case 7003: return "OPTIX_ERROR_INVALID_OPERATION";
        case 7050: return "OPTIX_ERROR_LAUNCH_FAILURE";
        case 7051: return "OPTIX_ERROR_INVALID_DEVICE_CONTEXT";
        case 7800: return "OPTIX_ERROR_NOT_SUPPORTED";
        case 7801: return "OPTIX_ERROR_UNSUPPORTED_ABI_VERSION";
        case 7802: return "OPTIX_ERROR_FUNCTION_TABLE_SIZE_MISMATCH";
        case 7900: return "OPTIX_ERROR_CUDA_ERROR";
        default: return "OPTIX_ERROR_UNKNOWN";
    }
}
const char* s_error_string(int r) { return b200rt_error_string(r); }

// ---- device context -------------------------------------------------------------------------------------------------------------------
int s_context_create(void* /*CUcontext fromContext: 0 = current*/, const ContextOptions* opt, ShimContext** out)
{
    if (!out) return OPTIX_ERROR_INVALID_VALUE_;
    *out = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return B200RT_ERROR_CUDA_ERROR;
    ShimContext* c = new (std::nothrow) ShimContext();
    if (!c) return OPTIX_ERROR_HOST_OUT_OF_MEMORY_;
    if (opt) { c->cb = opt->logCallbackFunction; c->cbdata = opt->logCallbackData; c->level = opt->logCallbackLevel; }
    const int rc = b200rt_context_create(dev, (b200rt_log_cb)c->cb, c->cbdata, c->level, &c->ctx);
    if (rc) { delete c; return rc; }
    say(c, 4, "B200RT", std::string("optixDeviceContextCreate answered by ") + b200rt_version());
    *out = c;
    return OPTIX_SUCCESS_;
}
int s_context_destroy(ShimContext* c)
{
    if (!c) return OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
    b200rt_context_destroy(c->ctx);
    delete c;
    return OPTIX_SUCCESS_;
}
int s_context_get_property(ShimContext* c, unsigned prop, void* value, size_t size)
{
    if (!c || !value) return OPTIX_ERROR_INVALID_VALUE_;
    unsigned v = 0;
    if (prop == PROP_RTCORE_VERSION) v = 0;           // no RT cores are used (B200 has none)
    else if (prop == PROP_MAX_TRACE_DEPTH) v = 31;
    else return OPTIX_ERROR_NOT_SUPPORTED_;
    if (size != sizeof(unsigned)) return OPTIX_ERROR_INVALID_VALUE_;
    memcpy(value, &v, sizeof v);
    return OPTIX_SUCCESS_;
}
int s_context_set_log_callback(ShimContext* c, LogCb cb, void* cbdata, unsigned level)
{
    if (!c) return OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
    c->cb = cb; c->cbdata = cbdata; c->level = (int)level;
    return OPTIX_SUCCESS_;
}
// the disk cache holds compiled modules; nothing is compiled here
int s_cache_set_enabled(ShimContext*, int) { return OPTIX_SUCCESS_; }
int s_cache_set_location(ShimContext*, const char*) { return OPTIX_SUCCESS_; }
int s_cache_set_sizes(ShimContext*, size_t, size_t) { return OPTIX_SUCCESS_; }
int s_cache_get_enabled(ShimContext*, int* enabled) { if (enabled) *enabled = 0; return OPTIX_SUCCESS_; }
int s_cache_get_location(ShimContext*, char* location, size_t size) { if (location && size) location[0] = 0; return OPTIX_SUCCESS_; }
int s_cache_get_sizes(ShimContext*, size_t* lo, size_t* hi) { if (lo) *lo = 0; if (hi) *hi = 0; return OPTIX_SUCCESS_; }

// ---- modules --------------------------------------------------------------------------------------------------------------------------
int s_module_create(ShimContext* c, const void* /*moduleCompileOptions*/, const PipelineCompileOptions* /*pco*/, const char* input, size_t input_size,
                    char* log, size_t* log_size, ShimModule** out)
{
    if (!c || !out || !input || !input_size) return OPTIX_ERROR_INVALID_VALUE_;
    ShimModule* m = new (std::nothrow) ShimModule{c};
    if (!m) return OPTIX_ERROR_HOST_OUT_OF_MEMORY_;
    put_log(log, log_size, "b200rt: module input not compiled; the pipeline's device programs are recognised by entry-function name");
    *out = m;
    return OPTIX_SUCCESS_;
}
int s_module_destroy(ShimModule* m) { delete m; return OPTIX_SUCCESS_; }
int s_not_supported() { return OPTIX_ERROR_NOT_SUPPORTED_; }

// ---- program groups -------------------------------------------------------------------------------------------------------------------
int s_program_group_create(ShimContext* c, const ProgramGroupDesc* descs, unsigned n, const void* /*options*/, char* log, size_t* log_size,
                           ShimProgramGroup** out)
{
    if (!c || !descs || !out) return OPTIX_ERROR_INVALID_VALUE_;
    for (unsigned i = 0; i < n; ++i) {
        const ProgramGroupDesc& d = descs[i];
        ShimProgramGroup* g = new (std::nothrow) ShimProgramGroup{c, d.kind, "", "", ""};
        if (!g) {
            for (unsigned k = 0; k < i; ++k) { delete out[k]; out[k] = nullptr; }
            return OPTIX_ERROR_HOST_OUT_OF_MEMORY_;
        }
        if (d.kind == KIND_HITGROUP) {
            if (d.hitgroup.nameCH) g->name = d.hitgroup.nameCH;
            if (d.hitgroup.nameAH) g->name_ah = d.hitgroup.nameAH;
            if (d.hitgroup.nameIS) g->name_is = d.hitgroup.nameIS;
        } else if (d.kind == KIND_RAYGEN || d.kind == KIND_MISS || d.kind == KIND_EXCEPTION) {
            if (d.single.name) g->name = d.single.name;
        } else if (d.kind != KIND_CALLABLES) {
            delete g;
            for (unsigned k = 0; k < i; ++k) { delete out[k]; out[k] = nullptr; }  // nothing of a failed call is left behind
            return OPTIX_ERROR_INVALID_VALUE_;
        }
        out[i] = g;
    }
    put_log(log, log_size, "");
    return OPTIX_SUCCESS_;
}
int s_program_group_destroy(ShimProgramGroup* g) { delete g; return OPTIX_SUCCESS_; }
int s_program_group_get_stack_size(ShimProgramGroup* g, StackSizes* sizes, void* /*pipeline*/)
{
    if (!g || !sizes) return OPTIX_ERROR_INVALID_VALUE_;
    memset(sizes, 0, sizeof *sizes);  // the wavefront launches keep their state in queues, not on a continuation stack
    return OPTIX_SUCCESS_;
}

// ---- pipelines ------------------------------------------------------------------------------------------------------------------------
int s_pipeline_create(ShimContext* c, const PipelineCompileOptions* /*pco*/, const void* /*linkOptions*/, ShimProgramGroup* const* groups, unsigned n,
                      char* log, size_t* log_size, ShimPipeline** out)
{
    if (!c || !groups || !out) return OPTIX_ERROR_INVALID_VALUE_;
    std::string raygen, hits;
    bool miss_radiance = false, miss_ms = false, miss_buffer = false, miss_constant = false;
    bool ch_radiance = false, ch_buffer = false, ch_ch = false;
    for (unsigned i = 0; i < n; ++i) {
        if (!groups[i]) return OPTIX_ERROR_INVALID_VALUE_;
        const std::string& nm = groups[i]->name;
        if (groups[i]->kind == KIND_RAYGEN) raygen = nm;
        if (groups[i]->kind == KIND_MISS) {
            miss_radiance |= nm == "__miss__radiance"; miss_ms |= nm == "__miss__ms"; miss_buffer |= nm == "__miss__buffer_miss";
            miss_constant |= nm == "__miss__constant_radiance";
        }
        if (groups[i]->kind == KIND_HITGROUP) {
            ch_radiance |= nm == "__closesthit__radiance"; ch_buffer |= nm == "__closesthit__buffer_hit"; ch_ch |= nm == "__closesthit__ch";
            hits += (hits.empty() ? "" : ", ") + (nm.empty() ? std::string("(none)") : nm);
        }
    }
    // the whole set of entry points has to be a sample's: a pipeline that reuses a sample's raygen name with other hit programs would
    // otherwise silently run this library's restatement of the sample
    Which w = W_NONE;
    if (raygen == "__raygen__from_buffer" && miss_buffer && ch_buffer) w = W_RAYCAST;
    else if (raygen == "__raygen__pinhole" && miss_constant && ch_radiance) w = W_WHITTED;
    else if (raygen == "__raygen__rg" && miss_ms && ch_ch) w = W_PLAYGROUND;
    else if (raygen == "__raygen__rg" && miss_radiance && ch_radiance) w = W_PATHTRACER_FAMILY;
    if (w == W_NONE) {
        const std::string msg = "b200rt: no restatement of a pipeline with raygen program '" + raygen + "' and closest-hit programs {" + hits +
                                "} (known: optixPathTracer, optixMultiGPU, optixRaycasting, optixMeshViewer/whitted, imgui_test)";
        put_log(log, log_size, msg);
        say(c, 2, "B200RT", msg);
        return OPTIX_ERROR_NOT_SUPPORTED_;
    }
    ShimPipeline* p = new (std::nothrow) ShimPipeline{c, w};
    if (!p) return OPTIX_ERROR_HOST_OUT_OF_MEMORY_;
    if (w == W_PATHTRACER_FAMILY)
        say(c, 4, "B200RT", "path-tracer launches of this pipeline run with sample_groups " + std::to_string(env_sample_groups()) +
                                (env_sample_groups() == 1 ? " (the reference's summation order)" : " (B200RT_SAMPLE_GROUPS: the per-pixel fp32 sum is regrouped)"));
    put_log(log, log_size, "");
    *out = p;
    return OPTIX_SUCCESS_;
}
int s_pipeline_destroy(ShimPipeline* p) { delete p; return OPTIX_SUCCESS_; }
int s_pipeline_set_stack_size(ShimPipeline* p, unsigned, unsigned, unsigned, unsigned) { return p ? OPTIX_SUCCESS_ : OPTIX_ERROR_INVALID_VALUE_; }

// ---- acceleration structures: pass-through --------------------------------------------------------------------------------------------
int s_accel_compute_memory_usage(ShimContext* c, const b200rt_accel_build_options* o, const b200rt_build_input* in, unsigned n, b200rt_accel_buffer_sizes* sizes)
{
    return c ? b200rt_accel_compute_memory_usage(c->ctx, o, in, n, sizes) : OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
}
int s_accel_build(ShimContext* c, void* stream, const b200rt_accel_build_options* o, const b200rt_build_input* in, unsigned n, b200rt_deviceptr temp,
                  size_t temp_bytes, b200rt_deviceptr out, size_t out_bytes, b200rt_traversable* handle, const b200rt_accel_emit_desc* emitted, unsigned n_emitted)
{
    return c ? b200rt_accel_build(c->ctx, (b200rt_stream)stream, o, in, n, temp, temp_bytes, out, out_bytes, handle, emitted, n_emitted)
             : OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
}
int s_accel_compact(ShimContext* c, void* stream, b200rt_traversable in, b200rt_deviceptr out, size_t out_bytes, b200rt_traversable* handle)
{
    return c ? b200rt_accel_compact(c->ctx, (b200rt_stream)stream, in, out, out_bytes, handle) : OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
}

int s_accel_emit_property(ShimContext* c, void* stream, b200rt_traversable handle, const b200rt_accel_emit_desc* emitted)
{
    return c ? b200rt_accel_emit_property(c->ctx, (b200rt_stream)stream, handle, emitted, emitted ? 1u : 0u) : OPTIX_ERROR_INVALID_DEVICE_CONTEXT_;
}
// optixConvertPointerToTraversableHandle: a traversable handle of this library IS the device address of its blob
int s_convert_pointer(ShimContext* c, b200rt_deviceptr pointer, unsigned /*OptixTraversableType*/, b200rt_traversable* handle)
{
    if (!c || !handle) return OPTIX_ERROR_INVALID_VALUE_;
    *handle = pointer;
    return OPTIX_SUCCESS_;
}

// ---- SBT + launch ---------------------------------------------------------------------------------------------------------------------
int s_sbt_record_pack_header(ShimProgramGroup* g, void* header)
{
    if (!g || !header) return OPTIX_ERROR_INVALID_VALUE_;
    // 32 bytes (OPTIX_SBT_RECORD_HEADER_SIZE): a tag, the group kind and the start of the entry name — for a reader of a memory dump;
    // the launches find their programs through the pipeline, not through the record header
    unsigned char h[32] = {0};
    memcpy(h, "B2RT", 4);
    memcpy(h + 4, &g->kind, 4);
    strncpy((char*)h + 8, g->name.c_str(), 23);
    memcpy(header, h, sizeof h);
    return OPTIX_SUCCESS_;
}


int s_launch(ShimPipeline* p, void* stream, b200rt_deviceptr params, size_t params_size, const b200rt_shader_binding_table* sbt, unsigned w, unsigned h,
             unsigned d)
{
    if (!p || !sbt) return OPTIX_ERROR_INVALID_VALUE_;
    if (w == 0 || h == 0 || d == 0) return OPTIX_SUCCESS_;
    b200rt_context ctx = p->c->ctx;
    b200rt_pt_options opts;
    memset(&opts, 0, sizeof opts);
    opts.sample_groups = env_sample_groups();
    if (d != 1) return OPTIX_ERROR_INVALID_VALUE_;  // every sample on the path launches a 2-D (or 1-D) grid
    switch (p->which) {
        case W_PATHTRACER_FAMILY:
            if (params_size == 152) return b200rt_launch_pathtracer(ctx, (b200rt_stream)stream, params, sbt, w, h, &opts);
            if (params_size == 168 && h == 1) return b200rt_launch_multigpu(ctx, (b200rt_stream)stream, params, sbt, w, &opts);
            return OPTIX_ERROR_INVALID_VALUE_;
        case W_RAYCAST:
            return params_size == 24 ? b200rt_launch_raycast(ctx, (b200rt_stream)stream, params, sbt, w, h, 0) : OPTIX_ERROR_INVALID_VALUE_;
        case W_WHITTED:
            return params_size == 128 ? b200rt_launch_whitted(ctx, (b200rt_stream)stream, params, sbt, w, h) : OPTIX_ERROR_INVALID_VALUE_;
        case W_PLAYGROUND:
            return params_size == 128 ? b200rt_launch_playground(ctx, (b200rt_stream)stream, params, w, h, nullptr) : OPTIX_ERROR_INVALID_VALUE_;
        default:
            return OPTIX_ERROR_NOT_SUPPORTED_;
    }
}

}  // namespace

// OptixQueryFunctionTable_t (include/optix_function_table.h:345-353 / optix_stubs.h:225-228)
extern "C" int optixQueryFunctionTable(int abi_id, unsigned int num_options, const void* /*option keys*/, const void** /*option values*/,
                                       void* function_table, size_t size_of_table)
{
    if (abi_id != (int)SHIM_ABI_VERSION) return OPTIX_ERROR_UNSUPPORTED_ABI_VERSION_;
    if (num_options != 0) return OPTIX_ERROR_INVALID_VALUE_;
    if (!function_table) return OPTIX_ERROR_INVALID_VALUE_;
    if (size_of_table != SHIM_TABLE_ENTRIES * sizeof(void*)) return OPTIX_ERROR_FUNCTION_TABLE_SIZE_MISMATCH_;
    void** t = (void**)function_table;
    for (size_t i = 0; i < SHIM_TABLE_ENTRIES; ++i) t[i] = (void*)&s_not_supported;  // a call with any argument list returns NOT_SUPPORTED
    // order of include/optix_function_table.h (ABI 87)
    t[0] = (void*)&s_error_name;                  // optixGetErrorName
    t[1] = (void*)&s_error_string;                // optixGetErrorString
    t[2] = (void*)&s_context_create;              // optixDeviceContextCreate
    t[3] = (void*)&s_context_destroy;             // optixDeviceContextDestroy
    t[4] = (void*)&s_context_get_property;        // optixDeviceContextGetProperty
    t[5] = (void*)&s_context_set_log_callback;    // optixDeviceContextSetLogCallback
    t[6] = (void*)&s_cache_set_enabled;           // optixDeviceContextSetCacheEnabled
    t[7] = (void*)&s_cache_set_location;          // optixDeviceContextSetCacheLocation
    t[8] = (void*)&s_cache_set_sizes;             // optixDeviceContextSetCacheDatabaseSizes
    t[9] = (void*)&s_cache_get_enabled;           // optixDeviceContextGetCacheEnabled
    t[10] = (void*)&s_cache_get_location;         // optixDeviceContextGetCacheLocation
    t[11] = (void*)&s_cache_get_sizes;            // optixDeviceContextGetCacheDatabaseSizes
    t[12] = (void*)&s_module_create;              // optixModuleCreate
    //  13 optixModuleCreateWithTasks, 14 optixModuleGetCompilationState: not supported
    t[15] = (void*)&s_module_destroy;             // optixModuleDestroy
    //  16 optixBuiltinISModuleGet, 17 optixTaskExecute: not supported
    t[18] = (void*)&s_program_group_create;       // optixProgramGroupCreate
    t[19] = (void*)&s_program_group_destroy;      // optixProgramGroupDestroy
    t[20] = (void*)&s_program_group_get_stack_size;  // optixProgramGroupGetStackSize
    t[21] = (void*)&s_pipeline_create;            // optixPipelineCreate
    t[22] = (void*)&s_pipeline_destroy;           // optixPipelineDestroy
    t[23] = (void*)&s_pipeline_set_stack_size;    // optixPipelineSetStackSize
    t[24] = (void*)&s_accel_compute_memory_usage; // optixAccelComputeMemoryUsage
    t[25] = (void*)&s_accel_build;                // optixAccelBuild
    //  26-28 relocation: not supported
    t[29] = (void*)&s_accel_compact;              // optixAccelCompact
    t[30] = (void*)&s_accel_emit_property;        // optixAccelEmitProperty
    t[31] = (void*)&s_convert_pointer;            // optixConvertPointerToTraversableHandle
    //  32-37 micromaps: not supported
    t[38] = (void*)&s_sbt_record_pack_header;     // optixSbtRecordPackHeader
    t[39] = (void*)&s_launch;                     // optixLaunch
    //  40-47 denoiser: not supported
    return OPTIX_SUCCESS_;
}

// what the CPU test suite compares with the layouts measured on the reference's headers (tests/test_abi_cpu.py)
extern "C" int b200rt_optix_shim_layout(unsigned int* out, unsigned int n)
{
    const unsigned v[] = {SHIM_ABI_VERSION, (unsigned)(SHIM_TABLE_ENTRIES * sizeof(void*)), (unsigned)sizeof(ContextOptions), (unsigned)sizeof(ProgramGroupDesc),
                          (unsigned)offsetof(ProgramGroupDesc, single.name), (unsigned)offsetof(ProgramGroupDesc, hitgroup.nameCH),
                          (unsigned)offsetof(ProgramGroupDesc, hitgroup.nameAH), (unsigned)offsetof(ProgramGroupDesc, hitgroup.nameIS),
                          (unsigned)sizeof(PipelineCompileOptions), (unsigned)sizeof(StackSizes), KIND_RAYGEN, KIND_MISS, KIND_HITGROUP,
                          PROP_RTCORE_VERSION};
    const unsigned m = (unsigned)(sizeof v / sizeof v[0]);
    for (unsigned i = 0; i < n && i < m; ++i) out[i] = v[i];
    return (int)m;
}
