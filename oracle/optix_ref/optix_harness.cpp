// optix_harness.cpp — TEST INFRASTRUCTURE (never linked or loaded by the product).
//
// A headless host for the REFERENCE'S OWN device programs on the real OptiX runtime: the reference's
// .cu files are compiled to PTX where they lie under /root/reference (oracle/Makefile, target `optix`),
// this file drives libnvoptix.so.1 through the reference's own headers (include/optix_stubs.h:198-229 does
// the dlopen; include/optix_function_table.h:46-343 is the table), and the Python side
// (oracle/optix_ref/backend.py) feeds it exactly the Params / SBT payloads / build inputs that
// optix_raytracer_b200/host.py feeds libb200rt.so.  That makes OptiX-on-the-same-B200 a second oracle:
//   * hit records of identical ray batches (query_programs.cu: t, primitive, instance, barycentrics);
//   * optixRaycasting's Hit buffer, optixPathTracer's / optixMultiGPU's images (the reference's programs);
//   * the GPU baseline timing the north star asks for (cudaEvents around optixLaunch).
// It restates the generic call sequence of the samples (context -> accel build/compact -> module ->
// program groups -> pipeline + stack sizes -> SBT header packing -> launch:
// SDK/optixPathTracer/optixPathTracer.cpp:555-898, SDK/optixRaycasting/optixRaycasting.cpp:94-252,
// SDK/optixMultiGPU/optixMultiGPU.cpp:641-1018) without any windowing / sutil dependency.
// libnvoptix.so.1 exists only on the GPU box; in the build container this file merely compiles.
#include <cuda_runtime.h>
#include <optix.h>
#include <optix_function_table_definition.h>
#include <optix_stack_size.h>
#include <optix_stubs.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Pipeline {
    OptixModule module = nullptr;
    OptixPipeline pipeline = nullptr;
    OptixProgramGroup raygen = nullptr;
    std::vector<OptixProgramGroup> miss, hit;
};

OptixDeviceContext g_ctx = nullptr;
std::vector<Pipeline> g_pipelines;
std::string g_log;
int g_log_level = 0;

void log_cb(unsigned int level, const char* tag, const char* message, void*)
{
    char buf[2048];
    snprintf(buf, sizeof buf, "[%u][%s] %s\n", level, tag ? tag : "", message ? message : "");
    g_log += buf;
    if (g_log.size() > (1u << 20)) g_log.erase(0, g_log.size() - (1u << 19));
    if ((int)level <= g_log_level) fputs(buf, stderr);
}

int fail(const char* what, int code)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %d (%s)\n", what, code,
             (g_optixFunctionTable.optixGetErrorName && code >= 7000) ? optixGetErrorName((OptixResult)code) : "cuda/harness");
    g_log += buf;
    return code ? code : -1;
}

#define OCHK(call)                                      \
    do {                                                \
        OptixResult r_ = (call);                        \
        if (r_ != OPTIX_SUCCESS) return fail(#call, r_); \
    } while (0)
#define CCHK(call)                                                \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return fail(#call, 100000 + (int)e_); \
    } while (0)

std::vector<std::string> split(const char* csv)
{
    std::vector<std::string> out;
    if (!csv || !*csv) return out;
    std::stringstream ss(csv);
    std::string item;
    while (std::getline(ss, item, ',')) out.push_back(item);
    return out;
}

}  // namespace

extern "C" {

const char* oref_log() { return g_log.c_str(); }
void oref_log_clear() { g_log.clear(); }

// cudaFree(0) + optixInit + optixDeviceContextCreate (optixPathTracer.cpp:555-573)
int oref_init(int device, int log_level)
{
    if (g_ctx) return 0;
    g_log_level = log_level;
    CCHK(cudaSetDevice(device));
    CCHK(cudaFree(0));
    OCHK(optixInit());
    OptixDeviceContextOptions opt = {};
    opt.logCallbackFunction = &log_cb;
    opt.logCallbackLevel = 4;
    OCHK(optixDeviceContextCreate(nullptr, &opt, &g_ctx));
    // the on-disk shader cache would hide JIT cost differences between runs on fresh boxes; it is harmless, keep default
    return 0;
}

int oref_shutdown()
{
    for (auto& p : g_pipelines) {
        if (p.pipeline) optixPipelineDestroy(p.pipeline);
        if (p.raygen) optixProgramGroupDestroy(p.raygen);
        for (auto g : p.miss) optixProgramGroupDestroy(g);
        for (auto g : p.hit) optixProgramGroupDestroy(g);
        if (p.module) optixModuleDestroy(p.module);
    }
    g_pipelines.clear();
    if (g_ctx) optixDeviceContextDestroy(g_ctx);
    g_ctx = nullptr;
    return 0;
}

unsigned oref_rtcore_version()
{
    unsigned v = 0;
    if (g_ctx) optixDeviceContextGetProperty(g_ctx, OPTIX_DEVICE_PROPERTY_RTCORE_VERSION, &v, sizeof v);
    return v;
}

// ---- acceleration structures: straight pass-through (b200rt's build structs are layout-identical) -------
int oref_accel_compute_memory_usage(const void* options, const void* inputs, unsigned n, void* sizes)
{
    OCHK(optixAccelComputeMemoryUsage(g_ctx, (const OptixAccelBuildOptions*)options, (const OptixBuildInput*)inputs, n, (OptixAccelBufferSizes*)sizes));
    return 0;
}

int oref_accel_build(void* stream, const void* options, const void* inputs, unsigned n, unsigned long long temp, size_t temp_bytes,
                     unsigned long long out, size_t out_bytes, unsigned long long* handle, const void* emitted, unsigned n_emitted)
{
    OptixTraversableHandle h = 0;
    OCHK(optixAccelBuild(g_ctx, (CUstream)stream, (const OptixAccelBuildOptions*)options, (const OptixBuildInput*)inputs, n, (CUdeviceptr)temp,
                         temp_bytes, (CUdeviceptr)out, out_bytes, &h, (const OptixAccelEmitDesc*)emitted, n_emitted));
    *handle = h;
    return 0;
}

int oref_accel_compact(void* stream, unsigned long long in, unsigned long long out, size_t out_bytes, unsigned long long* handle)
{
    OptixTraversableHandle h = 0;
    OCHK(optixAccelCompact(g_ctx, (CUstream)stream, (OptixTraversableHandle)in, (CUdeviceptr)out, out_bytes, &h));
    *handle = h;
    return 0;
}

// ---- module + program groups + pipeline ---------------------------------------------------------------
// miss_csv / ch_csv / ah_csv: comma-separated entry names, "-" = none (null entry function).  ch and ah lists pair up.
// payload_semantics != NULL: one typed payload (optixPathTracer.h:51-80) and numPayloadValues = 0 in the pipeline options.
int oref_pipeline_create(const char* ptx_path, const char* raygen, const char* miss_csv, const char* ch_csv, const char* ah_csv,
                         int num_payload_values, const unsigned* payload_semantics, int n_semantics, unsigned traversable_graph_flags,
                         unsigned max_trace_depth, unsigned max_traversable_depth, int* pipeline_id)
{
    std::ifstream f(ptx_path, std::ios::binary);
    if (!f) return fail("open ptx", -2);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string ptx = ss.str();

    Pipeline P;
    OptixPayloadType ptype = {};
    OptixModuleCompileOptions mco = {};
    mco.optLevel = OPTIX_COMPILE_OPTIMIZATION_DEFAULT;
    mco.debugLevel = OPTIX_COMPILE_DEBUG_LEVEL_MINIMAL;
    if (payload_semantics) {
        ptype.numPayloadValues = (unsigned)n_semantics;
        ptype.payloadSemantics = payload_semantics;
        mco.numPayloadTypes = 1;
        mco.payloadTypes = &ptype;
    }
    OptixPipelineCompileOptions pco = {};
    pco.usesMotionBlur = false;
    pco.traversableGraphFlags = traversable_graph_flags;
    pco.numPayloadValues = payload_semantics ? 0 : num_payload_values;
    pco.numAttributeValues = 2;
    pco.exceptionFlags = OPTIX_EXCEPTION_FLAG_NONE;
    pco.pipelineLaunchParamsVariableName = "params";

    char log[4096];
    size_t log_size = sizeof log;
    OptixResult r = optixModuleCreate(g_ctx, &mco, &pco, ptx.data(), ptx.size(), log, &log_size, &P.module);
    if (log_size > 1) g_log += std::string("module: ") + log + "\n";
    if (r != OPTIX_SUCCESS) return fail("optixModuleCreate", r);

    OptixProgramGroupOptions pgo = {};
    std::vector<OptixProgramGroup> all;
    {
        OptixProgramGroupDesc d = {};
        d.kind = OPTIX_PROGRAM_GROUP_KIND_RAYGEN;
        d.raygen.module = P.module;
        d.raygen.entryFunctionName = raygen;
        log_size = sizeof log;
        OCHK(optixProgramGroupCreate(g_ctx, &d, 1, &pgo, log, &log_size, &P.raygen));
        all.push_back(P.raygen);
    }
    for (const std::string& name : split(miss_csv)) {
        OptixProgramGroupDesc d = {};
        d.kind = OPTIX_PROGRAM_GROUP_KIND_MISS;
        if (name != "-") { d.miss.module = P.module; d.miss.entryFunctionName = name.c_str(); }
        OptixProgramGroup g = nullptr;
        log_size = sizeof log;
        OCHK(optixProgramGroupCreate(g_ctx, &d, 1, &pgo, log, &log_size, &g));
        P.miss.push_back(g);
        all.push_back(g);
    }
    const std::vector<std::string> ch = split(ch_csv), ah = split(ah_csv);
    for (size_t i = 0; i < ch.size(); ++i) {
        OptixProgramGroupDesc d = {};
        d.kind = OPTIX_PROGRAM_GROUP_KIND_HITGROUP;
        if (ch[i] != "-") { d.hitgroup.moduleCH = P.module; d.hitgroup.entryFunctionNameCH = ch[i].c_str(); }
        if (i < ah.size() && ah[i] != "-") { d.hitgroup.moduleAH = P.module; d.hitgroup.entryFunctionNameAH = ah[i].c_str(); }
        OptixProgramGroup g = nullptr;
        log_size = sizeof log;
        OCHK(optixProgramGroupCreate(g_ctx, &d, 1, &pgo, log, &log_size, &g));
        P.hit.push_back(g);
        all.push_back(g);
    }

    OptixPipelineLinkOptions plo = {};
    plo.maxTraceDepth = max_trace_depth;
    log_size = sizeof log;
    r = optixPipelineCreate(g_ctx, &pco, &plo, all.data(), (unsigned)all.size(), log, &log_size, &P.pipeline);
    if (log_size > 1) g_log += std::string("pipeline: ") + log + "\n";
    if (r != OPTIX_SUCCESS) return fail("optixPipelineCreate", r);

    OptixStackSizes stack = {};
    for (auto g : all) OCHK(optixUtilAccumulateStackSizes(g, &stack, P.pipeline));
    unsigned dc_trav = 0, dc_state = 0, cont = 0;
    OCHK(optixUtilComputeStackSizes(&stack, max_trace_depth, 0, 0, &dc_trav, &dc_state, &cont));
    OCHK(optixPipelineSetStackSize(P.pipeline, dc_trav, dc_state, cont, max_traversable_depth));

    g_pipelines.push_back(P);
    *pipeline_id = (int)g_pipelines.size() - 1;
    return 0;
}

// kind: 0 raygen, 1 miss, 2 hit group.  Writes the 32-byte SBT record header at `record` (host memory).
int oref_pack_header(int pipeline_id, int kind, int index, void* record)
{
    if (pipeline_id < 0 || pipeline_id >= (int)g_pipelines.size()) return fail("pipeline id", -3);
    Pipeline& P = g_pipelines[pipeline_id];
    OptixProgramGroup g = nullptr;
    if (kind == 0) g = P.raygen;
    else if (kind == 1 && index < (int)P.miss.size()) g = P.miss[index];
    else if (kind == 2 && index < (int)P.hit.size()) g = P.hit[index];
    if (!g) return fail("program group index", -4);
    OCHK(optixSbtRecordPackHeader(g, record));
    return 0;
}

int oref_launch(int pipeline_id, void* stream, unsigned long long d_params, size_t params_size, const void* sbt, unsigned w, unsigned h, unsigned d)
{
    if (pipeline_id < 0 || pipeline_id >= (int)g_pipelines.size()) return fail("pipeline id", -3);
    OCHK(optixLaunch(g_pipelines[pipeline_id].pipeline, (CUstream)stream, (CUdeviceptr)d_params, params_size, (const OptixShaderBindingTable*)sbt, w, h, d));
    return 0;
}

}  // extern "C"
