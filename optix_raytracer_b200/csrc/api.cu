// api.cu — the C ABI of include/b200rt.h: context management, error reporting, argument checks,
// and host-side sutil mirrors.  No compute happens on the host; every compute entry point needs a
// CUDA device and fails loudly (B200RT_ERROR_CUDA_ERROR) without one.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "accel.h"
#include "internal.h"
#include "loop_graph.h"

// layouts promised by b200rt.h (reference include/optix_types.h)
static_assert(sizeof(b200rt_build_input) == 1032, "OptixBuildInput");
static_assert(sizeof(b200rt_build_input_triangle_array) == 240, "OptixBuildInputTriangleArray");
static_assert(offsetof(b200rt_build_input, triangleArray) == 8, "OptixBuildInput union offset");
static_assert(offsetof(b200rt_build_input_triangle_array, transformFormat) == 92, "OptixBuildInputTriangleArray tail");
static_assert(sizeof(b200rt_instance) == 80, "OptixInstance");
static_assert(sizeof(b200rt_accel_build_options) == 20, "OptixAccelBuildOptions");
static_assert(sizeof(b200rt_accel_buffer_sizes) == 24, "OptixAccelBufferSizes");
static_assert(sizeof(b200rt_accel_emit_desc) == 16, "OptixAccelEmitDesc");
static_assert(sizeof(b200rt_shader_binding_table) == 64, "OptixShaderBindingTable");

namespace b200rt {

int set_error(b200rt_context ctx, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        ctx->last_error = buf;
        if (ctx->log_cb && ctx->log_level >= 2) ctx->log_cb(2, "ERROR", buf, ctx->log_data);
    }
    return code;
}

void log_msg(b200rt_context ctx, int level, const char* tag, const char* fmt, ...)
{
    if (!ctx || !ctx->log_cb || ctx->log_level < level) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    ctx->log_cb((unsigned)level, tag, buf, ctx->log_data);
}

int ensure_workspace(b200rt_context ctx, size_t bytes, cudaStream_t stream)
{
    if (ctx->ws.bytes >= bytes) return 0;
    if (ctx->ws.ptr) {
        B2_CUDA(ctx, cudaStreamSynchronize(stream));
        B2_CUDA(ctx, cudaDeviceSynchronize());
        B2_CUDA(ctx, cudaFree(ctx->ws.ptr));
        ctx->ws.ptr = nullptr;
        ctx->ws.bytes = 0;
    }
    const size_t want = align_up(bytes + bytes / 8, 1 << 20);
    B2_CUDA(ctx, cudaMalloc(&ctx->ws.ptr, want));
    ctx->ws.bytes = want;
    log_msg(ctx, 4, "workspace", "wavefront workspace grown to %zu bytes", want);
    return 0;
}

}  // namespace b200rt

using namespace b200rt;

// every entry point that takes a context holds its mutex for the call: set_error writes ctx->last_error, launches share the workspace
#define CTX_CHECK(ctx)                                        \
    if (!(ctx)) return B200RT_ERROR_INVALID_DEVICE_CONTEXT;   \
    std::lock_guard<std::recursive_mutex> b2_ctx_lock((ctx)->mu)

extern "C" {

const char* b200rt_version(void) { return "b200rt 0.1 (sm_100a)"; }

const char* b200rt_error_name(int code)
{
    switch (code) {
        case B200RT_SUCCESS: return "B200RT_SUCCESS";
        case B200RT_ERROR_INVALID_VALUE: return "B200RT_ERROR_INVALID_VALUE";
        case B200RT_ERROR_HOST_OUT_OF_MEMORY: return "B200RT_ERROR_HOST_OUT_OF_MEMORY";
        case B200RT_ERROR_INVALID_OPERATION: return "B200RT_ERROR_INVALID_OPERATION";
        case B200RT_ERROR_LAUNCH_FAILURE: return "B200RT_ERROR_LAUNCH_FAILURE";
        case B200RT_ERROR_INVALID_DEVICE_CONTEXT: return "B200RT_ERROR_INVALID_DEVICE_CONTEXT";
        case B200RT_ERROR_NOT_SUPPORTED: return "B200RT_ERROR_NOT_SUPPORTED";
        case B200RT_ERROR_CUDA_ERROR: return "B200RT_ERROR_CUDA_ERROR";
        default: return "B200RT_ERROR_UNKNOWN";
    }
}

const char* b200rt_error_string(int code)
{
    switch (code) {
        case B200RT_SUCCESS: return "Success";
        case B200RT_ERROR_INVALID_VALUE: return "Invalid value";
        case B200RT_ERROR_HOST_OUT_OF_MEMORY: return "Host is out of memory";
        case B200RT_ERROR_INVALID_OPERATION: return "Invalid operation";
        case B200RT_ERROR_LAUNCH_FAILURE: return "Launch failure";
        case B200RT_ERROR_INVALID_DEVICE_CONTEXT: return "Invalid device context";
        case B200RT_ERROR_NOT_SUPPORTED: return "Not supported";
        case B200RT_ERROR_CUDA_ERROR: return "Error during CUDA call";
        default: return "Unknown error";
    }
}

const char* b200rt_last_error_message(b200rt_context ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int b200rt_context_create(int cuda_device, b200rt_log_cb cb, void* cbdata, int level, b200rt_context* out)
{
    if (!out) return B200RT_ERROR_INVALID_VALUE;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || cuda_device < 0 || cuda_device >= count) {
        if (cb && level >= 1) cb(1, "FATAL", "no usable CUDA device: b200rt has no CPU path", cbdata);
        return B200RT_ERROR_CUDA_ERROR;
    }
    b200rt_context ctx = new (std::nothrow) b200rt_context_t();
    if (!ctx) return B200RT_ERROR_HOST_OUT_OF_MEMORY;
    ctx->device = cuda_device;
    ctx->log_cb = cb;
    ctx->log_data = cbdata;
    ctx->log_level = level;
    DeviceGuard guard(cuda_device);
    cudaDeviceProp prop;
    if (cudaFree(0) != cudaSuccess || cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess ||
        cudaMallocHost(&ctx->pinned, 4096) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return B200RT_ERROR_CUDA_ERROR;
    }
    memset(ctx->pinned, 0, 4096);  // counter read-back area + the asynchronous launch-error flags the kernels write (whitted.cu)
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    log_msg(ctx, 4, "context", "device %d: %s, %d SMs, L2 %zu MiB", cuda_device, prop.name, ctx->sm_count, ctx->l2_bytes >> 20);
    *out = ctx;
    return 0;
}

int b200rt_context_destroy(b200rt_context ctx)
{
    CTX_CHECK(ctx);
    {
        DeviceGuard guard(ctx->device);
        cudaDeviceSynchronize();
        pathtracer_release(ctx);
        whitted_release(ctx);
        retire_loops(ctx, true);
        if (ctx->ws.ptr) cudaFree(ctx->ws.ptr);
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        if (ctx->ev) cudaEventDestroy(ctx->ev);
        for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    }
    delete ctx;
    return 0;
}

uint64_t b200rt_context_kernel_launches(b200rt_context ctx) { return ctx ? ctx->launches + pathtracer_graph_kernels(ctx) : 0; }

int b200rt_shared_buffer_create(b200rt_context ctx, size_t bytes, b200rt_deviceptr* ptr, unsigned char* handle64)
{
    CTX_CHECK(ctx);
    B2_REQUIRE(ctx, bytes && ptr && handle64, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    DeviceGuard guard(ctx->device);
    void* p = nullptr;
    B2_CUDA(ctx, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "shared buffer: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *ptr = (b200rt_deviceptr)p;
    return B200RT_SUCCESS;
}

int b200rt_shared_buffer_open(b200rt_context ctx, const unsigned char* handle64, b200rt_deviceptr* ptr)
{
    CTX_CHECK(ctx);
    B2_REQUIRE(ctx, handle64 && ptr, "bad argument");
    DeviceGuard guard(ctx->device);  // opened by the device whose launches will store through it: peer access is set up for this pair
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    B2_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = (b200rt_deviceptr)p;
    return B200RT_SUCCESS;
}

int b200rt_shared_buffer_close(b200rt_context ctx, b200rt_deviceptr ptr)
{
    CTX_CHECK(ctx);
    DeviceGuard guard(ctx->device);
    if (ptr) B2_CUDA(ctx, cudaIpcCloseMemHandle((void*)ptr));
    return B200RT_SUCCESS;
}

int b200rt_shared_buffer_destroy(b200rt_context ctx, b200rt_deviceptr ptr)
{
    CTX_CHECK(ctx);
    DeviceGuard guard(ctx->device);
    if (ptr) B2_CUDA(ctx, cudaFree((void*)ptr));
    return B200RT_SUCCESS;
}

int b200rt_enable_peer_access(b200rt_context ctx, int peer_device)
{
    CTX_CHECK(ctx);
    if (peer_device == ctx->device) return B200RT_SUCCESS;
    DeviceGuard guard(ctx->device);
    int can = 0;
    B2_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return set_error(ctx, B200RT_ERROR_NOT_SUPPORTED, "GPU %d has no peer access to GPU %d", ctx->device, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return B200RT_SUCCESS; }
    if (e != cudaSuccess) return set_error(ctx, B200RT_ERROR_CUDA_ERROR, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    return B200RT_SUCCESS;
}

int b200rt_accel_compute_memory_usage(b200rt_context ctx, const b200rt_accel_build_options* options, const b200rt_build_input* inputs,
                                      unsigned int num_inputs, b200rt_accel_buffer_sizes* sizes)
{
    CTX_CHECK(ctx);
    return accel_compute_memory_usage(ctx, options, inputs, num_inputs, sizes);
}

int b200rt_accel_build(b200rt_context ctx, b200rt_stream stream, const b200rt_accel_build_options* options, const b200rt_build_input* inputs,
                       unsigned int num_inputs, b200rt_deviceptr temp_buffer, size_t temp_bytes, b200rt_deviceptr output_buffer,
                       size_t output_bytes, b200rt_traversable* handle, const b200rt_accel_emit_desc* emitted, unsigned int num_emitted)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return accel_build(ctx, (cudaStream_t)stream, options, inputs, num_inputs, temp_buffer, temp_bytes, output_buffer, output_bytes, handle,
                       emitted, num_emitted);
}

int b200rt_accel_compact(b200rt_context ctx, b200rt_stream stream, b200rt_traversable input, b200rt_deviceptr output_buffer,
                         size_t output_bytes, b200rt_traversable* handle)
{
    CTX_CHECK(ctx);
    return accel_compact(ctx, (cudaStream_t)stream, input, output_buffer, output_bytes, handle);
}

int b200rt_accel_emit_property(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, const b200rt_accel_emit_desc* emitted,
                               unsigned int num_emitted)
{
    CTX_CHECK(ctx);
    return accel_emit_property(ctx, (cudaStream_t)stream, handle, emitted, num_emitted);
}

int b200rt_accel_get_info(b200rt_context ctx, b200rt_traversable handle, b200rt_accel_info* info)
{
    CTX_CHECK(ctx);
    return accel_get_info(ctx, handle, info);
}

int b200rt_launch_pathtracer(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt,
                             unsigned int width, unsigned int height, const b200rt_pt_options* options)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return launch_pathtracer(ctx, (cudaStream_t)stream, d_params, sbt, width, height, options, 0);
}

int b200rt_launch_multigpu(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt,
                           unsigned int num_samples, const b200rt_pt_options* options)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return launch_pathtracer(ctx, (cudaStream_t)stream, d_params, sbt, num_samples, 1, options, 1);
}

int b200rt_fill_samples(b200rt_context ctx, b200rt_stream stream, int gpu_idx, int num_gpus, int width, int height,
                        b200rt_deviceptr sample_indices_int2, int num_samples)
{
    CTX_CHECK(ctx);
    return fill_samples(ctx, (cudaStream_t)stream, gpu_idx, num_gpus, width, height, sample_indices_int2, num_samples);
}

int b200rt_deinterleave(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr gathered_float4, int num_gpus, int num_samples, int width,
                        int height, b200rt_deviceptr accum_float4, b200rt_deviceptr frame_uchar4)
{
    CTX_CHECK(ctx);
    return deinterleave(ctx, (cudaStream_t)stream, gathered_float4, num_gpus, num_samples, width, height, accum_float4, frame_uchar4);
}

int b200rt_launch_raycast(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt,
                          unsigned int width, unsigned int height, b200rt_deviceptr ext_hits)
{
    CTX_CHECK(ctx);
    return launch_raycast(ctx, (cudaStream_t)stream, d_params, sbt, width, height, ext_hits);
}

int b200rt_create_rays_ortho(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr rays, int width, int height, const float bbmin[3],
                             const float bbmax[3], float padding)
{
    CTX_CHECK(ctx);
    return create_rays_ortho(ctx, (cudaStream_t)stream, rays, width, height, bbmin, bbmax, padding);
}

int b200rt_translate_rays(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr rays, int count, const float offset[3])
{
    CTX_CHECK(ctx);
    return translate_rays(ctx, (cudaStream_t)stream, rays, count, offset);
}

int b200rt_shade_hits(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr image_float3, int count, b200rt_deviceptr hits)
{
    CTX_CHECK(ctx);
    return shade_hits(ctx, (cudaStream_t)stream, image_float3, count, hits);
}

int b200rt_trace_closest(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n,
                         unsigned int ray_flags, b200rt_deviceptr ext_hits)
{
    CTX_CHECK(ctx);
    return trace_closest(ctx, (cudaStream_t)stream, handle, rays, n, ray_flags, ext_hits);
}

int b200rt_trace_any(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n,
                     unsigned int ray_flags, b200rt_deviceptr occluded_u32)
{
    CTX_CHECK(ctx);
    return trace_any(ctx, (cudaStream_t)stream, handle, rays, n, ray_flags, occluded_u32);
}

int b200rt_trace_stats(b200rt_context ctx, b200rt_stream stream, b200rt_traversable handle, b200rt_deviceptr rays, uint64_t n,
                       uint64_t* nodes_fetched, uint64_t* tris_tested)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return trace_stats(ctx, (cudaStream_t)stream, handle, rays, n, nodes_fetched, tris_tested);
}

int b200rt_generate_synthetic_mesh(b200rt_context ctx, b200rt_stream stream, uint64_t num_triangles, uint32_t seed,
                                   b200rt_deviceptr vertices_float4, b200rt_deviceptr mat_indices_u32, float bounds_out[6])
{
    CTX_CHECK(ctx);
    return generate_synthetic_mesh(ctx, (cudaStream_t)stream, num_triangles, seed, vertices_float4, mat_indices_u32, bounds_out);
}

int b200rt_launch_whitted(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, const b200rt_shader_binding_table* sbt,
                          unsigned int width, unsigned int height)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return launch_whitted(ctx, (cudaStream_t)stream, d_params, sbt, width, height);
}

int b200rt_texture_create(b200rt_context ctx, int width, int height, const void* rgba8, int address_s, int address_t, int linear_filter,
                          uint64_t* texture_object, uint64_t* cuda_array)
{
    CTX_CHECK(ctx);
    return texture_create(ctx, width, height, rgba8, address_s, address_t, linear_filter, texture_object, cuda_array);
}

int b200rt_texture_view(b200rt_context ctx, uint64_t cuda_array, int address_s, int address_t, int linear_filter, uint64_t* texture_object)
{
    CTX_CHECK(ctx);
    return texture_view(ctx, cuda_array, address_s, address_t, linear_filter, texture_object);
}

int b200rt_texture_destroy(b200rt_context ctx, uint64_t texture_object, uint64_t cuda_array)
{
    CTX_CHECK(ctx);
    return texture_destroy(ctx, texture_object, cuda_array);
}

int b200rt_launch_playground(b200rt_context ctx, b200rt_stream stream, b200rt_deviceptr d_params, unsigned int width, unsigned int height,
                             const b200rt_pt_options* options)
{
    CTX_CHECK(ctx);
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    return launch_playground(ctx, (cudaStream_t)stream, d_params, width, height, options);
}

int b200rt_generate_playground_scene(b200rt_context ctx, b200rt_stream stream, uint32_t rows, uint32_t seed, b200rt_deviceptr vertices_float3,
                                     b200rt_deviceptr normals_float3, b200rt_deviceptr mat_indices_i32, uint64_t* num_triangles)
{
    CTX_CHECK(ctx);
    return generate_playground_scene(ctx, (cudaStream_t)stream, rows, seed, vertices_float3, normals_float3, mat_indices_i32, num_triangles);
}

// ---- host-side sutil mirrors -------------------------------------------------------------------
// sutil::Camera::UVWFrame (reference SDK/sutil/Camera.cpp:34-46), host arithmetic without contraction
void b200rt_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fov_y_deg, float aspect, float U[3], float V[3],
                       float W[3])
{
    struct v3 { float x, y, z; };
    auto dot = [](v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    auto cross = [](v3 a, v3 b) { return v3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
    auto scale = [](v3 a, float s) { return v3{a.x * s, a.y * s, a.z * s}; };
    auto norm = [&](v3 a) { return scale(a, 1.0f / sqrtf(dot(a, a))); };
    v3 w{lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
    const float wlen = sqrtf(dot(w, w));
    v3 u = norm(cross(w, v3{up[0], up[1], up[2]}));
    v3 v = norm(cross(u, w));
    const float vlen = wlen * tanf(0.5f * fov_y_deg * 3.14159265358979323846f / 180.0f);
    v = scale(v, vlen);
    const float ulen = vlen * aspect;
    u = scale(u, ulen);
    U[0] = u.x; U[1] = u.y; U[2] = u.z;
    V[0] = v.x; V[1] = v.y; V[2] = v.z;
    W[0] = w.x; W[1] = w.y; W[2] = w.z;
}

// imgui_test Camera as main.cpp:236-243 sets it up: set_eye / set_up / set_lookat / set_aperture / set_fd (relative to |lookat - eye|) /
// set_fov / set_ortho, then compute_uvw (reference SDK/imgui_test/camera.h:19-117).  Writes the 92-byte object the device programs read.
void b200rt_playground_camera(const float eye[3], const float up[3], const float lookat[3], float aperture, float fd, float fov_deg, int ortho,
                              void* camera92)
{
    struct v3 { float x, y, z; };
    auto dot = [](v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    auto cross = [](v3 a, v3 b) { return v3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
    auto scale = [](v3 a, float s) { return v3{a.x * s, a.y * s, a.z * s}; };
    auto norm = [&](v3 a) { return scale(a, 1.0f / sqrtf(dot(a, a))); };
    const v3 e{eye[0], eye[1], eye[2]}, l{lookat[0], lookat[1], lookat[2]}, upv{up[0], up[1], up[2]};
    const v3 le{l.x - e.x, l.y - e.y, l.z - e.z};
    const float m_fd = fd / sqrtf(dot(le, le));
    v3 w = scale(le, m_fd);
    const float wlen = sqrtf(dot(w, w));
    v3 u = norm(cross(w, upv));
    v3 v = norm(cross(u, w));
    const float vlen = wlen * tanf(0.5f * fov_deg * 3.14159265358979323846f / 180.0f);
    v = scale(v, vlen);
    u = scale(u, vlen);
    struct Cam { float eye[3], lookat[3], up[3]; unsigned char ortho, pad[3]; float fov, fd, aperture, speed; float u[3], v[3], w[3]; } c;
    static_assert(sizeof(Cam) == 92, "imgui_test Camera");
    memset(&c, 0, sizeof c);
    c.eye[0] = e.x; c.eye[1] = e.y; c.eye[2] = e.z; c.lookat[0] = l.x; c.lookat[1] = l.y; c.lookat[2] = l.z;
    c.up[0] = upv.x; c.up[1] = upv.y; c.up[2] = upv.z;
    c.ortho = ortho ? 1 : 0; c.fov = fov_deg; c.fd = m_fd; c.aperture = aperture; c.speed = 0.01f;
    c.u[0] = u.x; c.u[1] = u.y; c.u[2] = u.z; c.v[0] = v.x; c.v[1] = v.y; c.v[2] = v.z; c.w[0] = w.x; c.w[1] = w.y; c.w[2] = w.z;
    memcpy(camera92, &c, sizeof c);
}

// what the triangle tests of an instance see of the ray's flags (accel.h: cull_word)
unsigned int b200rt_triangle_flag_word(unsigned int ray_flags, unsigned int instance_flags) { return b200rt::cull_word(ray_flags, instance_flags); }

// StaticWorkDistribution (reference SDK/sutil/WorkDistribution.h:50-81): 8x4 tiles, strips of 8*N columns
int b200rt_wd_num_samples(int width, int height, int num_gpus)
{
    const int strip_w = 8 * num_gpus;
    const int cols = (width + strip_w - 1) / strip_w, rows = (height + 3) / 4;
    return rows * cols * 32;
}

void b200rt_wd_sample_pixel(int width, int height, int num_gpus, int gpu_idx, int sample_idx, int xy[2])
{
    (void)height;
    const int strip_w = 8 * num_gpus;
    const int cols = (width + strip_w - 1) / strip_w;
    const int tile = sample_idx >> 5, in_tile = sample_idx & 31;
    const int row = tile / cols, col = tile % cols;
    const int rot = (gpu_idx + row % num_gpus) % num_gpus;
    xy[0] = col * strip_w + rot * 8 + (in_tile & 7);
    xy[1] = row * 4 + (in_tile >> 3);
}

}  // extern "C"
