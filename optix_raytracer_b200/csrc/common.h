// common.h — context object, error plumbing and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200rt.h"

namespace b200rt {

struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
};
struct PTGraphCache;
class LoopGraph;
struct WLoopCache;

}  // namespace b200rt

struct b200rt_context_t {
    int device = 0;
    int sm_count = 148;
    size_t l2_bytes = 0;
    b200rt_log_cb log_cb = nullptr;
    void* log_data = nullptr;
    int log_level = 0;
    std::string last_error;
    // wavefront workspace (lane state, queues, counters); grown on demand, owned by the context
    b200rt::Workspace ws;
    void* pinned = nullptr;   // small pinned host block for counter read-back and the asynchronous error flags the kernels write
    // Launches that keep their state in `ws` are ordered by their stream; a launch on ANOTHER stream first waits for `ev`, recorded at
    // the end of the last launch that used the workspace (ws_acquire / ws_release): two streams never see the lane state of each other.
    cudaEvent_t ev = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_busy = false;
    std::vector<b200rt::LoopGraph*> loops;      // build loops in flight (loop_graph.h)
    b200rt::WLoopCache* w_loop = nullptr;       // the BLEND level loop of the whitted launches (whitted.cu)
    b200rt::PTGraphCache* pt_graphs = nullptr;  // instantiated wavefront-loop graphs of the path-tracer launches (pathtracer.cu)
    std::recursive_mutex mu;  // held by every C ABI entry point that takes the context (CTX_CHECK): last_error, the workspace and the caches are per context
    std::vector<cudaEvent_t> timing_events;  // pool for B200RT_PT_STATS_TIMING
    unsigned int counter_slot = 0;  // rotating fetch-counter slot for persistent ray launches
    // whitted launches: light count the workspace was sized for, learnt from the first launch with a given d_params (whitted.cu)
    uint64_t w_params = 0, w_sbt = 0;
    unsigned int w_lights = 0, w_sbt_count = 0;
    bool w_blend = false;     // the hit-group records hold an ALPHA_MODE_BLEND material (continuation levels are run)
    // imgui_test launches: samples_per_frame / nlights of the Params block at pg_params, read on the first launch with it (playground.cu)
    uint64_t pg_params = 0;
    unsigned int pg_spf = 0;
    int pg_nlights = 0;
    uint64_t launches = 0;    // kernels launched through this context (bench: gpu_launches)
};

namespace b200rt {

// grid of a grid-stride kernel: what the work needs, at most ctas_per_sm resident CTAs on every SM
inline unsigned persistent_grid(b200rt_context ctx, uint64_t n, int block, int ctas_per_sm)
{
    const uint64_t need = (n + block - 1) / block;
    const uint64_t cap = (uint64_t)ctx->sm_count * ctas_per_sm;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}

int set_error(b200rt_context ctx, int code, const char* fmt, ...);
void log_msg(b200rt_context ctx, int level, const char* tag, const char* fmt, ...);
int ensure_workspace(b200rt_context ctx, size_t bytes, cudaStream_t stream);
// bracket every launch whose kernels read or write ctx->ws (caller holds ctx->mu)
inline void ws_acquire(b200rt_context ctx, cudaStream_t s)
{
    if (ctx->ws_busy && ctx->ws_stream != s) cudaStreamWaitEvent(s, ctx->ev, 0);
}
inline void ws_release(b200rt_context ctx, cudaStream_t s)
{
    cudaEventRecord(ctx->ev, s);
    ctx->ws_stream = s;
    ctx->ws_busy = true;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define B2_CUDA(ctx, expr)                                                                                  \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return b200rt::set_error(ctx, B200RT_ERROR_CUDA_ERROR, "%s failed: %s (%s:%d)", #expr,          \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);                           \
    } while (0)

#define B2_LAUNCH_CHECK(ctx)                                                                                \
    do {                                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                                \
        if (_e != cudaSuccess)                                                                              \
            return b200rt::set_error(ctx, B200RT_ERROR_LAUNCH_FAILURE, "kernel launch failed: %s (%s:%d)",  \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);                           \
        (ctx)->launches++;                                                                                  \
    } while (0)

#define B2_REQUIRE(ctx, cond, ...)                                                                          \
    do {                                                                                                    \
        if (!(cond)) return b200rt::set_error(ctx, B200RT_ERROR_INVALID_VALUE, __VA_ARGS__);                \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline unsigned int div_up(size_t n, size_t d) { return (unsigned int)((n + d - 1) / d); }

#ifdef __CUDACC__
// Programmatic dependent launch for chains of short kernels on one stream (the viewer's frame: seven launches for 0.22 ms).  A kernel
// launched with launch_chain may be made resident while its predecessor is still running; it calls chain_enter() first thing, which
// (a) lets ITS successor do the same and (b) waits until the predecessor has finished and its writes are visible.  Every kernel of a
// chain must call chain_enter() — also one that returns at once — because a kernel's wait only covers the kernel right before it; the
// ones further back are covered by that kernel having waited in turn.  Without the launch attribute both instructions do nothing.
// B200RT_PDL=0 launches the chains the plain way.
__device__ __forceinline__ void chain_enter()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool chain_launches_enabled()
{
    static const bool on = [] { const char* e = getenv("B200RT_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = chain_launches_enabled() ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Release-ordered atomics for "write a result, then count yourself in" protocols.  __threadfence() is fence.sc.gpu, which on sm_100 is
// MEMBAR.SC.GPU + ERRBAR + CCTL.IVALL — it invalidates the SM's WHOLE L1 every time (in a traversal kernel: the BVH nodes of every warp
// on the SM; in the tree builder: 7 GB of keys re-read from DRAM and a kernel that ran 6 ms instead of 1).  A release on the atomic is
// all such a protocol needs from the writer (MEMBAR.ALL.GPU, no invalidation); the reader that takes the last ticket reads the others'
// results past the L1 (__ldcg / volatile), behind the control dependency on the ticket.
__device__ __forceinline__ unsigned int atomic_exch_release(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.release.gpu.global.exch.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int atomic_add_release(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
#endif

}  // namespace b200rt
