// loop_graph.h — `while (device condition) { kernel; kernel; ... }` as a CUDA graph with a conditional WHILE node, launched into the
// caller's stream: the host never reads a count back to decide whether another round is needed.  The body is a chain of kernel nodes
// with fixed launch shapes (persistent / grid-stride kernels that read their problem size from device memory); its last kernel sets the
// condition with cudaGraphSetConditional(handle, more).  Used by the accel build (clustering rounds, collapse levels); the path tracer
// keeps its own cached variant (pathtracer.cu).
#pragma once
#include "common.h"

namespace b200rt {

class LoopGraph {
public:
    explicit LoopGraph(b200rt_context ctx) : ctx_(ctx) {}
    ~LoopGraph()
    {
        if (done_) cudaEventDestroy(done_);
        if (exec_) cudaGraphExecDestroy(exec_);
        if (graph_) cudaGraphDestroy(graph_);
    }
    LoopGraph(const LoopGraph&) = delete;
    LoopGraph& operator=(const LoopGraph&) = delete;

    // default_value: condition at launch (1 = run the body at least once)
    int begin(unsigned default_value = 1u)
    {
        B2_CUDA(ctx_, cudaGraphCreate(&graph_, 0));
        B2_CUDA(ctx_, cudaGraphConditionalHandleCreate(&cond_, graph_, default_value, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = cond_;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t loop;
        B2_CUDA(ctx_, cudaGraphAddNode(&loop, graph_, nullptr, 0, &np));
        body_ = np.conditional.phGraph_out[0];
        return 0;
    }
    cudaGraphConditionalHandle cond() const { return cond_; }

    // append a kernel to the body (runs after the one appended before it); arguments are copied now
    template <class... Args>
    int add(const void* func, unsigned grid, unsigned block, size_t smem, Args... args)
    {
        void* a[] = {(void*)&args...};
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof(kp));
        kp.func = (void*)func;
        kp.gridDim = dim3(grid);
        kp.blockDim = dim3(block);
        kp.sharedMemBytes = (unsigned)smem;
        kp.kernelParams = a;
        cudaGraphNode_t n;
        B2_CUDA(ctx_, cudaGraphAddKernelNode(&n, body_, prev_ ? &prev_ : nullptr, prev_ ? 1 : 0, &kp));
        prev_ = n;
        ++kernels_;
        return 0;
    }

    // instantiated on the first launch; a loop whose kernels and arguments stay the same can be launched again and again
    int launch(cudaStream_t s)
    {
        if (!exec_) B2_CUDA(ctx_, cudaGraphInstantiate(&exec_, graph_, 0));
        B2_CUDA(ctx_, cudaGraphLaunch(exec_, s));
        if (!done_) B2_CUDA(ctx_, cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
        B2_CUDA(ctx_, cudaEventRecord(done_, s));
        return 0;
    }
    // has the launched loop finished?  (a loop that was never launched counts as finished)
    bool finished() const { return !done_ || cudaEventQuery(done_) == cudaSuccess; }
    unsigned kernels_per_round() const { return kernels_; }

private:
    b200rt_context ctx_;
    cudaGraph_t graph_ = nullptr, body_ = nullptr;
    cudaGraphExec_t exec_ = nullptr;
    cudaGraphConditionalHandle cond_ = 0;
    cudaGraphNode_t prev_ = nullptr;
    cudaEvent_t done_ = nullptr;
    unsigned kernels_ = 0;
};

// Loops that may still be running are parked in the context and destroyed once their completion event has fired (destroying an
// executable graph that is in flight makes the host wait for it): retire_loops() at the start of every build, and at context destruction.
void retire_loops(b200rt_context ctx, bool all);
void park_loop(b200rt_context ctx, LoopGraph* g);

}  // namespace b200rt
