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
        // graph executions in flight are not terminated by destroying their objects; the resources go when the execution completes
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

    int launch(cudaStream_t s)
    {
        B2_CUDA(ctx_, cudaGraphInstantiate(&exec_, graph_, 0));
        B2_CUDA(ctx_, cudaGraphLaunch(exec_, s));
        return 0;
    }
    unsigned kernels_per_round() const { return kernels_; }

private:
    b200rt_context ctx_;
    cudaGraph_t graph_ = nullptr, body_ = nullptr;
    cudaGraphExec_t exec_ = nullptr;
    cudaGraphConditionalHandle cond_ = 0;
    cudaGraphNode_t prev_ = nullptr;
    unsigned kernels_ = 0;
};

}  // namespace b200rt
