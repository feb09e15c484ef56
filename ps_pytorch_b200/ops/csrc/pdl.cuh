// Programmatic dependent launch (PDL) for the kernels of the training step.
//
// A step is ~250 dependent kernels of 5..100 us each; between two of them the GPU idles for the launch latency of the next grid plus its
// prologue (barrier init, TMEM allocation, tensor-map prefetch, coefficient loads).  With PDL the next grid may become resident while the
// previous one is still draining: every converted kernel
//   * calls pdl_trigger() first thing  -> its successor (if launched with the attribute) may be scheduled as soon as SM resources allow;
//   * calls pdl_wait() after its prologue, before the FIRST access to global memory that a predecessor may have written (or that a
//     predecessor may still be reading, for buffers this kernel writes) -> blocks until all predecessor grids have completed and flushed.
// Both instructions are no-ops for a grid launched without the attribute, so a kernel is safe under either launch mode, and a kernel that was
// NOT converted is launched the ordinary way (full stream order).  Inside a stream capture the attribute becomes a programmatic edge of the graph.
// MEASURED (profiles/run30_pdl.md): correct (whole GPU suite green with it on) but 10 % SLOWER for the ResNet-18 step, 6.87 vs 6.24 ms: the step hides its
// weight gradients on a low-priority side stream inside the memory-bound BatchNorm kernels, and grids parked behind griddepcontrol.wait occupy exactly the
// thread / shared-memory slots those kernels need.  The attribute is therefore OFF by default; PSB200_PDL=1 (ops/__init__.py -> set_pdl) turns it on.
#pragma once
#include <cuda_runtime.h>

namespace psb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int pdl_enabled();            // bindings.cpp
void pdl_set(int on);

// fills the PDL attribute into `at[n]` when enabled; returns the new attribute count
inline int pdl_attr(cudaLaunchAttribute* at, int n) {
    if (!pdl_enabled()) return n;
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    return n + 1;
}

// <<<grid, block, smem, stream>>> with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    cfg.numAttrs = pdl_attr(at, 0);
    cfg.attrs = at;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace psb
