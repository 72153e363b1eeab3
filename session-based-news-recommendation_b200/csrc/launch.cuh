// Programmatic dependent launch (PDL) for the chains of short kernels of the train / eval step: every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with griddepcontrol.wait, which returns
// only when the preceding kernel of the stream has completed and its writes are visible -- the stream order is
// preserved exactly, but the launch latency and the CTA set-up of a kernel overlap the tail of its predecessor.
// -DTCAR_NO_PDL restores plain launches.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace tcar {

#ifdef TCAR_NO_PDL
#define PDL_ENTER() ((void)0)
#else
#define PDL_ENTER()                                                        \
    do {                                                                   \
        asm volatile("griddepcontrol.wait;" ::: "memory");                 \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    \
    } while (0)
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
#ifndef TCAR_NO_PDL
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#endif
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tcar
