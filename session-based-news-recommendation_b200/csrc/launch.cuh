// Programmatic dependent launch (PDL) for the chains of short kernels of the train / eval step: every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with griddepcontrol.wait, which returns
// only when the preceding kernel of the stream has completed and its writes are visible -- the stream order is
// preserved exactly, but the launch latency and the CTA set-up of a kernel overlap the tail of its predecessor.
// -DTCAR_NO_PDL restores plain launches.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and device instead of before every launch.
#define TCAR_SET_SMEM_ONCE(kernel, bytes)                                                                      \
    do {                                                                                                       \
        static int tcar_smem_dev_ = -1;                                                                        \
        int tcar_dev_now_ = -1;                                                                                \
        cudaGetDevice(&tcar_dev_now_);                                                                         \
        if (tcar_dev_now_ != tcar_smem_dev_) {                                                                 \
            const cudaError_t tcar_e_ =                                                                        \
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);              \
            if (tcar_e_ != cudaSuccess) return (int)tcar_e_;                                                   \
            tcar_smem_dev_ = tcar_dev_now_;                                                                    \
        }                                                                                                      \
    } while (0)

}  // namespace tcar

// ---------------------------------------------------------------------------------------------- tensor-map cache
// cuTensorMapEncodeTiled costs a few microseconds and a train step encodes ~100 maps (three per GEMM operand segment
// of the grouped projection launches), always over the same persistent buffers: memoise by the complete argument list.
// A descriptor depends on nothing else, so a hit is exact.  Fixed-size table, overwritten on collision.
#ifdef CUDA_VERSION
#include <mutex>
#include <cstring>
namespace tcar {
struct alignas(64) TmapSlot {
    CUtensorMap map;
    unsigned long long key[16];
    int used;
};
template <typename EncodeFn>
static inline CUresult tmap_encode_cached(EncodeFn fn, CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                                          const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                          const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw,
                                          CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    constexpr int kSlots = 2048;
    static TmapSlot* table = nullptr;
    static std::mutex mu;
    unsigned long long key[16] = {};
    key[0] = (unsigned long long)dt | ((unsigned long long)rank << 8) | ((unsigned long long)il << 16) |
             ((unsigned long long)sw << 24) | ((unsigned long long)l2 << 32) | ((unsigned long long)oob << 40);
    key[1] = (unsigned long long)reinterpret_cast<uintptr_t>(base);
    for (cuuint32_t i = 0; i < rank && i < 5; ++i) {
        key[2 + i] = dims[i];
        key[7 + i] = ((unsigned long long)box[i] << 32) | estr[i];
        if (i + 1 < rank) key[12 + i] = strides[i];
    }
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) { h ^= key[i]; h *= 1099511628211ull; }
    const int at = (int)((h >> 20) % kSlots);
    std::lock_guard<std::mutex> lock(mu);
    if (!table) table = new TmapSlot[kSlots]();
    TmapSlot& s = table[at];
    if (s.used && std::memcmp(s.key, key, sizeof(key)) == 0) {
        std::memcpy(out, &s.map, sizeof(CUtensorMap));
        return CUDA_SUCCESS;
    }
    const CUresult r = fn(out, dt, rank, base, dims, strides, box, estr, il, sw, l2, oob);
    if (r == CUDA_SUCCESS) {
        std::memcpy(&s.map, out, sizeof(CUtensorMap));
        std::memcpy(s.key, key, sizeof(key));
        s.used = 1;
    }
    return r;
}
}  // namespace tcar
#endif
