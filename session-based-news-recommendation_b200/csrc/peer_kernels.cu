// Peer-memory plumbing of the catalog-sharded train step: every rank owns a contiguous range
// of item-table rows (fp32 master copy + Adam moments); the rows a rank's sessions gather (clicks, labels, negatives)
// are read straight out of the owner's HBM over NVLink / NVSwitch with plain loads -- no staging buffer, no
// collective, ~1 KB per row.  The reference has no multi-device code; this replaces what would otherwise be an
// all-gather of the whole updated table (364 MB per step) after model_combine.py:163.
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string.h>
#include "tcar_b200.h"
#include "launch.cuh"

namespace tcar {

struct PeerTable {
    const float* base[TCAR_MAX_PEERS];   // table origin (row 0) in every rank's address space mapping
    int bound[TCAR_MAX_PEERS + 1];       // rank g owns rows [bound[g], bound[g+1])
    int G, self;
};

// one warp per listed row: 2 x 128-bit loads per lane from the owner (L1 bypassed: the owner rewrites its rows every
// step), 2 x 128-bit stores into the local replica.  Rows the caller owns, rows outside the table and duplicates (same
// bytes written twice) need no special care.
__global__ void __launch_bounds__(256)
peer_fetch_rows_kernel(const int32_t* __restrict__ rows, int n, int row_add, const PeerTable pt,
                       float* __restrict__ table) {
    PDL_ENTER();
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n) return;
    const int row = rows[e] + row_add;
    if (row < pt.bound[0] || row >= pt.bound[pt.G]) return;
    int owner = 0;
    while (row >= pt.bound[owner + 1]) ++owner;
    if (owner == pt.self) return;
    const float4* src = reinterpret_cast<const float4*>(pt.base[owner] + (size_t)row * TCAR_HP);
    float4* dst = reinterpret_cast<float4*>(table + (size_t)row * TCAR_HP);
    const float4 v0 = __ldcg(src + lane), v1 = __ldcg(src + 32 + lane);
    dst[lane] = v0;
    dst[32 + lane] = v1;
}

typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static AddrRangeFn addr_range_fn() {
    static AddrRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<AddrRangeFn>(p);
    }
    return fn;
}

}  // namespace tcar

using namespace tcar;

extern "C" int tcar_peer_export(const void* ptr, unsigned char* handle, long long* offset) {
    if (!ptr || !handle || !offset) return TCAR_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == TCAR_PEER_HANDLE_BYTES, "IPC handle size");
    AddrRangeFn fn = addr_range_fn();
    if (!fn) return TCAR_ERR_DRIVER;
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return TCAR_ERR_DRIVER;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
    if (e != cudaSuccess) return (int)e;
    memcpy(handle, &h, sizeof(h));
    *offset = (long long)(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return 0;
}

extern "C" int tcar_peer_open(const unsigned char* handle, long long offset, void** ptr) {
    if (!handle || !ptr || offset < 0) return TCAR_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *ptr = static_cast<char*>(base) + offset;
    return 0;
}

extern "C" int tcar_peer_close(void* ptr, long long offset) {
    if (!ptr) return TCAR_ERR_ARG;
    return (int)cudaIpcCloseMemHandle(static_cast<char*>(ptr) - offset);
}

extern "C" int tcar_peer_fetch_rows(const int32_t* rows, int n, int row_add, const void* const* peers,
                                    const int32_t* row_bounds, int G, int self, float* table, void* stream) {
    if (n < 0 || G < 1 || G > TCAR_MAX_PEERS || self < 0 || self >= G || !peers || !row_bounds || !table)
        return TCAR_ERR_ARG;
    if (n == 0) return 0;
    if (!rows) return TCAR_ERR_ARG;
    PeerTable pt = {};
    for (int g = 0; g < G; ++g) {
        if (!peers[g] || row_bounds[g + 1] < row_bounds[g]) return TCAR_ERR_ARG;
        pt.base[g] = static_cast<const float*>(peers[g]);
    }
    for (int g = 0; g <= G; ++g) pt.bound[g] = row_bounds[g];
    pt.G = G;
    pt.self = self;
    launch_pdl(peer_fetch_rows_kernel, dim3((n + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), rows, n,
               row_add, pt, table);
    return (int)cudaGetLastError();
}
