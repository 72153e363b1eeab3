// Candidate-matrix build, per-tensor gradient norms and the fused clip_by_norm + TF-Adam update.
// Reference: model_combine.py:86-92,135-136 (candidate matrix), :155-163 (AdamOptimizer + clip_by_norm).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include "tcar_b200.h"
#include "launch.cuh"

namespace tcar {

constexpr int H = TCAR_H, HP = TCAR_HP, KEXT = TCAR_KEXT;
__device__ __constant__ int kBinOffO[6] = {0, 13, 45, 53, 78, 139};

// Iext row n = [item[n+1, :250] | content[n+1, :250] | one-hot bins | 0]; one warp per row, 16-byte stores.
__global__ void __launch_bounds__(256)
build_iext_kernel(const float* __restrict__ item, const float* __restrict__ content,
                  const int32_t* __restrict__ mwdhm, __nv_bfloat16* __restrict__ iext, int N, int n_pad) {
    PDL_ENTER();
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= n_pad) return;
    uint4* dst = reinterpret_cast<uint4*>(iext + (size_t)n * KEXT);  // 80 x 16 B per row
    int bins[5] = {-1, -1, -1, -1, -1};
    if (n < N)
        for (int k = 0; k < 5; ++k) bins[k] = 2 * H + kBinOffO[k] + mwdhm[(size_t)n * 5 + k];
    for (int g = lane; g < KEXT / 8; g += 32) {
        __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            float x = 0.f;
            if (n < N) {
                if (c < H) x = item[((size_t)n + 1) * HP + c];
                else if (c < 2 * H) x = content[((size_t)n + 1) * HP + (c - H)];
                else x = (c == bins[0] || c == bins[1] || c == bins[2] || c == bins[3] || c == bins[4]) ? 1.f : 0.f;
            }
            v[j] = __float2bfloat16(x);
        }
        dst[g] = *reinterpret_cast<uint4*>(v);
    }
}

// TCAR_NORM_SPLIT CTAs per segment (tensor) of a flat fp32 buffer: each reduces a contiguous slice in a fixed order
// and writes one partial; the consumer (adam_small_kernel) adds the partials in index order.
constexpr int kSplit = TCAR_NORM_SPLIT;
__global__ void __launch_bounds__(512)
sqnorm_segments_kernel(const float* __restrict__ flat, const int32_t* __restrict__ seg_off, float* __restrict__ out) {
    PDL_ENTER();
    __shared__ float red[16];
    const int s = blockIdx.y, j = blockIdx.x;
    const int lo = seg_off[s], hi = seg_off[s + 1];           // 16-byte aligned starts (params.py pads to 4 floats)
    const int n4 = (hi - lo) >> 2;                            // pads are zero, so whole float4s are safe
    const int per = (n4 + kSplit - 1) / kSplit;
    const int a = j * per, b = min(a + per, n4);
    const float4* x = reinterpret_cast<const float4*>(flat + lo);
    float acc0 = 0.f, acc1 = 0.f;
    int i = a + threadIdx.x;
    for (; i + 512 < b; i += 1024) {
        const float4 u = x[i], v = x[i + 512];
        acc0 += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
        acc1 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (i < b) {
        const float4 u = x[i];
        acc0 += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    }
    float acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 16; ++k) t += red[k];
        out[s * kSplit + j] = t;
    }
}

// All clip norms of one training step in one launch, plus the step counter: rows 0..nseg-1 of the grid are
// sqnorm_segments_kernel; row nseg reduces the item-gradient norm from the per-CTA sums of the dense gradient GEMM (a)
// and the per-row corrections of the scatter (b) -- kSplit CTAs write one partial each and the CTA that arrives last
// adds the partials in index order (result independent of the arrival order); thread 0 of CTA (0, 0) increments the
// step counter, which no thread of this kernel reads.
__global__ void __launch_bounds__(512)
update_norms_kernel(const float* __restrict__ flat, const int32_t* __restrict__ seg_off, float* __restrict__ out_small,
                    int nseg, const float* __restrict__ a, int na, const float4* __restrict__ b, int nb4,
                    float* __restrict__ out_item, float* __restrict__ item_part, int32_t* __restrict__ ticket,
                    int32_t* __restrict__ step) {
    PDL_ENTER();
    __shared__ float red[16];
    __shared__ int s_last;
    const int s = blockIdx.y, j = blockIdx.x;
    float acc0 = 0.f, acc1 = 0.f;
    if (s < nseg) {
        const int lo = seg_off[s], hi = seg_off[s + 1];
        const int n4 = (hi - lo) >> 2;
        const int per = (n4 + kSplit - 1) / kSplit;
        const int p0 = j * per, p1 = min(p0 + per, n4);
        const float4* x = reinterpret_cast<const float4*>(flat + lo);
        int i = p0 + threadIdx.x;
        for (; i + 512 < p1; i += 1024) {
            const float4 u = x[i], v = x[i + 512];
            acc0 += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
            acc1 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        if (i < p1) {
            const float4 u = x[i];
            acc0 += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
        }
    } else {
        if (j == 0)
            for (int i = threadIdx.x; i < na; i += 512) acc0 += a[i];
        const int per = (nb4 + kSplit - 1) / kSplit;
        const int p0 = j * per, p1 = min(p0 + per, nb4);
        int i = p0 + threadIdx.x;
        for (; i + 512 < p1; i += 1024) {
            const float4 u = b[i], v = b[i + 512];
            acc0 += (u.x + u.y) + (u.z + u.w);
            acc1 += (v.x + v.y) + (v.z + v.w);
        }
        if (i < p1) {
            const float4 u = b[i];
            acc0 += (u.x + u.y) + (u.z + u.w);
        }
    }
    float acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x != 0) return;
    float t = 0.f;
    for (int k = 0; k < 16; ++k) t += red[k];
    if (s < nseg) {
        out_small[s * kSplit + j] = t;
        if (s == 0 && j == 0 && step) step[0] += 1;
        return;
    }
    item_part[j] = t;
    __threadfence();
    if (atomicAdd(ticket, 1) == kSplit - 1) {
        __threadfence();
        float tot = 0.f;
        for (int k = 0; k < kSplit; ++k) tot += __ldcg(item_part + k);
        out_item[0] = tot;
        *ticket = 0;
    }
}

// item columns of Iext from the fp32 item table (rows owned by OTHER ranks after the sharded Adam + all-gather)
__global__ void __launch_bounds__(256)
refresh_iext_items_kernel(const float4* __restrict__ item, __nv_bfloat16* __restrict__ iext, long long n4) {
    PDL_ENTER();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> 6;
        const int c = (int)(i & 63) * 4;
        if (row >= 1 && c < H) {
            const float4 p = item[i];
            __nv_bfloat16* dst = iext + (size_t)(row - 1) * KEXT + c;
            *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(p.x, p.y);
            if (c + 2 < H) *reinterpret_cast<__nv_bfloat162*>(dst + 2) = __floats2bfloat162_rn(p.z, p.w);
        }
    }
}

constexpr int kNormBlocks = 1184;  // 8 x 148 SMs
__global__ void __launch_bounds__(256)
sqnorm_big_partial_kernel(const float4* __restrict__ x, float* __restrict__ partial, long long n4) {
    PDL_ENTER();
    __shared__ float red[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024)
sqnorm_big_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int n) {
    PDL_ENTER();
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 32; ++i) t += red[i];
        out[0] = t;
    }
}

// out = sum(a) + sum(b): each thread sums a fixed strided subset (128-bit loads, independent accumulators), then a
// fixed-order block reduction.  nb must be a multiple of 4 and b 16-byte aligned.
__global__ void __launch_bounds__(1024)
sqnorm_combine_kernel(const float* __restrict__ a, int na, const float4* __restrict__ b, int nb4,
                      float* __restrict__ out) {
    PDL_ENTER();
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < na; i += blockDim.x) acc += a[i];
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    int i = threadIdx.x;
    for (; i + 1024 < nb4; i += 2048) {
        const float4 u = b[i], v = b[i + 1024];
        s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
        s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    }
    if (i < nb4) {
        const float4 u = b[i];
        s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
    }
    acc += ((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 32; ++k) t += red[k];
        out[0] = t;
    }
}

// tf.clip_by_norm(g, c) = g * c / max(||g||, c);  TF Adam: lr_t = lr sqrt(1-b2^t)/(1-b1^t); p -= lr_t m/(sqrt(v)+eps)
__device__ __forceinline__ float clip_factor(float sqnorm, float max_grad) {
    const float n = sqrtf(sqnorm);
    return max_grad / fmaxf(n, max_grad);
}
__device__ __forceinline__ float adam_lr_t(int t, float lr) {
    const float b1t = powf(0.9f, (float)t), b2t = powf(0.999f, (float)t);
    return lr * sqrtf(1.f - b2t) / (1.f - b1t);
}
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float lr_t) {
    m = m + (g - m) * (1.f - 0.9f);
    v = v + (g * g - v) * (1.f - 0.999f);
    p -= lr_t * m / (sqrtf(v) + 1e-8f);
}

__global__ void __launch_bounds__(256)
adam_small_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                  const float* __restrict__ g, const int32_t* __restrict__ seg_off,
                  const float* __restrict__ sqnorm, const int32_t* __restrict__ step, float lr, float max_grad) {
    PDL_ENTER();
    const int s = blockIdx.y;
    const int lo = seg_off[s], hi = seg_off[s + 1];
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kSplit; ++j) sq += sqnorm[s * kSplit + j];
    const float cf = clip_factor(sq, max_grad);
    const float lr_t = adam_lr_t(step[0], lr);
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        float p = theta[i], mm = m[i], vv = v[i];
        adam_update(p, mm, vv, g[i] * cf, lr_t);
        theta[i] = p; m[i] = mm; v[i] = vv;
    }
}

// item table [N+1, 256]: one float4 per thread per step; also re-quantises the row into Iext (bf16).
// The six pad columns of the 256-float pitch (float4 slot 63 of every row) are neither read nor written.
// `flags` (optional, one int32 per table row): rows whose flag equals the current step number were already updated by
// adam_item_rows_kernel (the rows the next batch gathers, see Seq2SeqAttNN.train_step) and are skipped here.
__device__ __forceinline__ void store_iext_items(__nv_bfloat16* __restrict__ iext, long long row, int c, const float4 p) {
    if (iext && row >= 1) {       // iext == NULL: the caller rebuilds the scoring operand itself
        __nv_bfloat16* dst = iext + (size_t)(row - 1) * KEXT + c;
        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(p.x, p.y);
        if (c + 2 < H) *reinterpret_cast<__nv_bfloat162*>(dst + 2) = __floats2bfloat162_rn(p.z, p.w);
    }
}

__global__ void __launch_bounds__(256)
adam_item_kernel(float4* __restrict__ item, float4* __restrict__ m, float4* __restrict__ v,
                 const float4* __restrict__ g, const float* __restrict__ sqnorm, const int32_t* __restrict__ step,
                 float lr, float max_grad, __nv_bfloat16* __restrict__ iext, long long n4, long long row0,
                 const int32_t* __restrict__ flags) {
    PDL_ENTER();
    const int t = step[0];
    const float cf = clip_factor(sqnorm[0], max_grad);
    const float lr_t = adam_lr_t(t, lr);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
        // two independent float4 streams per thread: all eight loads are in flight before the first use
        const long long i1 = i0 + stride;
        bool on0 = (i0 & 63) != 63, on1 = i1 < n4 && (i1 & 63) != 63;
        if (flags) {
            if (on0 && flags[row0 + (i0 >> 6)] == t) on0 = false;
            if (on1 && flags[row0 + (i1 >> 6)] == t) on1 = false;
        }
        float4 p0, m0, v0, g0, p1, m1, v1, g1;
        if (on0) { p0 = item[i0]; m0 = m[i0]; v0 = v[i0]; g0 = g[i0]; }
        if (on1) { p1 = item[i1]; m1 = m[i1]; v1 = v[i1]; g1 = g[i1]; }
        if (on0) {
            adam_update(p0.x, m0.x, v0.x, g0.x * cf, lr_t);
            adam_update(p0.y, m0.y, v0.y, g0.y * cf, lr_t);
            adam_update(p0.z, m0.z, v0.z, g0.z * cf, lr_t);
            adam_update(p0.w, m0.w, v0.w, g0.w * cf, lr_t);
            item[i0] = p0; m[i0] = m0; v[i0] = v0;
            store_iext_items(iext, row0 + (i0 >> 6), (int)(i0 & 63) * 4, p0);
        }
        if (on1) {
            adam_update(p1.x, m1.x, v1.x, g1.x * cf, lr_t);
            adam_update(p1.y, m1.y, v1.y, g1.y * cf, lr_t);
            adam_update(p1.z, m1.z, v1.z, g1.z * cf, lr_t);
            adam_update(p1.w, m1.w, v1.w, g1.w * cf, lr_t);
            item[i1] = p1; m[i1] = m1; v[i1] = v1;
            store_iext_items(iext, row0 + (i1 >> 6), (int)(i1 & 63) * 4, p1);
        }
    }
}

// The same update for a short list of rows, ahead of the table-wide pass: one warp per entry (the clicked items of
// the NEXT batch, then its labels + 1).  A row listed several times is claimed once (atomicExch of the step number
// into its flag); adam_item_kernel skips every claimed row, so each row is updated exactly once per step by the same
// arithmetic whichever kernel does it.
__global__ void __launch_bounds__(256)
adam_item_rows_kernel(float4* __restrict__ item, float4* __restrict__ m, float4* __restrict__ v,
                      const float4* __restrict__ g, const float* __restrict__ sqnorm,
                      const int32_t* __restrict__ step, float lr, float max_grad, __nv_bfloat16* __restrict__ iext,
                      const int32_t* __restrict__ seq, int n_seq, const int32_t* __restrict__ label, int n_label,
                      int32_t* __restrict__ flags, int row_lo, int n_rows) {
    PDL_ENTER();
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n_seq + n_label) return;
    const int row = e < n_seq ? seq[e] : label[e - n_seq] + 1;
    // rows outside [row_lo, n_rows) are not this caller's to update (other catalog shard) and are not claimed
    if (row < row_lo || row >= n_rows) return;
    const int t = step[0];
    int old = 0;
    if (lane == 0) old = atomicExch(&flags[row], t);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == t) return;
    const float cf = clip_factor(sqnorm[0], max_grad);
    const float lr_t = adam_lr_t(t, lr);
    const size_t base = (size_t)row * (HP / 4);
    const bool two = lane < 31;                       // float4 slot 63 = pad columns
    const size_t i0 = base + lane, i1 = base + 32 + lane;
    float4 p0 = item[i0], m0 = m[i0], v0 = v[i0];
    const float4 g0 = g[i0];
    float4 p1 = p0, m1 = m0, v1 = v0, g1 = g0;
    if (two) { p1 = item[i1]; m1 = m[i1]; v1 = v[i1]; g1 = g[i1]; }
    adam_update(p0.x, m0.x, v0.x, g0.x * cf, lr_t);
    adam_update(p0.y, m0.y, v0.y, g0.y * cf, lr_t);
    adam_update(p0.z, m0.z, v0.z, g0.z * cf, lr_t);
    adam_update(p0.w, m0.w, v0.w, g0.w * cf, lr_t);
    item[i0] = p0; m[i0] = m0; v[i0] = v0;
    store_iext_items(iext, row, lane * 4, p0);
    if (two) {
        adam_update(p1.x, m1.x, v1.x, g1.x * cf, lr_t);
        adam_update(p1.y, m1.y, v1.y, g1.y * cf, lr_t);
        adam_update(p1.z, m1.z, v1.z, g1.z * cf, lr_t);
        adam_update(p1.w, m1.w, v1.w, g1.w * cf, lr_t);
        item[i1] = p1; m[i1] = m1; v[i1] = v1;
        store_iext_items(iext, row, (32 + lane) * 4, p1);
    }
}

}  // namespace tcar

using namespace tcar;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int tcar_build_iext(const float* item, const float* content, const int32_t* mwdhm, void* iext_bf16, int N,
                               int n_pad, void* stream) {
    if (N < 1 || n_pad < N || n_pad % 256) return TCAR_ERR_ARG;
    launch_pdl(build_iext_kernel, dim3((n_pad + 7) / 8), dim3(256), 0, STREAM, item, content, mwdhm,
                                                           static_cast<__nv_bfloat16*>(iext_bf16), N, n_pad);
    return (int)cudaGetLastError();
}

extern "C" int tcar_sqnorm_segments(const float* flat, const int32_t* seg_off, float* sqnorm, int nseg,
                                    void* stream) {
    if (nseg < 1) return TCAR_ERR_ARG;
    launch_pdl(sqnorm_segments_kernel, dim3(dim3(kSplit, nseg)), dim3(512), 0, STREAM, flat, seg_off, sqnorm);
    return (int)cudaGetLastError();
}

extern "C" int tcar_update_norms(const float* flat, const int32_t* seg_off, float* sqnorm_small, int nseg,
                                 const float* a, int na, const float* b, int nb, float* sqnorm_item,
                                 float* item_part, int32_t* ticket, int32_t* step, void* stream) {
    // sqnorm_item == NULL: small tensors only; nseg == 0: item norm only; step == NULL: no counter increment
    if (nseg < 0 || na < 0 || nb < 0 || (nb & 3) || (nseg == 0 && !sqnorm_item) || (sqnorm_item && (!item_part || !ticket)) ||
        (step && nseg == 0))
        return TCAR_ERR_ARG;
    launch_pdl(update_norms_kernel, dim3(kSplit, nseg + (sqnorm_item ? 1 : 0)), dim3(512), 0, STREAM, flat, seg_off, sqnorm_small, nseg, a, na,
               reinterpret_cast<const float4*>(b), nb / 4, sqnorm_item, item_part, ticket, step);
    return (int)cudaGetLastError();
}

extern "C" int tcar_sqnorm_big(const float* x, float* partial, float* sqnorm, long long n, void* stream) {
    if (n % 4) return TCAR_ERR_ARG;
    launch_pdl(sqnorm_big_partial_kernel, dim3(kNormBlocks), dim3(256), 0, STREAM, reinterpret_cast<const float4*>(x), partial, n / 4);
    int rc = (int)cudaGetLastError();
    if (rc) return rc;
    launch_pdl(sqnorm_big_final_kernel, dim3(1), dim3(1024), 0, STREAM, partial, sqnorm, kNormBlocks);
    return (int)cudaGetLastError();
}

extern "C" int tcar_refresh_iext_items(const float* item, void* iext_bf16, int N, void* stream) {
    if (N < 1) return TCAR_ERR_ARG;
    launch_pdl(refresh_iext_items_kernel, dim3(148 * 16), dim3(256), 0, STREAM, reinterpret_cast<const float4*>(item),
                                                            static_cast<__nv_bfloat16*>(iext_bf16),
                                                            (long long)(N + 1) * (HP / 4));
    return (int)cudaGetLastError();
}

extern "C" int tcar_sqnorm_combine(const float* a, int na, const float* b, int nb, float* out, void* stream) {
    if (na < 0 || nb < 0 || (nb & 3) || !out) return TCAR_ERR_ARG;
    launch_pdl(sqnorm_combine_kernel, dim3(1), dim3(1024), 0, STREAM, a, na, reinterpret_cast<const float4*>(b), nb / 4, out);
    return (int)cudaGetLastError();
}

extern "C" int tcar_adam_small(float* theta, float* m, float* v, const float* g, const int32_t* seg_off,
                               const float* sqnorm, int nseg, const int32_t* step, float lr, float max_grad,
                               void* stream) {
    if (nseg < 1) return TCAR_ERR_ARG;
    launch_pdl(adam_small_kernel, dim3(dim3(64, nseg)), dim3(256), 0, STREAM, theta, m, v, g, seg_off, sqnorm, step, lr, max_grad);
    return (int)cudaGetLastError();
}

extern "C" int tcar_adam_item(float* item, float* m, float* v, const float* g, const float* sqnorm,
                              const int32_t* step, float lr, float max_grad, void* iext_bf16, int row0, int nrows,
                              const int32_t* row_flags, int ctas_per_sm, void* stream) {
    if (row0 < 0 || nrows < 1 || ctas_per_sm < 0 || ctas_per_sm > 1024) return TCAR_ERR_ARG;
    const long long n4 = (long long)nrows * (HP / 4);
    const int grid = 148 * (ctas_per_sm ? ctas_per_sm : 64);
    launch_pdl(adam_item_kernel, dim3(grid), dim3(256), 0, STREAM, reinterpret_cast<float4*>(item), reinterpret_cast<float4*>(m),
                                               reinterpret_cast<float4*>(v), reinterpret_cast<const float4*>(g),
                                               sqnorm, step, lr, max_grad, static_cast<__nv_bfloat16*>(iext_bf16), n4,
                                               (long long)row0, row_flags);
    return (int)cudaGetLastError();
}

extern "C" int tcar_adam_item_rows(float* item, float* m, float* v, const float* g, const float* sqnorm,
                                   const int32_t* step, float lr, float max_grad, void* iext_bf16,
                                   const int32_t* seq, int n_seq, const int32_t* label, int n_label,
                                   int32_t* row_flags, int n_rows, void* stream) {
    if (n_seq < 0 || n_label < 0 || n_rows < 1 || !row_flags) return TCAR_ERR_ARG;
    const int entries = n_seq + n_label;
    if (entries == 0) return 0;
    launch_pdl(adam_item_rows_kernel, dim3((entries + 7) / 8), dim3(256), 0, STREAM, 
        reinterpret_cast<float4*>(item), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
        reinterpret_cast<const float4*>(g), sqnorm, step, lr, max_grad, static_cast<__nv_bfloat16*>(iext_bf16), seq,
        n_seq, label, n_label, row_flags, 0, n_rows);
    return (int)cudaGetLastError();
}

extern "C" int tcar_adam_item_rows_groups(float* item, float* m, float* v, const float* g, const float* sqnorm,
                                          const int32_t* step, float lr, float max_grad, void* iext_bf16,
                                          const int32_t* ids, long long ids_stride, const int* n_rows, int groups,
                                          int T, int Nn, int32_t* row_flags, int row_lo, int row_hi, void* stream) {
    if (!ids || !n_rows || groups < 1 || T < 1 || Nn < 0 || !row_flags || row_lo < 0 || row_hi < row_lo)
        return TCAR_ERR_ARG;
    for (int gr = 0; gr < groups; ++gr) {
        const int B = n_rows[gr];
        if (B <= 0) continue;
        // packed batch of rank gr: [7*B*T idx | 2*B ctx | B label | B*Nn neg]
        const int32_t* base = ids + gr * ids_stride;
        const size_t M = (size_t)B * T;
        const int32_t* lists[2] = {base + 7 * M + 2 * (size_t)B, base + 7 * M + 3 * (size_t)B};
        const int n_first[2] = {(int)M, 0}, n_second[2] = {B, B * Nn};
        for (int k = 0; k < 2; ++k) {
            const int entries = n_first[k] + n_second[k];
            if (entries == 0) continue;
            launch_pdl(adam_item_rows_kernel, dim3((entries + 7) / 8), dim3(256), 0, STREAM,
                reinterpret_cast<float4*>(item), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                reinterpret_cast<const float4*>(g), sqnorm, step, lr, max_grad,
                static_cast<__nv_bfloat16*>(iext_bf16), base, n_first[k], lists[k], n_second[k], row_flags, row_lo,
                row_hi);
            const int rc = (int)cudaGetLastError();
            if (rc) return rc;
        }
    }
    return 0;
}
