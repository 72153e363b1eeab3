// Full-catalog top-20 evaluation without materialising the [B,N] score matrix.
// Reference: model_combine.py:283-306 (scores -> argsort()[::-1][:20]) and util.py:8-18 (rank = #(S > S[label]) + 1).
//
// The scoring kernel (eval mode) leaves chunkmax[b, j] = max of the bf16-GEMM scores of items 8j..8j+7.  Every item
// of the true top-20 lives in one of the 20 chunks with the largest chunk maxima, so we take the 32 best chunks
// (12 chunks of slack for bf16 rounding), re-score their 256 items exactly in fp32 and sort those by
// (score desc, id asc).  Ties therefore resolve to the lower item id (north_star), which agrees with the
// reference on tie-free inputs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "tcar_b200.h"
#include "launch.cuh"

namespace tcar {

constexpr int H = TCAR_H, HP = TCAR_HP, XW = TCAR_XW, NB = TCAR_NBINS, TOPK = TCAR_TOPK;
constexpr int NCH = TCAR_NCAND_CHUNKS, CH = TCAR_CHUNK, NCAND = NCH * CH;  // 32 chunks x 8 = 256 candidates
__device__ __constant__ int kBinOffE[6] = {0, 13, 45, 53, 78, 139};

__device__ __forceinline__ float warp_sum_e(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// order-preserving map float -> uint32 (larger float <=> larger key); -inf is the smallest finite-comparable key
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// same arithmetic as exact_score() in session_kernels.cu (kept identical so label and candidates compare exactly)
__device__ __forceinline__ float exact_score_e(const float* s_aic, const float* s_tq, const float* __restrict__ item,
                                               const float* __restrict__ content,
                                               const int32_t* __restrict__ mwdhm, int n, int lane) {
    const float4* ir = reinterpret_cast<const float4*>(item + ((size_t)n + 1) * HP);
    const float4* cr = reinterpret_cast<const float4*>(content + ((size_t)n + 1) * HP);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = j * 128 + lane * 4;
        const float4 iv = __ldg(ir + j * 32 + lane), cv = __ldg(cr + j * 32 + lane);
        if (c + 0 < H) { acc = fmaf(iv.x, s_aic[c + 0], acc); acc = fmaf(cv.x, s_aic[H + c + 0], acc); }
        if (c + 1 < H) { acc = fmaf(iv.y, s_aic[c + 1], acc); acc = fmaf(cv.y, s_aic[H + c + 1], acc); }
        if (c + 2 < H) { acc = fmaf(iv.z, s_aic[c + 2], acc); acc = fmaf(cv.z, s_aic[H + c + 2], acc); }
        if (c + 3 < H) { acc = fmaf(iv.w, s_aic[c + 3], acc); acc = fmaf(cv.w, s_aic[H + c + 3], acc); }
    }
    acc = warp_sum_e(acc);
    float tsum = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) tsum += s_tq[kBinOffE[k] + mwdhm[(size_t)n * 5 + k]];
    return acc + tsum;
}

// (score desc, id asc) "a ranks before b"
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// Block-wide selection of the NCH largest of vals[0..n) (n > NCH) by (value desc, index asc): 4 x 8-bit radix passes
// on the order-preserving key find the NCH-th largest key, then strictly-greater entries are taken in any order and
// equal-key entries lowest index first.  256 threads; s_sel receives NCH indices.
__device__ __forceinline__ void select_top(const float* __restrict__ vals, int n, int* s_sel, int* s_hist, int* s_warp,
                                           int* s_misc /* [4]: cnt, need, digit, above */) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t prefix = 0, mask = 0;
    int kth = NCH;
    if (tid == 0) s_misc[0] = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 256) {
            const uint32_t k = fkey(vals[i]);
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int above = 0, d = 255;
            for (; d > 0; --d) {
                if (above + s_hist[d] >= kth) break;
                above += s_hist[d];
            }
            s_misc[2] = d;
            s_misc[3] = above;
        }
        __syncthreads();
        prefix |= (uint32_t)s_misc[2] << shift;
        mask |= 255u << shift;
        kth -= s_misc[3];
        __syncthreads();
    }
    // prefix = key of the NCH-th largest; kth = how many keys equal to it are still needed
    for (int i = tid; i < n; i += 256)
        if (fkey(vals[i]) > prefix) s_sel[atomicAdd(&s_misc[0], 1)] = i;
    __syncthreads();
    int base = s_misc[0];
    __syncthreads();
    for (int i0 = 0; i0 < n && base < NCH; i0 += 256) {
        const int i = i0 + tid;
        const bool eq = i < n && fkey(vals[i]) == prefix;
        const uint32_t bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int j = 0; j < w; ++j) off += s_warp[j];
        int tot = 0;
        for (int j = 0; j < 8; ++j) tot += s_warp[j];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (eq && pos < NCH) s_sel[pos] = i;
        base += tot;
        __syncthreads();
    }
    __syncthreads();
}

constexpr int TILE_CH = 128 / CH;          // 16 chunks per 128-item tile
constexpr int NCC = NCH * TILE_CH;         // 512 candidate chunks after the tile-level selection

// one CTA (256 threads) per query.  Two-level selection: the 32 best 128-item tiles by tile max (every one of the 32
// best chunks lives in one of them: the 32 largest tile maxima are 32 distinct chunk values, so the 32nd largest chunk
// is >= the 32nd largest tile max), then the 32 best of their 512 chunks, then the exact fp32 re-scoring.
__global__ void __launch_bounds__(256)
eval_topk_kernel(const float* __restrict__ chunkmax, const float* __restrict__ tilemax, const float* __restrict__ a_ic,
                 const float* __restrict__ Tq, const float* __restrict__ item, const float* __restrict__ content,
                 const int32_t* __restrict__ mwdhm, const int32_t* __restrict__ label, int32_t* __restrict__ top_ids,
                 float* __restrict__ top_scores, int32_t* __restrict__ n_greater, int N, int n_pad, int item_offset) {
    PDL_ENTER();
    __shared__ float s_aic[XW], s_tq[NB + 1];
    __shared__ int s_hist[256];
    __shared__ int s_sel[NCH];
    __shared__ int s_misc[4];
    __shared__ float s_cv[NCC];
    __shared__ int s_ci[NCC];
    __shared__ float s_sc[NCAND];
    __shared__ int s_id[NCAND];
    __shared__ int s_warp[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nchunks = (N + CH - 1) / CH;
    const int ntiles = (N + 127) / 128;
    const float* cm = chunkmax + (size_t)b * (n_pad / CH);
    const float* tm = tilemax + (size_t)b * (n_pad / 128);
    for (int c = tid; c < XW; c += 256) s_aic[c] = a_ic[(size_t)b * XW + c];
    for (int c = tid; c < NB; c += 256) s_tq[c] = Tq[(size_t)b * NB + c];
    for (int i = tid; i < NCH; i += 256) s_sel[i] = -1;
    __syncthreads();

    if (nchunks > NCH) {
        // ---- level 1: tiles
        if (ntiles <= NCH) {
            for (int i = tid; i < ntiles; i += 256) s_sel[i] = i;
            __syncthreads();
        } else {
            select_top(tm, ntiles, s_sel, s_hist, s_warp, s_misc);
        }
        // ---- level 2: the 16 chunks of every selected tile, sorted by (max desc, chunk index asc)
        for (int i = tid; i < NCC; i += 256) {
            const int tile = s_sel[i / TILE_CH];
            const int chunk = tile * TILE_CH + (i % TILE_CH);
            const bool ok = tile >= 0 && chunk < nchunks;
            s_cv[i] = ok ? cm[chunk] : -INFINITY;
            s_ci[i] = ok ? chunk : 0x7fffffff;
        }
        for (int k = 2; k <= NCC; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int i = tid; i < NCC; i += 256) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const float sa = s_cv[i], sb = s_cv[ixj];
                        const int ia = s_ci[i], ib = s_ci[ixj];
                        const bool up = (i & k) == 0;
                        const bool swap = up ? before(sb, ib, sa, ia) : before(sa, ia, sb, ib);
                        if (swap) { s_cv[i] = sb; s_cv[ixj] = sa; s_ci[i] = ib; s_ci[ixj] = ia; }
                    }
                }
            }
        }
        __syncthreads();
        if (tid < NCH) s_sel[tid] = s_ci[tid] == 0x7fffffff ? -1 : s_ci[tid];
    } else {
        for (int i = tid; i < nchunks; i += 256) s_sel[i] = i;
    }
    __syncthreads();

    // exact re-scoring: warp w handles candidates w*32 .. w*32+31
    const int lab = label[b];
    float lab_score = 0.f;
    {
        const float ls = exact_score_e(s_aic, s_tq, item, content, mwdhm, lab, lane);
        lab_score = ls;
    }
    for (int j = 0; j < 32; ++j) {
        const int ci = w * 32 + j;
        const int chunk = s_sel[ci / CH];
        const int n = chunk < 0 ? -1 : chunk * CH + (ci % CH);          // local item id
        const bool ok = n >= 0 && n < N;
        float sc = -INFINITY;
        if (ok) sc = exact_score_e(s_aic, s_tq, item, content, mwdhm, n + item_offset, lane);
        if (lane == 0) {
            s_sc[ci] = sc;
            s_id[ci] = ok ? n + item_offset : 0x7fffffff;
        }
    }
    __syncthreads();
    // n_greater: strict, never counts the label itself                       (util.py:14)
    {
        const bool gt = s_id[tid] != 0x7fffffff && s_id[tid] != lab && s_sc[tid] > lab_score;
        const uint32_t bal = __ballot_sync(0xffffffffu, gt);
        if (lane == 0) s_warp[w] = __popc(bal);
    }
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int j = 0; j < 8; ++j) t += s_warp[j];
        n_greater[b] = t;
    }
    // bitonic sort of 256 (score, id) pairs, ascending in "before" order
    for (int k = 2; k <= NCAND; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            const int ixj = tid ^ j;
            if (ixj > tid) {
                const float sa = s_sc[tid], sb = s_sc[ixj];
                const int ia = s_id[tid], ib = s_id[ixj];
                const bool up = (tid & k) == 0;
                const bool swap = up ? before(sb, ib, sa, ia) : before(sa, ia, sb, ib);
                if (swap) { s_sc[tid] = sb; s_sc[ixj] = sa; s_id[tid] = ib; s_id[ixj] = ia; }
            }
        }
    }
    __syncthreads();
    if (tid < TOPK) {
        top_ids[(size_t)b * TOPK + tid] = s_id[tid] == 0x7fffffff ? -1 : s_id[tid];
        top_scores[(size_t)b * TOPK + tid] = s_sc[tid];
    }
}

// merge G shard lists: one warp per query, serial selection (G*20 <= 160 entries)
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int32_t* __restrict__ ids, const float* __restrict__ scores, int32_t* __restrict__ out_ids,
                  float* __restrict__ out_scores, int G, int B) {
    PDL_ENTER();
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n = G * TOPK;
    // each lane owns entries lane, lane+32, ...; `taken` bit per owned entry
    uint32_t taken = 0;
    for (int r = 0; r < TOPK; ++r) {
        float bs = -INFINITY; int bi = 0x7fffffff, bslot = -1;
        for (int e = lane, s = 0; e < n; e += 32, ++s) {
            if (taken & (1u << s)) continue;
            const int g = e / TOPK, j = e % TOPK;
            const float sc = scores[((size_t)g * B + b) * TOPK + j];
            int id = ids[((size_t)g * B + b) * TOPK + j];
            if (id < 0) id = 0x7fffffff;
            if (before(sc, id, bs, bi)) { bs = sc; bi = id; bslot = s; }
        }
        // warp arg-best
        float ws = bs; int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ws, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (before(os, oi, ws, wi)) { ws = os; wi = oi; }
        }
        if (bslot >= 0 && bs == ws && bi == wi && wi != 0x7fffffff) taken |= 1u << bslot;
        if (lane == 0) {
            out_ids[(size_t)b * TOPK + r] = wi == 0x7fffffff ? -1 : wi;
            out_scores[(size_t)b * TOPK + r] = ws;
        }
    }
}

}  // namespace tcar

using namespace tcar;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int tcar_eval_topk(const float* chunkmax, const float* tilemax, const float* a_ic, const float* Tq,
                              const float* item,
                              const float* content, const int32_t* mwdhm, const int32_t* label, int32_t* top_ids,
                              float* top_scores, int32_t* n_greater, int B, int N, int n_pad, int item_offset,
                              void* stream) {
    if (B < 1 || B > TCAR_QROWS || N < 1 || n_pad < N || !tilemax) return TCAR_ERR_ARG;
    launch_pdl(eval_topk_kernel, dim3(B), dim3(256), 0, STREAM, chunkmax, tilemax, a_ic, Tq, item, content, mwdhm, label, top_ids, top_scores,
                                            n_greater, N, n_pad, item_offset);
    return (int)cudaGetLastError();
}

extern "C" int tcar_topk_merge(const int32_t* ids, const float* scores, int32_t* out_ids, float* out_scores, int G,
                               int B, void* stream) {
    if (G < 1 || G > 8 || B < 1) return TCAR_ERR_ARG;
    launch_pdl(topk_merge_kernel, dim3((B + 7) / 8), dim3(256), 0, STREAM, ids, scores, out_ids, out_scores, G, B);
    return (int)cudaGetLastError();
}
