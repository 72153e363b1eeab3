// Session-side kernels of the TCAR hot path (HBM / latency bound, fp32): fused gather with max_norm clip,
// attention pooling forward / backward, query-operand build, losses, and the gradient plumbing around the
// scoring GEMMs.  Reference call sites are cited per kernel; formulas follow SURVEY.md Appendix A.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include "tcar_b200.h"
#include "launch.cuh"

namespace tcar {

constexpr int H = TCAR_H, HP = TCAR_HP, TH = TCAR_TH, XW = TCAR_XW, PW = TCAR_PW, NB = TCAR_NBINS;
__device__ __constant__ int kBinOff[6] = {0, 13, 45, 53, 78, 139};  // month, day, week, hour, minute row offsets

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// fixed-order block reduction (result valid in all threads); `red` >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}
// tf.nn.embedding_lookup(max_norm=1) scale: 1 / max(||x||, 1)            (modules.py:36)
__device__ __forceinline__ float clip_scale(float sq) {
    const float n = sqrtf(sq);
    return n > 1.f ? 1.f / n : 1.f;
}

// ------------------------------------------------------------------------------------------------ (1) gather
// one warp per click (rows of X/P/D) and one warp per session for the click-time context CT.
// A row of X is 125 aligned float4 slots: slots 0..61 hold item+pos columns 0..247, slot 62 holds item+pos columns
// 248,249 and content columns 0,1, slots 63..124 hold content columns 2..249.  Lane l owns slots l, l+32, l+64, l+96,
// fetches exactly the source words that land in them (float4 / float2 loads, all issued before the first use) and
// writes each slot with one 128-bit store, so a warp store instruction covers 512 contiguous bytes.
__device__ __forceinline__ float sq4(const float4 v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
__device__ __forceinline__ float4 join2(const float2 a, const float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

__global__ void __launch_bounds__(128)
gather_fwd_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ ctx, const float* __restrict__ item,
                  const float* __restrict__ content, const float* __restrict__ pos,
                  const float* __restrict__ month, const float* __restrict__ day, const float* __restrict__ week,
                  const float* __restrict__ hour, const float* __restrict__ minute, const float* __restrict__ dur,
                  float* __restrict__ X, float* __restrict__ P, float* __restrict__ D, float* __restrict__ CT, int B,
                  int T) {
    PDL_ENTER();
    const int M = B * T;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w < M) {
        const int m = w, t = m % T;
        // the seven index words of the click: lanes 0..6 fetch one each
        const int my = lane < 7 ? __ldg(idx + (size_t)lane * M + m) : 0;
        const int id = __shfl_sync(0xffffffffu, my, 0);
        const float4* ir4 = reinterpret_cast<const float4*>(item + (size_t)id * HP);
        const float2* ir2 = reinterpret_cast<const float2*>(item + (size_t)id * HP);
        const float2* cr2 = reinterpret_cast<const float2*>(content + (size_t)id * HP);
        const float2* pr2 = reinterpret_cast<const float2*>(pos + (size_t)t * H);     // rows of 1000 B: 8-byte aligned
        const float* tabs[6] = {month, day, week, hour, minute, dur};
        float2 tv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int r = __shfl_sync(0xffffffffu, my, k + 1);
            tv[k] = __ldg(reinterpret_cast<const float2*>(tabs[k] + (size_t)r * TH) + lane);
        }
        // slot 0 (item + pos), slot 1 (item + pos | mixed | content), slots 2, 3 (content)
        const float4 a0 = __ldg(ir4 + lane);
        const float4 p0 = join2(__ldg(pr2 + 2 * lane), __ldg(pr2 + 2 * lane + 1));
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = a1;
        float sqi = sq4(a0), sqp = sq4(p0), sqc = 0.f;
        if (lane < 30) {
            a1 = __ldg(ir4 + 32 + lane);
            p1 = join2(__ldg(pr2 + 64 + 2 * lane), __ldg(pr2 + 65 + 2 * lane));
            sqi += sq4(a1);
            sqp += sq4(p1);
        } else if (lane == 30) {
            const float2 i2 = __ldg(ir2 + 124), q2 = __ldg(pr2 + 124), c2 = __ldg(cr2);
            a1 = join2(i2, c2);
            p1 = make_float4(q2.x, q2.y, 0.f, 0.f);
            sqi += i2.x * i2.x + i2.y * i2.y;
            sqp += q2.x * q2.x + q2.y * q2.y;
            sqc += c2.x * c2.x + c2.y * c2.y;
        } else {
            a1 = join2(__ldg(cr2 + 1), __ldg(cr2 + 2));                                  // slot 63: content 2..5
            sqc += sq4(a1);
        }
        // slot f >= 63 holds content columns 4 (f - 63) + 2 .. + 5  =  float2 words 2 (f - 63) + 1, + 2
        const float4 a2 = join2(__ldg(cr2 + 2 * lane + 3), __ldg(cr2 + 2 * lane + 4));   // f = 64 + lane
        float4 a3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 29) a3 = join2(__ldg(cr2 + 2 * lane + 67), __ldg(cr2 + 2 * lane + 68));   // f = 96 + lane
        sqc += sq4(a2) + sq4(a3);
        float tsq[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) tsq[k] = tv[k].x * tv[k].x + tv[k].y * tv[k].y;
        // nine independent butterfly reductions, interleaved
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sqi += __shfl_xor_sync(0xffffffffu, sqi, o);
            sqp += __shfl_xor_sync(0xffffffffu, sqp, o);
            sqc += __shfl_xor_sync(0xffffffffu, sqc, o);
#pragma unroll
            for (int k = 0; k < 6; ++k) tsq[k] += __shfl_xor_sync(0xffffffffu, tsq[k], o);
        }
        const float si = clip_scale(sqi), sp = clip_scale(sqp), sc = clip_scale(sqc);
        float4* x4 = reinterpret_cast<float4*>(X + (size_t)m * XW);
        x4[lane] = make_float4(fmaf(p0.x, sp, a0.x * si), fmaf(p0.y, sp, a0.y * si), fmaf(p0.z, sp, a0.z * si),
                               fmaf(p0.w, sp, a0.w * si));
        float4 o1;
        if (lane < 30)
            o1 = make_float4(fmaf(p1.x, sp, a1.x * si), fmaf(p1.y, sp, a1.y * si), fmaf(p1.z, sp, a1.z * si),
                             fmaf(p1.w, sp, a1.w * si));
        else if (lane == 30)
            o1 = make_float4(fmaf(p1.x, sp, a1.x * si), fmaf(p1.y, sp, a1.y * si), a1.z * sc, a1.w * sc);
        else
            o1 = make_float4(a1.x * sc, a1.y * sc, a1.z * sc, a1.w * sc);
        x4[32 + lane] = o1;
        x4[64 + lane] = make_float4(a2.x * sc, a2.y * sc, a2.z * sc, a2.w * sc);
        if (lane < 29) x4[96 + lane] = make_float4(a3.x * sc, a3.y * sc, a3.z * sc, a3.w * sc);
        // five publish-time rows and the dwell-time row, 64 floats each: one float2 per lane
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float s = clip_scale(tsq[k]);
            const float2 o = make_float2(tv[k].x * s, tv[k].y * s);
            if (k < 5) reinterpret_cast<float2*>(P + (size_t)m * PW + k * TH)[lane] = o;
            else reinterpret_cast<float2*>(D + (size_t)m * TH)[lane] = o;
        }
    } else if (w < M + B) {
        const int b = w - M;
        const int cw = ctx[b], ch = ctx[B + b];
        const float2 a = __ldg(reinterpret_cast<const float2*>(week + (size_t)cw * TH) + lane);
        const float2 h = __ldg(reinterpret_cast<const float2*>(hour + (size_t)ch * TH) + lane);
        const float sa = clip_scale(warp_sum(a.x * a.x + a.y * a.y));
        const float sh = clip_scale(warp_sum(h.x * h.x + h.y * h.y));
        reinterpret_cast<float2*>(CT + (size_t)b * 2 * TH)[lane] = make_float2(a.x * sa, a.y * sa);
        reinterpret_cast<float2*>(CT + (size_t)b * 2 * TH + TH)[lane] = make_float2(h.x * sh, h.y * sh);
    }
}

// ------------------------------------------------------------------------------------------------ (0) batch assembly
// GPU-resident sampler: the packed batch [7*B*T idx | 2*B ctx | B label | B*Nn neg] that
// Sampler.next_packed() builds on the host (sampler.py:52-113) is gathered on the device from the columnar cache of
// one session-length bucket: seq [n, T+1], feats [6, n, T], ctx [2, n]; `rows` [B] = bucket rows of the batch.
// Negatives are either copied from neg_in (host-drawn: the reference's NumPy stream, bit-exact) or drawn on the device:
// Philox4x32-10 with key = seed, counter = (offset + element / 4), word element % 4, mapped to [0, item_num) by
// (x * item_num) >> 32  (restated in oracle/philox_oracle.py).
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ uint32_t philox_word(unsigned long long seed, unsigned long long ctr, int word) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c[word];
}

__global__ void __launch_bounds__(256)
assemble_batch_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ seq,
                      const int32_t* __restrict__ feats, const int32_t* __restrict__ ctx, int n, int B, int T, int Nn,
                      const int32_t* __restrict__ neg_in, int item_num, unsigned long long seed,
                      unsigned long long offset, int32_t* __restrict__ out) {
    PDL_ENTER();
    const int M = B * T;
    const int total = 7 * M + 3 * B + B * Nn;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int v;
        if (i < 7 * M) {
            const int k = i / M, m = i - k * M, b = m / T, t = m - b * T;
            const int r = rows[b];
            v = k == 0 ? seq[(size_t)r * (T + 1) + t] : feats[((size_t)(k - 1) * n + r) * T + t];
        } else if (i < 7 * M + 2 * B) {
            const int j = i - 7 * M, k = j / B, b = j - k * B;
            v = ctx[(size_t)k * n + rows[b]];
        } else if (i < 7 * M + 3 * B) {
            const int b = i - 7 * M - 2 * B;
            v = seq[(size_t)rows[b] * (T + 1) + T] - 1;                     // target, 0-based (sampler.py:72)
        } else {
            const int e = i - 7 * M - 3 * B;
            if (neg_in) v = neg_in[e];
            else v = (int)(((unsigned long long)philox_word(seed, offset + (unsigned long long)(e >> 2), e & 3) *
                            (unsigned long long)item_num) >> 32);
        }
        out[i] = v;
    }
}

// Impression-list negatives (sampler.py:118-131, neg_neighbor_from_impre) on the device, one thread per session:
// up to 21 uniform draws from the session's impression list (`random.choice`), a draw counts when the article is in
// item_dict (impr_ids >= 0, already mapped to 0-based item ids) until Nn are found; the rest is filled with uniform
// draws from [0, item_num) (np.random.randint).  Same algorithm, counter-based stream instead of Python's Mersenne
// twister: session b of the batch owns the TCAR_IMPR_BLOCKS(Nn) Philox counters from offset + b * TCAR_IMPR_BLOCKS(Nn);
// word j < 21 drives try j, word 21 + k fills slot k (restated in oracle/philox_oracle.py).
__global__ void __launch_bounds__(128)
impression_negatives_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ impr_off,
                            const int32_t* __restrict__ impr_ids, int B, int Nn, int item_num,
                            unsigned long long seed, unsigned long long offset, int32_t* __restrict__ out) {
    PDL_ENTER();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int r = rows[b];
    const int lo = impr_off[r], len = impr_off[r + 1] - lo;
    const unsigned long long base = offset + (unsigned long long)b * TCAR_IMPR_BLOCKS(Nn);
    int32_t* neg = out + (size_t)b * Nn;
    int found = 0;
    if (len > 0) {
        for (int j = 0; j < 21 && found < Nn; ++j) {
            const uint32_t wv = philox_word(seed, base + (unsigned long long)(j >> 2), j & 3);
            const int id = impr_ids[lo + (int)(((unsigned long long)wv * (unsigned long long)len) >> 32)];
            if (id >= 0) neg[found++] = id;
        }
    }
    for (int k = found; k < Nn; ++k) {
        const int j = 21 + k;
        const uint32_t wv = philox_word(seed, base + (unsigned long long)(j >> 2), j & 3);
        neg[k] = (int)(((unsigned long long)wv * (unsigned long long)item_num) >> 32);
    }
}

// ------------------------------------------------------------------------------------------------ (2) pooling fwd
// one CTA per session.  modules.py:126-142 (count_alpha_m), :94-100 (count_alpha_s), :116-117 / :82-83 (pool).
// Every row is fetched with 128-bit loads that are all issued before the first use (8 independent loads per lane and
// click), so a click costs one memory round trip instead of ~30 dependent ones.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

__global__ void __launch_bounds__(256, 4)
pool_fwd_kernel(const float* __restrict__ X, const float* __restrict__ P, float* __restrict__ U1,
                float* __restrict__ U2, const float* __restrict__ q, const float* __restrict__ w_r,
                const float* __restrict__ w_t, float* __restrict__ alpha, float* __restrict__ pooled,
                float* __restrict__ pooled_t, int B, int T) {
    PDL_ENTER();
    __shared__ float s_e[3][TCAR_MAXT];
    __shared__ float s_a[3][TCAR_MAXT];
    __shared__ __align__(16) float s_q[XW + 12];
    __shared__ __align__(16) float s_wr[HP], s_wt[HP];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int M = B * T;
    for (int c = threadIdx.x; c < XW + 12; c += blockDim.x) s_q[c] = c < XW ? q[(size_t)b * XW + c] : 0.f;
    for (int c = threadIdx.x; c < HP; c += blockDim.x) {
        s_wr[c] = c < H ? w_r[c] : 0.f;          // zero weights on the 6 pad columns of the 256-float pitch
        s_wt[c] = c < H ? w_t[c] : 0.f;
    }
    __syncthreads();
    const float4* q4 = reinterpret_cast<const float4*>(s_q);
    const float4* wr4 = reinterpret_cast<const float4*>(s_wr);
    const float4* wt4 = reinterpret_cast<const float4*>(s_wt);
    for (int t = w; t < T; t += 8) {
        const size_t m = (size_t)b * T + t;
        float4* u1 = reinterpret_cast<float4*>(U1 + m * HP);
        float4* u2 = reinterpret_cast<float4*>(U2 + m * HP);
        const float4* x4 = reinterpret_cast<const float4*>(X + m * XW);       // 125 float4 per row
        const float4 a0 = u1[lane], a1 = u1[lane + 32], c0 = u2[lane], c1 = u2[lane + 32];
        const float4 x0 = x4[lane], x1 = x4[lane + 32], x2 = x4[lane + 64];
        const float4 x3 = lane + 96 < XW / 4 ? x4[lane + 96] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 s0 = make_float4(sigmoidf_(a0.x), sigmoidf_(a0.y), sigmoidf_(a0.z), sigmoidf_(a0.w));
        float4 s1 = make_float4(sigmoidf_(a1.x), sigmoidf_(a1.y), sigmoidf_(a1.z), sigmoidf_(a1.w));
        float4 r0 = make_float4(sigmoidf_(c0.x), sigmoidf_(c0.y), sigmoidf_(c0.z), sigmoidf_(c0.w));
        float4 r1 = make_float4(sigmoidf_(c1.x), sigmoidf_(c1.y), sigmoidf_(c1.z), sigmoidf_(c1.w));
        if (lane == 30) { s1.z = s1.w = 0.f; r1.z = r1.w = 0.f; }               // columns 250, 251
        if (lane == 31) { s1 = make_float4(0.f, 0.f, 0.f, 0.f); r1 = s1; }      // columns 252..255
        u1[lane] = s0; u1[lane + 32] = s1;
        u2[lane] = r0; u2[lane + 32] = r1;
        float e1 = dot4(s0, wr4[lane]) + dot4(s1, wr4[lane + 32]);
        float et = dot4(r0, wt4[lane]) + dot4(r1, wt4[lane + 32]);
        float e2 = (dot4(x0, q4[lane]) + dot4(x1, q4[lane + 32])) + (dot4(x2, q4[lane + 64]) + dot4(x3, q4[lane + 96]));
        e1 = warp_sum(e1); e2 = warp_sum(e2); et = warp_sum(et);
        if (lane == 0) { s_e[0][t] = e1; s_e[1][t] = e2; s_e[2][t] = et; }
    }
    __syncthreads();
    if (w < 3) {
        // nrm(x) = exp(x) / (sum exp(x) + 1e-9), no max subtraction            (util.py:92-100)
        const float x0 = lane < T ? expf(s_e[w][lane]) : 0.f;
        const float x1 = lane + 32 < T ? expf(s_e[w][lane + 32]) : 0.f;
        const float sum = warp_sum(x0 + x1) + 1e-9f;
        if (lane < T) { s_a[w][lane] = x0 / sum; alpha[(size_t)w * M + (size_t)b * T + lane] = x0 / sum; }
        if (lane + 32 < T) { s_a[w][lane + 32] = x1 / sum; alpha[(size_t)w * M + (size_t)b * T + lane + 32] = x1 / sum; }
    }
    __syncthreads();
    // pooled = sum_t (alpha1 + alpha2)[t] X[t, :] (125 float4 columns), pooled_t = sum_t alpha_t[t] P[t, :] (80)
    for (int c = threadIdx.x; c < XW / 4 + PW / 4; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool isx = c < XW / 4;
        const float4* src = isx ? reinterpret_cast<const float4*>(X + (size_t)b * T * XW) + c
                                : reinterpret_cast<const float4*>(P + (size_t)b * T * PW) + (c - XW / 4);
        const int stride = isx ? XW / 4 : PW / 4;
        int t = 0;
        for (; t + 4 <= T; t += 4) {
            const float4 v0 = src[(size_t)t * stride], v1 = src[(size_t)(t + 1) * stride];
            const float4 v2 = src[(size_t)(t + 2) * stride], v3 = src[(size_t)(t + 3) * stride];
            const float g0 = isx ? s_a[0][t] + s_a[1][t] : s_a[2][t];
            const float g1 = isx ? s_a[0][t + 1] + s_a[1][t + 1] : s_a[2][t + 1];
            const float g2 = isx ? s_a[0][t + 2] + s_a[1][t + 2] : s_a[2][t + 2];
            const float g3 = isx ? s_a[0][t + 3] + s_a[1][t + 3] : s_a[2][t + 3];
            acc.x = fmaf(g0, v0.x, acc.x); acc.y = fmaf(g0, v0.y, acc.y); acc.z = fmaf(g0, v0.z, acc.z); acc.w = fmaf(g0, v0.w, acc.w);
            acc.x = fmaf(g1, v1.x, acc.x); acc.y = fmaf(g1, v1.y, acc.y); acc.z = fmaf(g1, v1.z, acc.z); acc.w = fmaf(g1, v1.w, acc.w);
            acc.x = fmaf(g2, v2.x, acc.x); acc.y = fmaf(g2, v2.y, acc.y); acc.z = fmaf(g2, v2.z, acc.z); acc.w = fmaf(g2, v2.w, acc.w);
            acc.x = fmaf(g3, v3.x, acc.x); acc.y = fmaf(g3, v3.y, acc.y); acc.z = fmaf(g3, v3.z, acc.z); acc.w = fmaf(g3, v3.w, acc.w);
        }
        for (; t < T; ++t) {
            const float4 v0 = src[(size_t)t * stride];
            const float g0 = isx ? s_a[0][t] + s_a[1][t] : s_a[2][t];
            acc.x = fmaf(g0, v0.x, acc.x); acc.y = fmaf(g0, v0.y, acc.y); acc.z = fmaf(g0, v0.z, acc.z); acc.w = fmaf(g0, v0.w, acc.w);
        }
        if (isx) reinterpret_cast<float4*>(pooled + (size_t)b * XW)[c] = acc;
        else reinterpret_cast<float4*>(pooled_t + (size_t)b * PW)[c - XW / 4] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ (2b) pooling bwd
__global__ void __launch_bounds__(256, 4)
pool_bwd_kernel(const float* __restrict__ X, const float* __restrict__ P, const float* __restrict__ S1,
                const float* __restrict__ S2, const float* __restrict__ q, const float* __restrict__ w_r,
                const float* __restrict__ w_t, const float* __restrict__ alpha, const float* __restrict__ dpooled,
                const float* __restrict__ dpooled_t, float* __restrict__ dU1, float* __restrict__ dU2,
                float* __restrict__ dXi, float* __restrict__ dP, float* __restrict__ dq, float* __restrict__ de,
                int B, int T) {
    PDL_ENTER();
    __shared__ __align__(16) float s_dp[XW + 12], s_dpt[PW], s_q[XW + 12];
    __shared__ __align__(16) float s_wr[HP], s_wt[HP];
    __shared__ float s_a[3][TCAR_MAXT], s_da[2][TCAR_MAXT], s_de[3][TCAR_MAXT];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int M = B * T;
    for (int c = threadIdx.x; c < XW + 12; c += blockDim.x) {
        s_dp[c] = c < XW ? dpooled[(size_t)b * XW + c] : 0.f;
        s_q[c] = c < XW ? q[(size_t)b * XW + c] : 0.f;
    }
    for (int c = threadIdx.x; c < PW; c += blockDim.x) s_dpt[c] = dpooled_t[(size_t)b * PW + c];
    for (int c = threadIdx.x; c < HP; c += blockDim.x) {
        s_wr[c] = c < H ? w_r[c] : 0.f;
        s_wt[c] = c < H ? w_t[c] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) s_a[i / T][i % T] = alpha[(size_t)(i / T) * M + (size_t)b * T + i % T];
    __syncthreads();
    const float4* dp4 = reinterpret_cast<const float4*>(s_dp);
    const float4* dpt4 = reinterpret_cast<const float4*>(s_dpt);
    const float4* q4 = reinterpret_cast<const float4*>(s_q);
    const float4* wr4 = reinterpret_cast<const float4*>(s_wr);
    const float4* wt4 = reinterpret_cast<const float4*>(s_wt);
    // d alpha[t] = X[t,:] . dpooled ; d alpha_t[t] = P[t,:] . dpooled_t
    for (int t = w; t < T; t += 8) {
        const size_t m = (size_t)b * T + t;
        const float4* x4 = reinterpret_cast<const float4*>(X + m * XW);
        const float4* p4 = reinterpret_cast<const float4*>(P + m * PW);       // 80 float4 per row
        const float4 x0 = x4[lane], x1 = x4[lane + 32], x2 = x4[lane + 64];
        const float4 x3 = lane + 96 < XW / 4 ? x4[lane + 96] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 p0 = p4[lane], p1 = p4[lane + 32];
        const float4 p2 = lane + 64 < PW / 4 ? p4[lane + 64] : make_float4(0.f, 0.f, 0.f, 0.f);
        float a = (dot4(x0, dp4[lane]) + dot4(x1, dp4[lane + 32])) + (dot4(x2, dp4[lane + 64]) + dot4(x3, dp4[lane + 96]));
        float at = dot4(p0, dpt4[lane]) + dot4(p1, dpt4[lane + 32]) + (lane + 64 < PW / 4 ? dot4(p2, dpt4[lane + 64]) : 0.f);
        a = warp_sum(a); at = warp_sum(at);
        if (lane == 0) { s_da[0][t] = a; s_da[1][t] = at; }
    }
    __syncthreads();
    // softmax' backward: de[t] = alpha[t] (d alpha[t] - sum_s alpha[s] d alpha[s])
    if (w < 3) {
        const int src = (w == 2) ? 1 : 0;
        const float p0 = lane < T ? s_a[w][lane] * s_da[src][lane] : 0.f;
        const float p1 = lane + 32 < T ? s_a[w][lane + 32] * s_da[src][lane + 32] : 0.f;
        const float dot = warp_sum(p0 + p1);
        if (lane < T) {
            const float v = s_a[w][lane] * (s_da[src][lane] - dot);
            s_de[w][lane] = v; de[(size_t)w * M + (size_t)b * T + lane] = v;
        }
        if (lane + 32 < T) {
            const float v = s_a[w][lane + 32] * (s_da[src][lane + 32] - dot);
            s_de[w][lane + 32] = v; de[(size_t)w * M + (size_t)b * T + lane + 32] = v;
        }
    }
    __syncthreads();
    for (int t = w; t < T; t += 8) {
        const size_t m = (size_t)b * T + t;
        const float de1 = s_de[0][t], de2 = s_de[1][t], det = s_de[2][t];
        const float a12 = s_a[0][t] + s_a[1][t], at = s_a[2][t];
        const float4* s1r = reinterpret_cast<const float4*>(S1 + m * HP);
        const float4* s2r = reinterpret_cast<const float4*>(S2 + m * HP);
        float4* du1 = reinterpret_cast<float4*>(dU1 + m * HP);
        float4* du2 = reinterpret_cast<float4*>(dU2 + m * HP);
        float4* dxi = reinterpret_cast<float4*>(dXi + m * HP);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;                 // float4 index within the 256-float pitch (pads give zeros)
            const float4 s1 = s1r[i], s2 = s2r[i], wr = wr4[i], wt = wt4[i];
            // dXi uses columns 0..249 of dpooled / q (the item half of X)
            const float4 dpv = i < 63 ? dp4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 qv = i < 63 ? q4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o1, o2, ox;
            o1.x = de1 * wr.x * s1.x * (1.f - s1.x); o1.y = de1 * wr.y * s1.y * (1.f - s1.y);
            o1.z = de1 * wr.z * s1.z * (1.f - s1.z); o1.w = de1 * wr.w * s1.w * (1.f - s1.w);
            o2.x = det * wt.x * s2.x * (1.f - s2.x); o2.y = det * wt.y * s2.y * (1.f - s2.y);
            o2.z = det * wt.z * s2.z * (1.f - s2.z); o2.w = det * wt.w * s2.w * (1.f - s2.w);
            ox.x = fmaf(a12, dpv.x, de2 * qv.x); ox.y = fmaf(a12, dpv.y, de2 * qv.y);
            ox.z = fmaf(a12, dpv.z, de2 * qv.z); ox.w = fmaf(a12, dpv.w, de2 * qv.w);
            if (i == 62) { ox.z = ox.w = 0.f; }            // columns 250, 251 belong to the content half
            du1[i] = o1; du2[i] = o2; dxi[i] = ox;
        }
        float4* dp_out = reinterpret_cast<float4*>(dP + m * PW);
        for (int i = lane; i < PW / 4; i += 32) {
            const float4 v = dpt4[i];
            dp_out[i] = make_float4(at * v.x, at * v.y, at * v.z, at * v.w);
        }
    }
    // dq[c] = sum_t de2[t] X[t, c]
    for (int c = threadIdx.x; c < XW / 4; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* src = reinterpret_cast<const float4*>(X + (size_t)b * T * XW) + c;
        int t = 0;
        for (; t + 4 <= T; t += 4) {
            const float4 v0 = src[(size_t)t * (XW / 4)], v1 = src[(size_t)(t + 1) * (XW / 4)];
            const float4 v2 = src[(size_t)(t + 2) * (XW / 4)], v3 = src[(size_t)(t + 3) * (XW / 4)];
            const float g0 = s_de[1][t], g1 = s_de[1][t + 1], g2 = s_de[1][t + 2], g3 = s_de[1][t + 3];
            acc.x = fmaf(g0, v0.x, acc.x); acc.y = fmaf(g0, v0.y, acc.y); acc.z = fmaf(g0, v0.z, acc.z); acc.w = fmaf(g0, v0.w, acc.w);
            acc.x = fmaf(g1, v1.x, acc.x); acc.y = fmaf(g1, v1.y, acc.y); acc.z = fmaf(g1, v1.z, acc.z); acc.w = fmaf(g1, v1.w, acc.w);
            acc.x = fmaf(g2, v2.x, acc.x); acc.y = fmaf(g2, v2.y, acc.y); acc.z = fmaf(g2, v2.z, acc.z); acc.w = fmaf(g2, v2.w, acc.w);
            acc.x = fmaf(g3, v3.x, acc.x); acc.y = fmaf(g3, v3.y, acc.y); acc.z = fmaf(g3, v3.z, acc.z); acc.w = fmaf(g3, v3.w, acc.w);
        }
        for (; t < T; ++t) {
            const float4 v0 = src[(size_t)t * (XW / 4)];
            const float g0 = s_de[1][t];
            acc.x = fmaf(g0, v0.x, acc.x); acc.y = fmaf(g0, v0.y, acc.y); acc.z = fmaf(g0, v0.z, acc.z); acc.w = fmaf(g0, v0.w, acc.w);
        }
        reinterpret_cast<float4*>(dq + (size_t)b * XW)[c] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ time tables
__global__ void clip_time_tables_kernel(const float* __restrict__ month, const float* __restrict__ day,
                                        const float* __restrict__ week, const float* __restrict__ hour,
                                        const float* __restrict__ minute, float* __restrict__ ct_tab,
                                        float* __restrict__ ct_scale) {
    PDL_ENTER();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= NB) return;
    int k = 0;
    while (r >= kBinOff[k + 1]) ++k;
    const float* tabs[5] = {month, day, week, hour, minute};
    const float2 v = reinterpret_cast<const float2*>(tabs[k] + (size_t)(r - kBinOff[k]) * TH)[lane];
    const float s = clip_scale(warp_sum(v.x * v.x + v.y * v.y));
    reinterpret_cast<float2*>(ct_tab + (size_t)r * TH)[lane] = make_float2(v.x * s, v.y * s);
    if (lane == 0) ct_scale[r] = s;
}

// exact fp32 score of item n for session vectors in shared memory (warp-cooperative, fixed order).
// S[b,n] = a_ic . [item[n+1] | content[n+1]] + sum_k Tq[b, off_k + mwdhm[n,k]]     (model_combine.py:132-138)
__device__ __forceinline__ float exact_score(const float* s_aic, const float* s_tq, const float* __restrict__ item,
                                             const float* __restrict__ content, const int32_t* __restrict__ mwdhm,
                                             int n, int lane) {
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
    acc = warp_sum(acc);
    float tsum = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) tsum += s_tq[kBinOff[k] + mwdhm[(size_t)n * 5 + k]];
    return acc + tsum;
}

// ------------------------------------------------------------------------------------------------ (3b) query
__global__ void __launch_bounds__(256)
build_query_kernel(const float* __restrict__ a_ic, const float* __restrict__ a_pt, const float* __restrict__ ct_tab,
                   const float* __restrict__ item, const float* __restrict__ content,
                   const int32_t* __restrict__ mwdhm, const int32_t* __restrict__ label, float* __restrict__ Tq,
                   __nv_bfloat16* __restrict__ Q, float* __restrict__ c_ref, int B) {
    PDL_ENTER();
    __shared__ float s_aic[XW], s_apt[PW], s_tq[NB + 1];
    const int b = blockIdx.x;
    __nv_bfloat16* qr = Q + (size_t)b * TCAR_KEXT;
    if (b >= B) {
        for (int c = threadIdx.x; c < TCAR_KEXT; c += blockDim.x) qr[c] = __float2bfloat16(0.f);
        return;
    }
    for (int c = threadIdx.x; c < XW; c += blockDim.x) s_aic[c] = a_ic[(size_t)b * XW + c];
    for (int c = threadIdx.x; c < PW; c += blockDim.x) s_apt[c] = a_pt[(size_t)b * PW + c];
    __syncthreads();
    {
        // Tq[b, r] = a_pt[b, 64 k : 64 k + 64] . clip(table_k)[r]: one warp per table row, coalesced 256-byte row
        // reads, six rows in flight per warp (the 139 rows are spread over the 8 warps)
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll 6
        for (int r = w; r < NB; r += 8) {
            const int k = (r >= 13) + (r >= 45) + (r >= 53) + (r >= 78);
            const float v0 = __ldg(ct_tab + (size_t)r * TH + lane), v1 = __ldg(ct_tab + (size_t)r * TH + 32 + lane);
            const float acc = warp_sum(fmaf(v0, s_apt[k * TH + lane], v1 * s_apt[k * TH + 32 + lane]));
            if (lane == 0) { s_tq[r] = acc; Tq[(size_t)b * NB + r] = acc; }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < TCAR_KEXT; c += blockDim.x) {
        const float v = c < XW ? s_aic[c] : (c < XW + NB ? s_tq[c - XW] : 0.f);
        qr[c] = __float2bfloat16(v);
    }
    if (threadIdx.x < 32) {
        const float s = exact_score(s_aic, s_tq, item, content, mwdhm, label[b], threadIdx.x);
        if (threadIdx.x == 0) c_ref[b] = s;
    }
}

// ------------------------------------------------------------------------------------------------ (4a) CE finish
// 8 rows per CTA (a quarter-warp = 8 consecutive rows = one 32-byte sector of a partial, the four quarter-warps take
// consecutive partials), 32 warps stride over the partials with 8 independent loads in flight per thread, fixed combine
// order.  (16 rows per CTA left 116 of the 148 SMs idle: 17 us for 23 MB of partials at B = 512.)
//
// Overflow guard of the softmax (TF's sparse_softmax_cross_entropy_with_logits subtracts the row maximum,
// model_combine.py:145; the scoring kernel shifts by the label score):
//   pass 0  sums only (no guard)
//   pass 1  sums, and rowmax[b] = max over the tiles of pmax (largest exponent argument of the row, log2 units)
//   pass 2  after the scoring kernel re-ran with `rowmax` (rows above TCAR_EXP_LIMIT2 shifted by it): those rows are
//           summed again and ce[b] = log(sum) + rowmax[b] ln 2; every other row keeps its pass-1 result, and CTAs
//           without such a row exit at once
//   pass 3  maxima only (rowmax[b]; catalog-sharded step, where the sums travel with dQ)
// blockIdx.y = session group: partials of group g part_stride floats apart, outputs (sumexp, rowmax) out_stride apart
struct GroupRows { int n[TCAR_MAX_PEERS]; long long part_stride; long long out_stride; };

__global__ void __launch_bounds__(1024)
ce_finish_kernel(const float* __restrict__ part, const float* __restrict__ pmax, float* __restrict__ sumexp,
                 float* __restrict__ ce, float* __restrict__ rowmax, int n_tiles, int B, int sum_stride, int pass,
                 const __grid_constant__ GroupRows gr) {
    PDL_ENTER();
    __shared__ float s[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + (lane & 7);
    if (gridDim.y > 1 || gr.part_stride) {
        // several session groups in one launch: group g's partials lie part_stride floats apart, its row maxima go
        // to rowmax + g * 512
        B = gr.n[blockIdx.y];
        if (part) part += (size_t)blockIdx.y * gr.part_stride;
        if (pmax) pmax += (size_t)blockIdx.y * gr.part_stride;
        if (sumexp) sumexp += (size_t)blockIdx.y * gr.out_stride;
        if (rowmax) rowmax += (size_t)blockIdx.y * gr.out_stride;
        if ((int)blockIdx.x * 8 >= B) return;
    }
    bool redo = false;
    if (pass == 2) {
        redo = b < B && rowmax[b] > TCAR_EXP_LIMIT2;
        if (!__syncthreads_or(redo)) return;
    }
    float tot = 0.f;
    if (pass != 3) {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 0.f;
        if (b < B) {
            const float* src = part + b;
            int t = 4 * w + (lane >> 3);
            for (; t + 7 * 128 < n_tiles; t += 8 * 128) {
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] += src[(size_t)(t + 128 * k) * TCAR_QROWS];
            }
            for (; t < n_tiles; t += 128) a[0] += src[(size_t)t * TCAR_QROWS];
        }
        s[w][lane] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        __syncthreads();
        if (w == 0 && lane < 8)
            for (int i = 0; i < 32; ++i) tot += (s[i][lane] + s[i][lane + 8]) + (s[i][lane + 16] + s[i][lane + 24]);
        __syncthreads();
    }
    if (pass == 1 || pass == 3) {
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (b < B) {
            const float* src = pmax + b;
            int t = 4 * w + (lane >> 3);
            for (; t + 3 * 128 < n_tiles; t += 4 * 128) {
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], src[(size_t)(t + 128 * k) * TCAR_QROWS]);
            }
            for (; t < n_tiles; t += 128) m[0] = fmaxf(m[0], src[(size_t)t * TCAR_QROWS]);
        }
        s[w][lane] = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
        __syncthreads();
        if (w == 0 && lane < 8 && b < B) {
            float mx = -INFINITY;
            for (int i = 0; i < 32; ++i)
                mx = fmaxf(mx, fmaxf(fmaxf(s[i][lane], s[i][lane + 8]), fmaxf(s[i][lane + 16], s[i][lane + 24])));
            rowmax[b] = mx;
        }
    }
    if (pass != 3 && w == 0 && lane < 8 && b < B && (pass != 2 || redo)) {
        sumexp[(size_t)b * sum_stride] = tot;
        if (ce) ce[b] = pass == 2 ? fmaf(rowmax[b], 0.6931471805599453f, logf(tot)) : logf(tot);
    }
}

// ------------------------------------------------------------------------------------------------ (5a') column jobs
// Up to TCAR_COL_JOBS independent column reductions in one launch (blockIdx.y = job): 32 columns per CTA (lane =
// column), 32 warps stride over the rows with four rows in flight per thread; column sums combined in a fixed order.
//   mode 0 / 1: dz[r,c] = a[r,c] * act'(y[r,c]) (tanh: 1 - y^2, relu: y > 0), out[c] = sum_r dz[r,c]
//   mode 2    : out[c] = sum_r y[r,c] * a[r]          (weight-vector gradients of count_alpha_*, modules.py:99,134)
// Long jobs (rows > 512, given a scratch) are split over blockIdx.z: every CTA parks its partial sums in the scratch,
// and the CTA that arrives last (ticket counter behind the partials) adds them in split order -- the result does not
// depend on which CTA that is.
struct ColJobs { tcar_col_job j[TCAR_COL_JOBS]; int rsplit[TCAR_COL_JOBS]; };

__device__ __forceinline__ float col_term(const tcar_col_job& q, int r, int c) {
    const size_t i = (size_t)r * q.ld + c;
    const float yv = q.y[i];
    if (q.mode == 2) return yv * q.a[r];
    const float d = q.a[i] * (q.mode == 0 ? (1.f - yv * yv) : (yv > 0.f ? 1.f : 0.f));
    q.dz[i] = d;
    return d;
}

__global__ void __launch_bounds__(1024)
col_jobs_kernel(const __grid_constant__ ColJobs jobs) {
    PDL_ENTER();
    const tcar_col_job& q = jobs.j[blockIdx.y];
    const int rsplit = jobs.rsplit[blockIdx.y];
    if ((int)blockIdx.x * 32 >= q.cols || (int)blockIdx.z >= rsplit) return;
    __shared__ float s[32][33];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int per = (q.rows + rsplit - 1) / rsplit;
    const int r0 = blockIdx.z * per, r1 = min(r0 + per, q.rows);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < q.cols) {
        int r = r0 + w;
        for (; r + 96 < r1; r += 128) {
            a0 += col_term(q, r, c); a1 += col_term(q, r + 32, c);
            a2 += col_term(q, r + 64, c); a3 += col_term(q, r + 96, c);
        }
        for (; r < r1; r += 32) a0 += col_term(q, r, c);
    }
    s[w][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    float t = 0.f;
    if (w == 0) {
        for (int i = 0; i < 32; ++i) t += s[i][lane];
    }
    if (rsplit == 1) {
        if (w == 0 && c < q.cols) q.out[c] = t;
        return;
    }
    // scratch: [rsplit][cols] partials, then one ticket counter per column block
    float* part = q.scratch;
    int* ticket = reinterpret_cast<int*>(q.scratch + (size_t)rsplit * q.cols) + blockIdx.x;
    if (w == 0) {
        if (c < q.cols) part[(size_t)blockIdx.z * q.cols + c] = t;
        __threadfence();
        __syncwarp();
        if (lane == 0) s_last = atomicAdd(ticket, 1) == rsplit - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (w == 0) {
        __threadfence();
        if (c < q.cols) {
            float tot = 0.f;
            for (int z = 0; z < rsplit; ++z) tot += __ldcg(part + (size_t)z * q.cols + c);
            q.out[c] = tot;
        }
        if (lane == 0) *ticket = 0;          // ready for the next launch
    }
}

// ------------------------------------------------------------------------------------------------ (4b) neg loss
// ce == NULL: the cross loss is not known yet (the kernel runs beside the scoring GEMM, see Seq2SeqAttNN.forward_train);
// loss[] is then left to tcar_loss_combine.  No shared memory beyond the reduction scratch (the two columns a thread owns
// stay in registers), so that a CTA fits beside a scoring-GEMM CTA that holds all but 1.7 KB of the SM's shared memory.
__global__ void __launch_bounds__(256)
neg_loss_kernel(const float* __restrict__ a_ic, const float* __restrict__ item, const float* __restrict__ content,
                const int32_t* __restrict__ neg, const float* __restrict__ ce, float* __restrict__ negloss,
                float* __restrict__ loss, float* __restrict__ coef, float* __restrict__ dA_neg, int B, int Nn) {
    PDL_ENTER();
    __shared__ float red[32];
    const int b = blockIdx.x;
    float partial = 0.f;
    float sv[2] = {0.f, 0.f};                       // columns threadIdx.x and threadIdx.x + 256 (XW = 500 <= 512)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = threadIdx.x + k * 256;
        if (c < XW) {
            const float* tab = c < H ? item : content;
            const int cc = c < H ? c : c - H;
            float v = 0.f;
            for (int j = 0; j < Nn; ++j) v += tab[((size_t)neg[(size_t)b * Nn + j] + 1) * HP + cc];
            sv[k] = v;
            partial = fmaf(v, a_ic[(size_t)b * XW + c], partial);
        }
    }
    const float z = block_sum(partial, red);
    const float s = 1.f / (1.f + expf(-z));
    const float u = (1.f - s) + 1e-24f;
    const float cf = 0.01f * s * (1.f - s) / u;
    if (threadIdx.x == 0) {
        const float nl = -logf(u);
        negloss[b] = nl;
        if (ce) loss[b] = ce[b] + 0.01f * nl;
        coef[b] = cf;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = threadIdx.x + k * 256;
        if (c < XW) dA_neg[(size_t)b * XW + c] = cf * sv[k];
    }
}

// Catalog-sharded step: the softmax sums of a rank's sessions arrive in a strided column (the pad column of the
// reduce-scattered dQ): sumexp[b] = sums[b * stride], ce[b] = log(sumexp[b]) (+ shift ln 2 for rows the overflow guard
// shifted by rowmax[b] > TCAR_EXP_LIMIT2, log2 units).  One launch instead of six elementwise torch kernels.
__global__ void __launch_bounds__(256)
ce_from_sums_kernel(const float* __restrict__ sums, int stride, const float* __restrict__ rowmax,
                    float* __restrict__ sumexp, float* __restrict__ ce, int B) {
    PDL_ENTER();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float s = sums[(size_t)b * stride];
    sumexp[b] = s;
    float c = logf(s);
    if (rowmax) {
        const float m = rowmax[b];
        if (m > TCAR_EXP_LIMIT2) c = fmaf(m, 0.6931471805599453f, c);
    }
    ce[b] = c;
}

// loss = cross loss + 0.01 * negative-feedback loss (model_combine.py:146-147), for tcar_neg_loss(ce = NULL)
__global__ void __launch_bounds__(256)
loss_combine_kernel(const float* __restrict__ ce, const float* __restrict__ negloss, float* __restrict__ loss, int B) {
    PDL_ENTER();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) loss[b] = ce[b] + 0.01f * negloss[b];
}

// ------------------------------------------------------------------------------------------------ (3e) finish
__global__ void __launch_bounds__(256)
score_bwd_finish_kernel(const float* __restrict__ dq_raw, const float* __restrict__ sumexp,
                        const float* __restrict__ dA_neg, const float* __restrict__ a_ic,
                        const float* __restrict__ ct_tab, const float* __restrict__ item,
                        const float* __restrict__ content, const int32_t* __restrict__ mwdhm,
                        const int32_t* __restrict__ label, float* __restrict__ d_a_ic, float* __restrict__ d_a_pt,
                        float* __restrict__ dTq, __nv_bfloat16* __restrict__ Qs, int B) {
    PDL_ENTER();
    __shared__ float s_dt[NB + 1];
    const int b = blockIdx.x;
    __nv_bfloat16* qs = Qs + (size_t)b * HP;
    if (b >= B) {
        for (int c = threadIdx.x; c < HP; c += blockDim.x) qs[c] = __float2bfloat16(0.f);
        return;
    }
    const float inv = 1.f / sumexp[b];
    const int lab = label[b];
    for (int c = threadIdx.x; c < XW; c += blockDim.x) {
        const float lv = c < H ? item[((size_t)lab + 1) * HP + c] : content[((size_t)lab + 1) * HP + (c - H)];
        d_a_ic[(size_t)b * XW + c] = dq_raw[(size_t)b * TCAR_KEXT + c] * inv - lv + dA_neg[(size_t)b * XW + c];
    }
    for (int c = threadIdx.x; c < HP; c += blockDim.x)
        qs[c] = __float2bfloat16(c < H ? a_ic[(size_t)b * XW + c] * inv : 0.f);
    if (threadIdx.x < NB) {
        const int r = threadIdx.x;
        int k = 0;
        while (r >= kBinOff[k + 1]) ++k;
        const float onehot = (mwdhm[(size_t)lab * 5 + k] == r - kBinOff[k]) ? 1.f : 0.f;
        const float v = dq_raw[(size_t)b * TCAR_KEXT + XW + r] * inv - onehot;
        s_dt[r] = v;
        dTq[(size_t)b * NB + r] = v;
    }
    __syncthreads();
    // d a_pt[b, 64 k + d] = sum_{r in table k} dTq[b, r] clip(table_k)[r, d]: four independent partial sums per
    // thread so that the (L1-resident) table reads overlap
    for (int c = threadIdx.x; c < PW; c += blockDim.x) {
        const int k = c / TH, d = c % TH;
        const int r1 = kBinOff[k + 1];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int r = kBinOff[k];
        for (; r + 4 <= r1; r += 4) {
            const float t0 = __ldg(ct_tab + (size_t)r * TH + d), t1 = __ldg(ct_tab + (size_t)(r + 1) * TH + d);
            const float t2 = __ldg(ct_tab + (size_t)(r + 2) * TH + d), t3 = __ldg(ct_tab + (size_t)(r + 3) * TH + d);
            a0 = fmaf(s_dt[r], t0, a0); a1 = fmaf(s_dt[r + 1], t1, a1);
            a2 = fmaf(s_dt[r + 2], t2, a2); a3 = fmaf(s_dt[r + 3], t3, a3);
        }
        for (; r < r1; ++r) a0 = fmaf(s_dt[r], __ldg(ct_tab + (size_t)r * TH + d), a0);
        d_a_pt[(size_t)b * PW + c] = (a0 + a1) + (a2 + a3);
    }
}

// ------------------------------------------------------------------------------------------------ (5a) small tables
// clip Jacobian of y = x / max(||x||, 1):  dx = dy (||x|| <= 1)   else   s (dy - y (y . dy)),  s = 1/||x||, y = s x
//
// One CTA per table row (40 position rows, 139 publish-time rows, 11 dwell rows), 16 warps.  A time / dwell row scans
// the index column of its table (lane <-> click, 512 clicks per sweep of the CTA), and every warp adds the gradient
// rows of its matching clicks in increasing click order; the 16 warp partials, the click-context matches (week / hour
// rows), and the scoring-side term sum_b dTq[b,r] a_pt[b,:] are then combined in a fixed order and the clip Jacobian
// is applied.  No atomics, no scratch: the result does not depend on scheduling.
constexpr int TGD_THREADS = 512;
constexpr int TGD_WARPS = TGD_THREADS / 32;

// adds src[m * pitch + lane], src[m * pitch + 32 + lane] for every m in [0, count) with key[m] == want, visiting the
// warp's clicks (32 consecutive ones every 32 * TGD_WARPS) in increasing order; up to four matches are fetched at once
__device__ __forceinline__ void match_accumulate(const int32_t* __restrict__ key, int want, int count,
                                                 const float* __restrict__ src, size_t pitch, int w, int lane,
                                                 float& acc0, float& acc1) {
    for (int m0 = w * 32; m0 < count; m0 += 32 * TGD_WARPS) {
        const int m = m0 + lane;
        unsigned bits = __ballot_sync(0xffffffffu, m < count && __ldg(key + m) == want);
        while (bits) {
            int j[4];
            float v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                j[u] = bits ? __ffs(bits) - 1 : -1;
                if (bits) bits &= bits - 1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* r = src + (size_t)(m0 + max(j[u], 0)) * pitch;
                v0[u] = j[u] >= 0 ? r[lane] : 0.f;
                v1[u] = j[u] >= 0 ? r[32 + lane] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc0 += v0[u]; acc1 += v1[u]; }
        }
    }
}

__global__ void __launch_bounds__(TGD_THREADS)
table_grads_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ ctx, const float* __restrict__ dXi,
                   const float* __restrict__ dP, const float* __restrict__ dD, const float* __restrict__ dCT,
                   const float* __restrict__ dTq, const float* __restrict__ a_pt, const float* __restrict__ pos,
                   const float* __restrict__ month, const float* __restrict__ day, const float* __restrict__ week,
                   const float* __restrict__ hour, const float* __restrict__ minute, const float* __restrict__ dur,
                   float* __restrict__ g_pos, float* __restrict__ g_month, float* __restrict__ g_day,
                   float* __restrict__ g_week, float* __restrict__ g_hour, float* __restrict__ g_minute,
                   float* __restrict__ g_dur, int B, int T) {
    PDL_ENTER();
    __shared__ __align__(16) float s_part[TGD_WARPS][256];
    __shared__ float s_g[256];
    __shared__ float red[32];
    const int row = blockIdx.x;  // [0,40) pos, [40,179) time bins, [179,190) duration
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int M = B * T;
    const float* x;
    float* g;
    int width;
    if (row < TCAR_MAXT) {
        // position row t: sum over the sessions of dXi[b * T + t, :] -- thread <-> (float4 column, session group)
        width = H;
        x = pos + (size_t)row * H;
        g = g_pos + (size_t)row * H;
        const int c4 = tid & 63, grp = tid >> 6;                    // 64 float4 columns x 8 session groups
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        if (row < T) {
            const float4* src = reinterpret_cast<const float4*>(dXi + (size_t)row * HP) + c4;
            const size_t step = (size_t)T * (HP / 4);               // float4 stride between sessions
            int bb = grp;
            for (; bb + 24 < B; bb += 32) {
                const float4 u0 = src[(size_t)bb * step], u1 = src[(size_t)(bb + 8) * step];
                const float4 u2 = src[(size_t)(bb + 16) * step], u3 = src[(size_t)(bb + 24) * step];
                a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
                a1.x += u1.x; a1.y += u1.y; a1.z += u1.z; a1.w += u1.w;
                a2.x += u2.x; a2.y += u2.y; a2.z += u2.z; a2.w += u2.w;
                a3.x += u3.x; a3.y += u3.y; a3.z += u3.z; a3.w += u3.w;
            }
            for (; bb < B; bb += 8) {
                const float4 u0 = src[(size_t)bb * step];
                a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
            }
        }
        float4 t4;
        t4.x = (a0.x + a1.x) + (a2.x + a3.x); t4.y = (a0.y + a1.y) + (a2.y + a3.y);
        t4.z = (a0.z + a1.z) + (a2.z + a3.z); t4.w = (a0.w + a1.w) + (a2.w + a3.w);
        reinterpret_cast<float4*>(&s_part[grp][0])[c4] = t4;
        __syncthreads();
        if (tid < 256) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += s_part[i][tid];
            s_g[tid] = tid < H ? acc : 0.f;                          // dXi pad columns are never written
        }
    } else {
        width = TH;
        const int r = row - TCAR_MAXT;
        float acc0 = 0.f, acc1 = 0.f;
        int k = -1;
        if (r < NB) {
            k = (r >= 13) + (r >= 45) + (r >= 53) + (r >= 78);
            const int rr = r - kBinOff[k];
            const float* tabs[5] = {month, day, week, hour, minute};
            float* gs[5] = {g_month, g_day, g_week, g_hour, g_minute};
            x = tabs[k] + (size_t)rr * TH;
            g = gs[k] + (size_t)rr * TH;
            match_accumulate(idx + (size_t)(k + 1) * M, rr, M, dP + k * TH, PW, w, lane, acc0, acc1);
            // click-time context (sampler.py:106-107): week table by ctx[0][b], hour table by ctx[1][b]
            if (k == 2) match_accumulate(ctx, rr, B, dCT, 2 * TH, w, lane, acc0, acc1);
            if (k == 3) match_accumulate(ctx + B, rr, B, dCT + TH, 2 * TH, w, lane, acc0, acc1);
        } else {
            const int rr = r - NB;
            x = dur + (size_t)rr * TH;
            g = g_dur + (size_t)rr * TH;
            match_accumulate(idx + (size_t)6 * M, rr, M, dD, TH, w, lane, acc0, acc1);
        }
        s_part[w][lane] = acc0;
        s_part[w][32 + lane] = acc1;
        // scoring side: d/d clip(table)[r] of sum_b Tq[b,r] = sum_b dTq[b,r] a_pt[b, 64k:64k+64]; thread <-> (d, group)
        float sc = 0.f;
        if (k >= 0) {
            const int d = tid & 63, grp = tid >> 6;                  // 8 session groups
            const float* dq = dTq + r;
            const float* ap = a_pt + k * TH + d;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int bb = grp;
            for (; bb + 24 < B; bb += 32) {
                a0 = fmaf(dq[(size_t)bb * NB], ap[(size_t)bb * PW], a0);
                a1 = fmaf(dq[(size_t)(bb + 8) * NB], ap[(size_t)(bb + 8) * PW], a1);
                a2 = fmaf(dq[(size_t)(bb + 16) * NB], ap[(size_t)(bb + 16) * PW], a2);
                a3 = fmaf(dq[(size_t)(bb + 24) * NB], ap[(size_t)(bb + 24) * PW], a3);
            }
            for (; bb < B; bb += 8) a0 = fmaf(dq[(size_t)bb * NB], ap[(size_t)bb * PW], a0);
            sc = (a0 + a1) + (a2 + a3);
        }
        // thread (d, grp) parks its scoring partial in row grp, columns 128 + d
        if (k >= 0) s_part[tid >> 6][128 + (tid & 63)] = sc;
        __syncthreads();
        if (tid < 64) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < TGD_WARPS; ++i) acc += s_part[i][tid];
            if (k >= 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += s_part[i][128 + tid];
            }
            s_g[tid] = acc;
        }
    }
    __syncthreads();
    const int c = tid;
    const float xv = c < width ? x[c] : 0.f;
    const float gv = c < width ? s_g[c] : 0.f;
    const float sq = block_sum(xv * xv, red);
    const float n = sqrtf(sq);
    if (n > 1.f) {
        const float sI = 1.f / n;
        const float ydg = block_sum(xv * sI * gv, red);
        if (c < width) g[c] = sI * (gv - xv * sI * ydg);
    } else if (c < width) {
        g[c] = gv;
    }
}

// ------------------------------------------------------------------------------------------------ (5b) scatter-add
constexpr float kFix = 1099511627776.f;        // 2^40
constexpr float kInvFix = 1.f / 1099511627776.f;

__device__ __forceinline__ int hash_slot(int32_t* keys, int row, int mask) {
    uint32_t h = (uint32_t)row * 2654435761u;
    int slot = (int)((h >> 7) & (uint32_t)mask);
    while (true) {
        const int prev = atomicCAS(&keys[slot], -1, row);
        if (prev == -1 || prev == row) return slot;
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ int entry_row(int e, int M, int B, const int32_t* __restrict__ seq,
                                         const int32_t* __restrict__ label, const int32_t* __restrict__ neg) {
    if (e < M) return seq[e];
    if (e < M + B) return label[e - M] + 1;
    return neg[e - M - B] + 1;
}

// pass 1 -- one thread per sparse entry (clicks [0,M), labels [M,M+B), negatives [M+B, M+B+B*Nn)): claim the hash
// slot of its item row and count how many entries share it.
__device__ __forceinline__ void scatter_count_body(int e, const int32_t* __restrict__ seq,
                                                   const int32_t* __restrict__ label, const int32_t* __restrict__ neg,
                                                   int32_t* __restrict__ keys, int32_t* __restrict__ cnt,
                                                   int32_t* __restrict__ entry_slot, int mask, int B, int T, int Nn,
                                                   int row_lo, int row_hi) {
    const int M = B * T;
    if (e >= M + B + B * Nn) return;
    const int row = entry_row(e, M, B, seq, label, neg);
    // rows outside [row_lo, row_hi) belong to another catalog shard (tcar_scatter_add_rows_range): no slot
    if (row < row_lo || row >= row_hi) { entry_slot[e] = -1; return; }
    const int slot = hash_slot(keys, row, mask);
    atomicAdd(&cnt[slot], 1);
    entry_slot[e] = slot;
}

__global__ void __launch_bounds__(256)
scatter_count_kernel(const int32_t* __restrict__ seq, const int32_t* __restrict__ label,
                     const int32_t* __restrict__ neg, int32_t* __restrict__ keys, int32_t* __restrict__ cnt,
                     int32_t* __restrict__ entry_slot, int mask, int B, int T, int Nn, int row_lo, int row_hi) {
    PDL_ENTER();
    scatter_count_body(blockIdx.x * blockDim.x + threadIdx.x, seq, label, neg, keys, cnt, entry_slot, mask, B, T, Nn,
                       row_lo, row_hi);
}

// The sparse rows of SEVERAL session groups (the ranks of a catalog-sharded step) against ONE hash table: blockIdx.y =
// group, packed batch of group g at ids + g * ids_stride ([7*B*T idx | 2*B ctx | B label | B*Nn neg]), payload at
// payload + g * pay_stride ([a_ic 512x500 | coef 512 | dXi B*T x 256]), its entries' slots at entry_slot + g * es_stride.
// Counting ALL groups before any accumulation makes "touched once" a global property, so the in-place path stays
// race-free and the shared rows still add up in order-independent fixed point: three launches per step instead of
// three per source rank.
struct ScatterGroups {
    const int32_t* ids;
    const float* payload;
    long long ids_stride, pay_stride, es_stride;
    int n[TCAR_MAX_PEERS];
};

__global__ void __launch_bounds__(256)
scatter_count_groups_kernel(const __grid_constant__ ScatterGroups sg, int32_t* __restrict__ keys,
                            int32_t* __restrict__ cnt, int32_t* __restrict__ entry_slot, int mask, int T, int Nn,
                            int row_lo, int row_hi) {
    PDL_ENTER();
    const int g = blockIdx.y, B = sg.n[g];
    if (B <= 0) return;
    const int32_t* base = sg.ids + (size_t)g * sg.ids_stride;
    const size_t M = (size_t)B * T;
    scatter_count_body(blockIdx.x * blockDim.x + threadIdx.x, base, base + 7 * M + 2 * (size_t)B,
                       base + 7 * M + 3 * (size_t)B, keys, cnt, entry_slot + (size_t)g * sg.es_stride, mask, B, T, Nn,
                       row_lo, row_hi);
}

// pass 2 -- one warp per entry.  A row touched by exactly one entry (the common case: uniform negatives, labels, tail
// items) is updated in place with plain 128-bit read-modify-writes; rows shared by several entries accumulate in exact
// int64 fixed point (2^-40) so that the sum does not depend on the arrival order.
__device__ __forceinline__ void scatter_accum_body(int e, int lane, const int32_t* __restrict__ seq,
                                                   const int32_t* __restrict__ label, const int32_t* __restrict__ neg,
                                                   const float* __restrict__ dXi, const float* __restrict__ a_ic,
                                                   const float* __restrict__ coef, const float* __restrict__ item,
                                                   float* __restrict__ g_item, const int32_t* __restrict__ cnt,
                                                   const int32_t* __restrict__ entry_slot,
                                                   unsigned long long* __restrict__ acc, float* __restrict__ slot_sq,
                                                   int B, int T, int Nn) {
    const int M = B * T;
    if (e >= M + B + B * Nn) return;
    const int slot = entry_slot[e];
    if (slot < 0) return;                    // row of another catalog shard
    const int row = entry_row(e, M, B, seq, label, neg);
    float val[8];
    bool jac_on = false;
    float jac_s = 1.f, jac_ydy = 0.f, sc_b = 0.f;
    int b_idx = 0;
    if (e < M) {
        const float4* ir = reinterpret_cast<const float4*>(item + (size_t)row * HP);
        const float4 i0 = __ldg(ir + lane), i1 = __ldg(ir + 32 + lane);
        const float4* dr = reinterpret_cast<const float4*>(dXi + (size_t)e * HP);
        const float4 d0 = __ldg(dr + lane), d1 = __ldg(dr + 32 + lane);
        const float xv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        float dy[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float sq = 0.f, xdy = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = (j < 4) ? lane * 4 + j : 128 + lane * 4 + (j - 4);
            if (c >= H) dy[j] = 0.f;          // dXi pad columns are never written
            sq = fmaf(xv[j], xv[j], sq);
            xdy = fmaf(xv[j], dy[j], xdy);
        }
        sq = warp_sum(sq);
        xdy = warp_sum(xdy);
        const float n = sqrtf(sq);
        if (n > 1.f) {
            const float s = 1.f / n;
            const float ydy = xdy * s;  // y . dy
            jac_on = true; jac_s = s; jac_ydy = ydy;
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] = s * (dy[j] - xv[j] * s * ydy);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] = dy[j];
        }
    } else {
        int b;
        float scale;
        if (e < M + B) { b = e - M; scale = -1.f; }
        else { b = (e - M - B) / Nn; scale = coef[b]; }
        b_idx = b; sc_b = scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = (j < 4) ? lane * 4 + j : 128 + lane * 4 + (j - 4);
            val[j] = c < H ? scale * a_ic[(size_t)b * XW + c] : 0.f;
        }
    }
    if (cnt[slot] == 1) {
        float4* dst = reinterpret_cast<float4*>(g_item + (size_t)row * HP);
        float4 o0 = dst[lane], o1 = dst[32 + lane];
        const float4 n0 = make_float4(o0.x + val[0], o0.y + val[1], o0.z + val[2], o0.w + val[3]);
        const float4 n1 = make_float4(o1.x + val[4], o1.y + val[5], o1.z + val[6], o1.w + val[7]);
        dst[lane] = n0;
        dst[32 + lane] = n1;
        // change of the squared gradient norm caused by this row (see tcar_sqnorm_combine)
        float d = (n0.x * n0.x - o0.x * o0.x) + (n0.y * n0.y - o0.y * o0.y) + (n0.z * n0.z - o0.z * o0.z) +
                  (n0.w * n0.w - o0.w * o0.w) + (n1.x * n1.x - o1.x * o1.x) + (n1.y * n1.y - o1.y * o1.y) +
                  (n1.z * n1.z - o1.z * o1.z) + (n1.w * n1.w - o1.w * o1.w);
        d = warp_sum(d);
        if (lane == 0 && slot_sq) slot_sq[slot] = d;
    } else {
        // shared row: exact fixed-point accumulation.  Values are re-derived with lane <-> consecutive column (the
        // operands were just fetched, so these are L1 hits): one warp-wide atomic instruction then covers 256
        // contiguous bytes (8 sectors) instead of 32 scattered sectors.
        unsigned long long* dst = acc + (size_t)slot * HP;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = k * 32 + lane;
            if (c < H) {
                float vv;
                if (e < M) {
                    const float dyc = dXi[(size_t)e * HP + c];
                    vv = jac_on ? jac_s * (dyc - item[(size_t)row * HP + c] * jac_s * jac_ydy) : dyc;
                } else {
                    vv = sc_b * a_ic[(size_t)b_idx * XW + c];
                }
                atomicAdd(dst + c, (unsigned long long)__float2ll_rn(vv * kFix));
            }
        }
    }
}

__global__ void __launch_bounds__(256)
scatter_accum_kernel(const int32_t* __restrict__ seq, const int32_t* __restrict__ label,
                     const int32_t* __restrict__ neg, const float* __restrict__ dXi, const float* __restrict__ a_ic,
                     const float* __restrict__ coef, const float* __restrict__ item, float* __restrict__ g_item,
                     const int32_t* __restrict__ cnt, const int32_t* __restrict__ entry_slot,
                     unsigned long long* __restrict__ acc, float* __restrict__ slot_sq, int B, int T, int Nn) {
    PDL_ENTER();
    scatter_accum_body((blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31, seq, label, neg, dXi, a_ic, coef,
                       item, g_item, cnt, entry_slot, acc, slot_sq, B, T, Nn);
}

__global__ void __launch_bounds__(256)
scatter_accum_groups_kernel(const __grid_constant__ ScatterGroups sg, const float* __restrict__ item,
                            float* __restrict__ g_item, const int32_t* __restrict__ cnt,
                            const int32_t* __restrict__ entry_slot, unsigned long long* __restrict__ acc,
                            float* __restrict__ slot_sq, int T, int Nn) {
    PDL_ENTER();
    const int g = blockIdx.y, B = sg.n[g];
    if (B <= 0) return;
    const int32_t* base = sg.ids + (size_t)g * sg.ids_stride;
    const float* pay = sg.payload + (size_t)g * sg.pay_stride;
    const size_t M = (size_t)B * T;
    scatter_accum_body((blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31, base,
                       base + 7 * M + 2 * (size_t)B, Nn > 0 ? base + 7 * M + 3 * (size_t)B : nullptr,
                       pay + TCAR_QROWS * TCAR_XW + TCAR_QROWS, pay, pay + TCAR_QROWS * TCAR_XW, item, g_item, cnt,
                       entry_slot + (size_t)g * sg.es_stride, acc, slot_sq, B, T, Nn);
}

// pass 3 -- one thread per hash slot (most slots are empty or hold a row touched once: nothing to add); the slots of
// a warp that hold a shared row are then processed by the whole warp one after the other: add the exact integer sums
// to g_item, record the change of the squared norm, and restore the scratch (keys = -1, counts = 0, accumulators = 0).
__global__ void __launch_bounds__(256)
scatter_apply_kernel(int32_t* __restrict__ keys, int32_t* __restrict__ cnt, long long* __restrict__ acc,
                     float* __restrict__ g_item, float* __restrict__ slot_sq, int hash_size) {
    PDL_ENTER();
    const int slot0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31, lane = threadIdx.x & 31;
    const int mine = slot0 + lane;
    const int row_mine = mine < hash_size ? keys[mine] : -1;
    const int cnt_mine = row_mine >= 0 ? cnt[mine] : 0;
    if (mine < hash_size && row_mine < 0 && slot_sq) slot_sq[mine] = 0.f;
    unsigned shared = __ballot_sync(0xffffffffu, cnt_mine > 1);
    while (shared) {
        const int src = __ffs(shared) - 1;
        shared &= shared - 1;
        const int slot = slot0 + src;
        const int row = __shfl_sync(0xffffffffu, row_mine, src);
        long long* a = acc + (size_t)slot * HP;
        float* dst = g_item + (size_t)row * HP;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = k * 32 + lane;
            if (c < H) {
                const float o = dst[c];
                const float nv = o + (float)a[c] * kInvFix;
                dst[c] = nv;
                d += nv * nv - o * o;
                a[c] = 0;
            }
        }
        d = warp_sum(d);
        if (lane == 0 && slot_sq) slot_sq[slot] = d;
    }
    if (row_mine >= 0) { keys[mine] = -1; cnt[mine] = 0; }
}

}  // namespace tcar

using namespace tcar;
#define STREAM static_cast<cudaStream_t>(stream)
#define LAUNCH_RC() ((int)cudaGetLastError())

extern "C" int tcar_gather_fwd(const int32_t* idx, const int32_t* ctx, const float* item, const float* content,
                               const float* pos, const float* month, const float* day, const float* week,
                               const float* hour, const float* minute, const float* dur, float* X, float* P, float* D,
                               float* CT, int B, int T, void* stream) {
    if (B < 1 || T < 1 || T > TCAR_MAXT) return TCAR_ERR_ARG;
    const int warps = B * T + B;
    launch_pdl(gather_fwd_kernel, dim3((warps + 3) / 4), dim3(128), 0, STREAM, idx, ctx, item, content, pos, month, day, week, hour,
                                                           minute, dur, X, P, D, CT, B, T);
    return LAUNCH_RC();
}

extern "C" int tcar_assemble_batch(const int32_t* rows, const int32_t* seq, const int32_t* feats, const int32_t* ctx,
                                   int n_bucket, int B, int T, int Nn, const int32_t* neg_in, int item_num,
                                   unsigned long long seed, unsigned long long offset, int32_t* out, void* stream) {
    if (B < 1 || T < 1 || T > TCAR_MAXT || Nn < 0 || n_bucket < 1 || !rows || !seq || !feats || !ctx || !out ||
        (Nn > 0 && !neg_in && item_num < 1))
        return TCAR_ERR_ARG;
    const int total = 7 * B * T + 3 * B + B * Nn;
    launch_pdl(assemble_batch_kernel, dim3((total + 255) / 256), dim3(256), 0, STREAM, rows, seq, feats, ctx, n_bucket, B, T,
               Nn, neg_in, item_num, seed, offset, out);
    return LAUNCH_RC();
}

extern "C" int tcar_impression_negatives(const int32_t* rows, const int32_t* impr_off, const int32_t* impr_ids, int B,
                                         int Nn, int item_num, unsigned long long seed, unsigned long long offset,
                                         int32_t* neg_out, void* stream) {
    if (B < 1 || Nn < 1 || item_num < 1 || !rows || !impr_off || !impr_ids || !neg_out) return TCAR_ERR_ARG;
    launch_pdl(impression_negatives_kernel, dim3((B + 127) / 128), dim3(128), 0, STREAM, rows, impr_off, impr_ids, B,
               Nn, item_num, seed, offset, neg_out);
    return LAUNCH_RC();
}

extern "C" int tcar_pool_fwd(const float* X, const float* P, float* U1, float* U2, const float* q, const float* w_r,
                             const float* w_t, float* alpha, float* pooled, float* pooled_t, int B, int T,
                             void* stream) {
    if (B < 1 || T < 1 || T > TCAR_MAXT) return TCAR_ERR_ARG;
    launch_pdl(pool_fwd_kernel, dim3(B), dim3(256), 0, STREAM, X, P, U1, U2, q, w_r, w_t, alpha, pooled, pooled_t, B, T);
    return LAUNCH_RC();
}

extern "C" int tcar_pool_bwd(const float* X, const float* P, const float* S1, const float* S2, const float* q,
                             const float* w_r, const float* w_t, const float* alpha, const float* dpooled,
                             const float* dpooled_t, float* dU1, float* dU2, float* dXi, float* dP, float* dq,
                             float* de, int B, int T, void* stream) {
    if (B < 1 || T < 1 || T > TCAR_MAXT) return TCAR_ERR_ARG;
    launch_pdl(pool_bwd_kernel, dim3(B), dim3(256), 0, STREAM, X, P, S1, S2, q, w_r, w_t, alpha, dpooled, dpooled_t, dU1, dU2, dXi, dP,
                                           dq, de, B, T);
    return LAUNCH_RC();
}

extern "C" int tcar_clip_time_tables(const float* month, const float* day, const float* week, const float* hour,
                                     const float* minute, float* ct_tab, float* ct_scale, void* stream) {
    launch_pdl(clip_time_tables_kernel, dim3((NB * 32 + 255) / 256), dim3(256), 0, STREAM, month, day, week, hour, minute, ct_tab,
                                                                        ct_scale);
    return LAUNCH_RC();
}

extern "C" int tcar_build_query(const float* a_ic, const float* a_pt, const float* ct_tab, const float* item,
                                const float* content, const int32_t* mwdhm, const int32_t* label, float* Tq,
                                void* q_bf16, float* c_ref, int B, void* stream) {
    if (B < 1 || B > TCAR_QROWS) return TCAR_ERR_ARG;
    launch_pdl(build_query_kernel, dim3(TCAR_QROWS), dim3(256), 0, STREAM, a_ic, a_pt, ct_tab, item, content, mwdhm, label, Tq,
                                                       static_cast<__nv_bfloat16*>(q_bf16), c_ref, B);
    return LAUNCH_RC();
}

extern "C" int tcar_ce_finish(const float* rowsum_part, float* sumexp, float* ce, int n_tiles, int B, void* stream) {
    return tcar_ce_finish_guarded(rowsum_part, nullptr, sumexp, ce, nullptr, n_tiles, B, 0, stream);
}

extern "C" int tcar_ce_finish_guarded(const float* rowsum_part, const float* rowmax_part, float* sumexp, float* ce,
                                      float* rowmax, int n_tiles, int B, int pass, void* stream) {
    if (B < 1 || B > TCAR_QROWS || pass < 0 || pass > 3) return TCAR_ERR_ARG;
    if ((pass == 1 || pass == 3) && (!rowmax_part || !rowmax)) return TCAR_ERR_ARG;
    if (pass == 2 && !rowmax) return TCAR_ERR_ARG;
    if (pass != 3 && (!rowsum_part || !sumexp)) return TCAR_ERR_ARG;
    launch_pdl(ce_finish_kernel, dim3((B + 7) / 8), dim3(1024), 0, STREAM, rowsum_part, rowmax_part, sumexp, ce,
               rowmax, n_tiles, B, 1, pass, GroupRows{});
    return LAUNCH_RC();
}

extern "C" int tcar_rowmax_groups(const float* rowmax_part, long long part_stride, float* rowmax, int n_tiles,
                                  const int* n_rows, int groups, void* stream) {
    if (!rowmax_part || !rowmax || !n_rows || groups < 1 || groups > TCAR_MAX_PEERS || part_stride < 1)
        return TCAR_ERR_ARG;
    GroupRows gr = {};
    int bmax = 0;
    for (int g = 0; g < groups; ++g) {
        if (n_rows[g] > TCAR_QROWS) return TCAR_ERR_ARG;
        gr.n[g] = n_rows[g] > 0 ? n_rows[g] : 0;
        if (gr.n[g] > bmax) bmax = gr.n[g];
    }
    if (bmax == 0) return 0;
    gr.part_stride = part_stride;
    gr.out_stride = TCAR_QROWS;
    launch_pdl(ce_finish_kernel, dim3((bmax + 7) / 8, groups), dim3(1024), 0, STREAM,
               static_cast<const float*>(nullptr), rowmax_part, static_cast<float*>(nullptr),
               static_cast<float*>(nullptr), rowmax, n_tiles, bmax, 1, 3, gr);
    return LAUNCH_RC();
}

// Guarded softmax sums of several session groups in one launch (catalog-sharded evaluation): pass 1 / 2 of
// tcar_ce_finish_guarded per group; group g reads its partials at + g * part_stride and writes sumexp / rowmax at
// + g * out_stride (no ce output: the owner of the queries combines the ranges' sums).
extern "C" int tcar_ce_finish_groups(const float* rowsum_part, const float* rowmax_part, long long part_stride,
                                     float* sumexp, float* rowmax, long long out_stride, int n_tiles, const int* n_rows,
                                     int groups, int pass, void* stream) {
    if (!rowsum_part || !sumexp || !rowmax || !n_rows || groups < 1 || groups > TCAR_MAX_PEERS || part_stride < 1 ||
        out_stride < 1 || (pass != 1 && pass != 2) || (pass == 1 && !rowmax_part))
        return TCAR_ERR_ARG;
    GroupRows gr = {};
    int bmax = 0;
    for (int g = 0; g < groups; ++g) {
        if (n_rows[g] > TCAR_QROWS) return TCAR_ERR_ARG;
        gr.n[g] = n_rows[g] > 0 ? n_rows[g] : 0;
        if (gr.n[g] > bmax) bmax = gr.n[g];
    }
    if (bmax == 0) return 0;
    gr.part_stride = part_stride;
    gr.out_stride = out_stride;
    launch_pdl(ce_finish_kernel, dim3((bmax + 7) / 8, groups), dim3(1024), 0, STREAM, rowsum_part, rowmax_part,
               sumexp, static_cast<float*>(nullptr), rowmax, n_tiles, bmax, 1, pass, gr);
    return LAUNCH_RC();
}

extern "C" int tcar_rowsum_finish(const float* rowsum_part, float* out, int out_stride, int n_tiles, int B,
                                  void* stream) {
    if (B < 1 || B > TCAR_QROWS || out_stride < 1 || !out) return TCAR_ERR_ARG;
    launch_pdl(ce_finish_kernel, dim3((B + 7) / 8), dim3(1024), 0, STREAM, rowsum_part,
               static_cast<const float*>(nullptr), out, static_cast<float*>(nullptr), static_cast<float*>(nullptr),
               n_tiles, B, out_stride, 0, GroupRows{});
    return LAUNCH_RC();
}

extern "C" int tcar_col_jobs(const tcar_col_job* jobs, int njobs, void* stream) {
    if (!jobs || njobs < 1 || njobs > TCAR_COL_JOBS) return TCAR_ERR_ARG;
    ColJobs js = {};
    int maxcols = 0, maxsplit = 1;
    for (int i = 0; i < njobs; ++i) {
        const tcar_col_job& q = jobs[i];
        if (q.rows < 1 || q.cols < 1 || q.ld < q.cols || q.mode < 0 || q.mode > 2 || !q.a || !q.y || !q.out ||
            (q.mode != 2 && !q.dz))
            return TCAR_ERR_ARG;
        js.j[i] = q;
        int rs = q.scratch ? q.rows / 512 : 1;
        if (rs < 1) rs = 1;
        if (rs > TCAR_COL_MAX_SPLIT) rs = TCAR_COL_MAX_SPLIT;
        js.rsplit[i] = rs;
        if (q.cols > maxcols) maxcols = q.cols;
        if (rs > maxsplit) maxsplit = rs;
    }
    launch_pdl(col_jobs_kernel, dim3(dim3((maxcols + 31) / 32, njobs, maxsplit)), dim3(1024), 0, STREAM, js);
    return LAUNCH_RC();
}

extern "C" int tcar_act_bwd_colsum(const float* dy, const float* y, float* dz, float* gb, int rows, int cols, int ld,
                                   int mode, void* stream) {
    if (mode != 0 && mode != 1) return TCAR_ERR_ARG;
    tcar_col_job q = {dy, y, dz, gb, nullptr, rows, cols, ld, mode};
    return tcar_col_jobs(&q, 1, stream);
}

extern "C" int tcar_neg_loss(const float* a_ic, const float* item, const float* content, const int32_t* neg,
                             const float* ce, float* negloss, float* loss, float* coef, float* dA_neg, int B, int Nn,
                             void* stream) {
    if (B < 1 || Nn < 0) return TCAR_ERR_ARG;
    launch_pdl(neg_loss_kernel, dim3(B), dim3(256), 0, STREAM, a_ic, item, content, neg, ce, negloss, loss, coef, dA_neg, B, Nn);
    return LAUNCH_RC();
}

extern "C" int tcar_ce_from_sums(const float* sums, int stride, const float* rowmax, float* sumexp, float* ce, int B,
                                 void* stream) {
    if (B < 1 || stride < 1 || !sums || !sumexp || !ce) return TCAR_ERR_ARG;
    launch_pdl(ce_from_sums_kernel, dim3((B + 255) / 256), dim3(256), 0, STREAM, sums, stride, rowmax, sumexp, ce, B);
    return LAUNCH_RC();
}

extern "C" int tcar_loss_combine(const float* ce, const float* negloss, float* loss, int B, void* stream) {
    if (B < 1 || !ce || !negloss || !loss) return TCAR_ERR_ARG;
    launch_pdl(loss_combine_kernel, dim3((B + 255) / 256), dim3(256), 0, STREAM, ce, negloss, loss, B);
    return LAUNCH_RC();
}

extern "C" int tcar_score_bwd_finish(const float* dq_raw, const float* sumexp, const float* dA_neg,
                                     const float* a_ic, const float* ct_tab, const float* item, const float* content,
                                     const int32_t* mwdhm, const int32_t* label, float* d_a_ic, float* d_a_pt,
                                     float* dTq, void* qs_bf16, int B, void* stream) {
    if (B < 1 || B > TCAR_QROWS) return TCAR_ERR_ARG;
    launch_pdl(score_bwd_finish_kernel, dim3(TCAR_QROWS), dim3(256), 0, STREAM, dq_raw, sumexp, dA_neg, a_ic, ct_tab, item, content,
                                                            mwdhm, label, d_a_ic, d_a_pt, dTq,
                                                            static_cast<__nv_bfloat16*>(qs_bf16), B);
    return LAUNCH_RC();
}

extern "C" int tcar_small_table_grads(const int32_t* idx, const int32_t* ctx, const float* dXi, const float* dP,
                                      const float* dD, const float* dCT, const float* dTq, const float* a_pt,
                                      const float* pos, const float* month, const float* day, const float* week,
                                      const float* hour, const float* minute, const float* dur, float* g_pos,
                                      float* g_month, float* g_day, float* g_week, float* g_hour, float* g_minute,
                                      float* g_dur, float* part, int B, int T, void* stream) {
    (void)part;      // scratch of the former two-pass version; kept in the signature, no longer touched
    if (B < 1 || T < 1 || T > TCAR_MAXT) return TCAR_ERR_ARG;
    launch_pdl(table_grads_kernel, dim3(TCAR_MAXT + NB + 11), dim3(TGD_THREADS), 0, STREAM, 
        idx, ctx, dXi, dP, dD, dCT, dTq, a_pt, pos, month, day, week, hour, minute, dur, g_pos, g_month, g_day, g_week,
        g_hour, g_minute, g_dur, B, T);
    return LAUNCH_RC();
}

extern "C" int tcar_scatter_add_rows_range(const int32_t* seq, const int32_t* label, const int32_t* neg,
                                           const float* dXi, const float* a_ic, const float* coef, const float* item,
                                           float* g_item, int32_t* hash_keys, int32_t* hash_cnt, long long* hash_acc,
                                           int32_t* entry_slot, float* slot_sq, int hash_size, int B, int T, int Nn,
                                           int row_lo, int row_hi, void* stream) {
    const int entries = B * T + B + B * Nn;
    if (B < 1 || T < 1 || Nn < 0 || hash_size < 2 * entries || (hash_size & (hash_size - 1))) return TCAR_ERR_ARG;
    launch_pdl(scatter_count_kernel, dim3((entries + 255) / 256), dim3(256), 0, STREAM, seq, label, neg, hash_keys, hash_cnt, entry_slot,
                                                                    hash_size - 1, B, T, Nn, row_lo, row_hi);
    int rc = LAUNCH_RC();
    if (rc) return rc;
    launch_pdl(scatter_accum_kernel, dim3((entries + 7) / 8), dim3(256), 0, STREAM, 
        seq, label, neg, dXi, a_ic, coef, item, g_item, hash_cnt, entry_slot,
        reinterpret_cast<unsigned long long*>(hash_acc), slot_sq, B, T, Nn);
    rc = LAUNCH_RC();
    if (rc) return rc;
    launch_pdl(scatter_apply_kernel, dim3((hash_size + 255) / 256), dim3(256), 0, STREAM, hash_keys, hash_cnt, hash_acc, g_item, slot_sq,
                                                                  hash_size);
    return LAUNCH_RC();
}

extern "C" int tcar_scatter_add_rows_multi(const int32_t* ids, long long ids_stride, const float* payload,
                                           long long payload_stride, const float* item, float* g_item,
                                           int32_t* hash_keys, int32_t* hash_cnt, long long* hash_acc,
                                           int32_t* entry_slot, float* slot_sq, int hash_size, const int* n_rows,
                                           int groups, int T, int Nn, int row_lo, int row_hi, void* stream) {
    if (!n_rows || groups < 1 || groups > TCAR_MAX_PEERS || !ids || !payload || T < 1 || Nn < 0 ||
        (hash_size & (hash_size - 1)))
        return TCAR_ERR_ARG;
    ScatterGroups sg = {};
    sg.ids = ids;
    sg.payload = payload;
    sg.ids_stride = ids_stride;
    sg.pay_stride = payload_stride;
    long long total = 0;
    int emax = 0;
    for (int g = 0; g < groups; ++g) {
        const int B = n_rows[g] > 0 ? n_rows[g] : 0;
        if (B > TCAR_QROWS) return TCAR_ERR_ARG;
        sg.n[g] = B;
        const int e = B * T + B + B * Nn;
        total += e;
        if (e > emax) emax = e;
    }
    sg.es_stride = emax;
    // every entry may fall into [row_lo, row_hi): the table must hold them all at load factor <= 1/2, entry_slot needs
    // groups * emax words
    if (total == 0 || (long long)hash_size < 2 * total) return TCAR_ERR_ARG;
    launch_pdl(scatter_count_groups_kernel, dim3((emax + 255) / 256, groups), dim3(256), 0, STREAM, sg, hash_keys,
               hash_cnt, entry_slot, hash_size - 1, T, Nn, row_lo, row_hi);
    int rc = LAUNCH_RC();
    if (rc) return rc;
    launch_pdl(scatter_accum_groups_kernel, dim3((emax + 7) / 8, groups), dim3(256), 0, STREAM, sg, item, g_item,
               static_cast<const int32_t*>(hash_cnt), static_cast<const int32_t*>(entry_slot),
               reinterpret_cast<unsigned long long*>(hash_acc), slot_sq, T, Nn);
    rc = LAUNCH_RC();
    if (rc) return rc;
    launch_pdl(scatter_apply_kernel, dim3((hash_size + 255) / 256), dim3(256), 0, STREAM, hash_keys, hash_cnt, hash_acc,
               g_item, slot_sq, hash_size);
    return LAUNCH_RC();
}

extern "C" int tcar_scatter_add_rows(const int32_t* seq, const int32_t* label, const int32_t* neg, const float* dXi,
                                     const float* a_ic, const float* coef, const float* item, float* g_item,
                                     int32_t* hash_keys, int32_t* hash_cnt, long long* hash_acc, int32_t* entry_slot,
                                     float* slot_sq, int hash_size, int B, int T, int Nn, void* stream) {
    return tcar_scatter_add_rows_range(seq, label, neg, dXi, a_ic, coef, item, g_item, hash_keys, hash_cnt, hash_acc,
                                       entry_slot, slot_sq, hash_size, B, T, Nn, 0, 0x7fffffff, stream);
}
