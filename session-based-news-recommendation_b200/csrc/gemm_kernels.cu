// Session-side dense projections on 5th-gen tensor cores (tcgen05.mma.kind::tf32 + TMEM + TMA), sm_100a only.
//
// Replaces the TF matmuls behind linear_2d / linear_3d (modules.py:43-70) and the weight / data gradients TF derives
// from them (model_combine.py:156):
//
//   C[M,N] = act( sum_seg A_seg[M,K_seg] . B_seg[K_seg,N] + bias )         (or  C += ...)
//
// * `precise` (forward): 3xTF32 -- every fp32 operand is split into hi = top 19 bits and lo = a - hi, and each K step
//   issues  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  into the same fp32 TMEM accumulator (error ~2^-21, i.e. fp32-class;
//   the exact re-scoring of the evaluation top-20 depends on fp32-accurate session vectors).  The B operand (weights)
//   is pre-split by prep_weights_kernel once per step; the A operand (activations) is split in shared memory by the
//   four epilogue warps while the tensor core works on the previous stage.
// * fast (backward): single-pass TF32 on operands as they are (gradients carry a 2e-2 tolerance).
// * operands may be K-major or MN-major (UMMA descriptors take both), so X.W, X^T.dU and dU.W^T all map onto the same
//   kernel without transposes; the reduction dimension can be split across CTAs (fixed-order second pass).
#include "sm100_ptx.cuh"
#include "tcar_b200.h"
#include "launch.cuh"

namespace tcar {

#ifndef TCAR_GEMM_EPI_WARPS
#define TCAR_GEMM_EPI_WARPS 8
#endif
constexpr int G_EPI_WARPS = TCAR_GEMM_EPI_WARPS;  // a multiple of 4: G_EPI_WARPS / 4 warps per TMEM lane quarter
constexpr int G_THREADS = 128 + 32 * G_EPI_WARPS;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, then split + epilogue warps
constexpr int G_BM = 128;
constexpr int G_BK = 32;         // 32 fp32 = one 128-byte swizzle span
constexpr int G_A_TILE = G_BM * G_BK * 4;   // 16384
constexpr int G_MAX_SEG = 3;
constexpr int G_SMEM_BUDGET = 196608;       // stage memory (192 KB)
constexpr int G_SMEM = G_SMEM_BUDGET + 1024 + 256;

struct GemmSegDev {
    int nkb;          // K blocks of this segment
    int a_mn;         // 1: A operand is MN-major (tensor [K, M]), 0: K-major (tensor [M, K])
    int b_mn;         // 1: B operand is MN-major (tensor [K, N]), 0: K-major (tensor [N, K])
    int a_koff;       // first K index of the segment inside the A tensor (TMA coordinate offset)
};

struct GemmParams {
    GemmSegDev seg[G_MAX_SEG];
    int nseg;
    int M, N, bn;          // bn = UMMA N (64 | 128 | 256)
    int precise;           // 1: 3xTF32
    int stages, stage_bytes, b_tile_bytes;
    int splits, kb_per_split, kb_total;
    int mtiles, ntiles;
    const float* bias;     // [N] or null
    int act;               // 0 none, 1 relu, 2 tanh
    float* C;              // [M, ldc]   (or partials [splits][M_pad][ldc] when splits > 1)
    int ldc;
    int accumulate;        // C += result
    long long part_stride; // elements between split partials
    float* out;            // final destination of the split reduction
    int ld_out;
    float* out2;           // optional second destination for rows >= out2_row0 (same pitch as the first)
    int out2_row0;
};

struct GemmMaps {
    CUtensorMap a[G_MAX_SEG];
    CUtensorMap b[G_MAX_SEG];
    CUtensorMap blo[G_MAX_SEG];
};

// One launch = up to G_MAX_PROB independent problems (e.g. every weight gradient that consumes dU1 / dU2); CTAs are
// numbered problem after problem: [cta_start[g], cta_start[g+1]).
constexpr int G_MAX_PROB = TCAR_GEMM_MAX_GROUP;
struct GemmGroup {
    GemmMaps maps[G_MAX_PROB];
    GemmParams prm[G_MAX_PROB];
    int cta_start[G_MAX_PROB + 1];
    int red_start[G_MAX_PROB + 1];   // CTA ranges of the split-reduction kernel
    int nprob;
    long long* trace;                // debug (tcar_debug_gemm_trace): 8 clock64() stamps per CTA, else null
};
#define G_TRACE(slot) do { if (grp.trace) grp.trace[(size_t)blockIdx.x * 8 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32 instruction descriptor: D fp32, A/B tf32 (format 2), majors, N >> 3, M >> 4
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// MN-major tf32 operands exist only in the "128B swizzle with 32-byte atomicity" layout (descriptor layout type 1,
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of [32 mn (128 B, contiguous) x 4 k]; one UMMA (K = 8) spans two
// atoms along K (SBO = 512 B apart inside a [32 mn x 32 k] TMA box); MN atoms are one box (4096 B) apart (LBO).
__device__ __forceinline__ uint64_t sdesc_mnmajor_tf32(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((4096u >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((512u >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(1) << 61;
    return d;
}
// fp32 -> tf32 with round-to-nearest (low 13 bits zero afterwards, so the tensor core's own truncation is a no-op)
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// tanh(x) = 1 - 2 / (exp(2x) + 1) with ex2.approx / rcp.approx: absolute error < 3e-7 (the epilogue thread owns a
// whole output row, so the ~30-instruction libm tanhf would dominate the small projections)
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(x * 2.8853900817779268f);     // exp(2x)
    return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ GemmGroup grp) {
    PDL_ENTER();
    int gi = 0;
    while (gi + 1 < grp.nprob && (int)blockIdx.x >= grp.cta_start[gi + 1]) ++gi;
    const GemmMaps& maps = grp.maps[gi];
    const GemmParams& p = grp.prm[gi];
    if (threadIdx.x == 0) G_TRACE(0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_SMEM_BUDGET);
    uint64_t* full = bars;            // [8]  TMA bytes landed
    uint64_t* conv = bars + 8;        // [8]  A operand split done (precise mode)
    uint64_t* empty = bars + 16;      // [8]  MMAs of the stage retired
    uint64_t* acc_full = bars + 24;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const int local = (int)blockIdx.x - grp.cta_start[gi];
    const int mtile = local % p.mtiles, ntile = (local / p.mtiles) % p.ntiles, split = local / (p.mtiles * p.ntiles);
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nkb = max(kb1 - kb0, 0);
    // stage layout: [A hi 16K][A lo 16K (precise)][B hi b_tile][B lo b_tile (precise)]
    const int a_lo_off = G_A_TILE;
    const int b_off = p.precise ? 2 * G_A_TILE : G_A_TILE;
    const int b_lo_off = b_off + p.b_tile_bytes;

    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.nseg; ++s) {
            tma_prefetch_desc(&maps.a[s]);
            tma_prefetch_desc(&maps.b[s]);
            if (p.precise) tma_prefetch_desc(&maps.blo[s]);
        }
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&conv[i], G_EPI_WARPS);
            mbar_init(&empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) G_TRACE(1);

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            int kb = 0;  // global K-block index over all segments
            for (int s = 0; s < p.nseg; ++s) {
                const GemmSegDev sg = p.seg[s];
                for (int j = 0; j < sg.nkb; ++j, ++kb) {
                    if (kb < kb0 || kb >= kb1) continue;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + (size_t)stage * p.stage_bytes;
                    const uint32_t bytes = G_A_TILE + p.b_tile_bytes * (p.precise ? 2 : 1);
                    mbar_expect_tx(&full[stage], bytes);
                    const int k0 = j * G_BK;
                    const int ka = sg.a_koff + k0;
                    if (sg.a_mn) {
                        // tensor [K, M]: boxes of [32 m x 32 k]
#pragma unroll
                        for (int i = 0; i < G_BM / 32; ++i)
                            tma_load_2d(st + i * 4096, &maps.a[s], &full[stage], mtile * G_BM + i * 32, ka);
                    } else {
                        tma_load_2d(st, &maps.a[s], &full[stage], ka, mtile * G_BM);
                    }
                    if (sg.b_mn) {
                        for (int i = 0; i < p.bn / 32; ++i) {
                            tma_load_2d(st + b_off + i * 4096, &maps.b[s], &full[stage], ntile * p.bn + i * 32, k0);
                            if (p.precise)
                                tma_load_2d(st + b_lo_off + i * 4096, &maps.blo[s], &full[stage],
                                            ntile * p.bn + i * 32, k0);
                        }
                    } else {
                        tma_load_2d(st + b_off, &maps.b[s], &full[stage], k0, ntile * p.bn);
                        if (p.precise) tma_load_2d(st + b_lo_off, &maps.blo[s], &full[stage], k0, ntile * p.bn);
                    }
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
            }
            G_TRACE(2);
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0;
        int kb = 0, issued = 0;
        for (int s = 0; s < p.nseg; ++s) {
            const GemmSegDev sg = p.seg[s];
            const uint32_t idesc = make_idesc_tf32(G_BM, p.bn, sg.a_mn, sg.b_mn);
            for (int j = 0; j < sg.nkb; ++j, ++kb) {
                if (kb < kb0 || kb >= kb1) continue;
                mbar_wait(p.precise ? &conv[stage] : &full[stage], phase);
                tc_fence_after();
                if (issued == 0 && lane == 0) G_TRACE(3);
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
                    const uint32_t b_addr = a_addr + b_off;
#pragma unroll
                    for (int k = 0; k < G_BK / 8; ++k) {
                        // K-major: 8 tf32 = 32 bytes along the swizzle span; MN-major: 8 k-rows of 128 bytes
                        const uint32_t ao = sg.a_mn ? k * 1024 : k * 32;
                        const uint32_t bo = sg.b_mn ? k * 1024 : k * 32;
                        const uint64_t ad = sg.a_mn ? sdesc_mnmajor_tf32(a_addr + ao) : sdesc_kmajor(a_addr + ao);
                        const uint64_t bd = sg.b_mn ? sdesc_mnmajor_tf32(b_addr + bo) : sdesc_kmajor(b_addr + bo);
                        umma_tf32(tmem_base, ad, bd, idesc, (issued | k) != 0);
                        if (p.precise) {
                            const uint64_t adl = sg.a_mn ? sdesc_mnmajor_tf32(a_addr + a_lo_off + ao)
                                                         : sdesc_kmajor(a_addr + a_lo_off + ao);
                            const uint64_t bdl = sg.b_mn ? sdesc_mnmajor_tf32(a_addr + b_lo_off + bo)
                                                         : sdesc_kmajor(a_addr + b_lo_off + bo);
                            umma_tf32(tmem_base, adl, bd, idesc, 1);
                            umma_tf32(tmem_base, ad, bdl, idesc, 1);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) { umma_commit(acc_full); G_TRACE(4); }
                }
                __syncwarp();
                ++issued;
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const uint32_t q = warp & 3;              // TMEM lane quarter this warp may read (warp id mod 4)
        const int half = (int)(warp - 4) >> 2;    // which of the G_EPI_WARPS / 4 warps of the quarter
        const uint32_t tid = threadIdx.x - 128;   // 0 .. 32 * G_EPI_WARPS - 1
        if (p.precise) {
            // ================= operand split: A tile -> (hi in place, lo next to it) =================
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < nkb; ++it) {
                mbar_wait(&full[stage], phase);
                float4* hi = reinterpret_cast<float4*>(smem + (size_t)stage * p.stage_bytes);
                float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * p.stage_bytes + a_lo_off);
#pragma unroll
                for (int i = 0; i < G_A_TILE / 16 / (32 * G_EPI_WARPS); ++i) {
                    const float4 v = hi[tid + i * 32 * G_EPI_WARPS];
                    float4 h, l;
                    h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
                    h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
                    h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
                    h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
                    hi[tid + i * 32 * G_EPI_WARPS] = h;
                    lo[tid + i * 32 * G_EPI_WARPS] = l;
                }
                fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[stage]);
                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
            }
        }
        // ================= epilogue: thread <-> output row, the two warps of a TMEM lane quarter alternate over the
        // 32-column chunks.  (Measured with tcar_debug_gemm_trace: the epilogue of the small projections is bound by
        // the issue latency of ONE warp per scheduler walking its chunks -- about 1 us per chunk -- not by the store
        // pattern; a shared-memory transpose for coalesced stores was slower.  Hence: twice the warps, one bias load
        // per lane instead of 32 per thread.)
        const int row = mtile * G_BM + q * 32 + lane;
        if (nkb > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (tid == 0) G_TRACE(5);
        float* crow = p.C + (size_t)split * p.part_stride + (size_t)row * p.ldc;
        // second destination (unsplit problems only; split ones apply it in the reduction kernel)
        float* crow2 = (p.out2 && p.splits == 1 && row >= p.out2_row0) ? p.out2 + (size_t)(row - p.out2_row0) * p.ldc
                                                                       : nullptr;
        const bool vec = (p.ldc & 3) == 0;
        const float* bias = p.bias;
        const int act = p.act, accumulate = p.accumulate, Ncols = p.N;
#pragma unroll 1
        for (int ch = half; ch < p.bn / 32; ch += G_EPI_WARPS / 4) {
            const int c0 = ntile * p.bn + ch * 32;
            if (c0 >= Ncols) break;
            uint32_t v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((q * 32) << 16) + ch * 32, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            const float bl = (bias && c0 + (int)lane < Ncols) ? __ldg(bias + c0 + lane) : 0.f;
            if (nkb > 0) tmem_ld_wait();
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bl, j);
                if (act == 1) x = fmaxf(x, 0.f);
                else if (act == 2) x = tanh_fast(x);
                o[j] = x;
            }
            if (row < p.M) {
                if (vec && c0 + 32 <= Ncols) {
                    float4* dst = reinterpret_cast<float4*>(crow + c0);
                    if (accumulate) {
                        float4 old[8];
#pragma unroll
                        for (int g = 0; g < 8; ++g) old[g] = dst[g];
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            dst[g] = make_float4(o[g * 4] + old[g].x, o[g * 4 + 1] + old[g].y, o[g * 4 + 2] + old[g].z,
                                                 o[g * 4 + 3] + old[g].w);
                    } else {
#pragma unroll
                        for (int g = 0; g < 8; ++g) dst[g] = make_float4(o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < Ncols) crow[c0 + j] = accumulate ? crow[c0 + j] + o[j] : o[j];
                }
                if (crow2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < Ncols) crow2[c0 + j] = o[j];
                }
            }
        }
    }
    if (threadIdx.x == 128) G_TRACE(6);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) G_TRACE(7);
}

// out[r, c] = sum_s part[s][r][c]  (fixed order) for every split problem of the group
__global__ void __launch_bounds__(256)
gemm_reduce_splits_kernel(const __grid_constant__ GemmGroup grp) {
    PDL_ENTER();
    int gi = 0;
    while (gi + 1 < grp.nprob && (int)blockIdx.x >= grp.red_start[gi + 1]) ++gi;
    const GemmParams& p = grp.prm[gi];
    const int i = ((int)blockIdx.x - grp.red_start[gi]) * blockDim.x + threadIdx.x;
    if (p.splits <= 1 || i >= p.M * p.N) return;
    const int r = i / p.N, c = i % p.N;
    float acc = 0.f;
    for (int s = 0; s < p.splits; ++s) acc += p.C[(size_t)s * p.part_stride + (size_t)r * p.ldc + c];
    p.out[(size_t)r * p.ld_out + c] = acc;
    if (p.out2 && r >= p.out2_row0) p.out2[(size_t)(r - p.out2_row0) * p.ld_out + c] = acc;
}

// Pre-split of the dense weights (once per step, after Adam): for every tensor t of the table
//   hi[dst_off + r*dst_pitch + c] = tf32(w) (round to nearest),   lo[...] = tf32(w - hi)   (pad columns are zero).
// table rows: {src_off, rows, cols, dst_off, dst_pitch, src2_off, src2_row0}: a second tensor (same column count) is
// added to rows >= src2_row0 when src2_off >= 0.
__global__ void __launch_bounds__(256)
prep_weights_kernel(const float* __restrict__ theta, const int32_t* __restrict__ table, float* __restrict__ hi,
                    float* __restrict__ lo) {
    PDL_ENTER();
    const int32_t* t = table + blockIdx.y * 7;
    const int src = t[0], rows = t[1], cols = t[2], dst = t[3], pitch = t[4], src2 = t[5], row0 = t[6];
    const int n = rows * pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / pitch, c = i % pitch;
        float w = c < cols ? theta[src + r * cols + c] : 0.f;
        if (src2 >= 0 && r >= row0 && c < cols) w += theta[src2 + (r - row0) * cols + c];
        const float h = tf32_rn(w);
        hi[dst + i] = h;
        lo[dst + i] = tf32_rn(w - h);
    }
}

typedef CUresult (*EncodeTiledFn32)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn32 get_encode_fn32() {
    static EncodeTiledFn32 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn32>(ptr);
    }
    return fn;
}

// 2-D fp32 row-major tensor [rows, cols] (cols contiguous, pitch in elements); box = [box_cols, box_rows], SW128,
// out-of-bounds elements read as zero (this is what pads K and the M / N edges).
static int make_map_f32(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                        uint32_t box_cols, uint32_t box_rows, bool mn_major = false) {
    EncodeTiledFn32 fn = get_encode_fn32();
    if (!fn) return TCAR_ERR_DRIVER;
    if ((pitch_elems * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15)) return TCAR_ERR_ARG;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = tmap_encode_cached(fn, m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : TCAR_ERR_TENSORMAP;
}

}  // namespace tcar

using namespace tcar;

extern "C" int tcar_gemm_tf32_splits(int M, int N, int k_total, int want) {
    (void)M; (void)N;
    const int kb = (k_total + G_BK - 1) / G_BK;
    int s = want < 1 ? 1 : want;
    if (s > kb) s = kb;
    return s;
}

static int setup_problem(const tcar_gemm_problem& q, GemmParams& p, GemmMaps& maps, int group_ctas_hint) {
    const int M = q.M, N = q.N, nseg = q.nseg, precise = q.precise;
    int splits = q.splits;
    if (nseg < 1 || nseg > G_MAX_SEG || M < 1 || N < 1 || !q.C) return TCAR_ERR_ARG;
    if (splits < 1 || (splits > 1 && (!q.part || q.bias || q.act || precise || q.accumulate))) return TCAR_ERR_ARG;
    p = GemmParams{};
    p.nseg = nseg;
    p.M = M;
    p.N = N;
    p.precise = precise ? 1 : 0;
    // UMMA N: one tile when N <= 256, else 256-wide tiles; narrow outputs use the smallest legal tile.  Small
    // problems are bound by the per-SM operand ingest (~64 B/clk), not by the tensor pipe: narrower tiles spread
    // them over more SMs.
    p.bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    const int mt0 = (M + G_BM - 1) / G_BM;
    // 3xTF32 streams hi+lo of both operands: 128-wide tiles keep three 64 KB stages in flight instead of two 96 KB ones
    if (precise && p.bn > 128) p.bn = 128;
    // (split problems get their parallelism from the split factor: narrowing them would only re-read A more often)
    while (splits == 1 && p.bn > 64 && group_ctas_hint + mt0 * ((N + p.bn - 1) / p.bn) < 74) p.bn >>= 1;
    p.b_tile_bytes = p.bn * G_BK * 4;
    p.stage_bytes = (G_A_TILE + p.b_tile_bytes) * (precise ? 2 : 1);
    p.stages = G_SMEM_BUDGET / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return TCAR_ERR_ARG;
    int kb_total = 0;
    for (int s = 0; s < nseg; ++s) {
        const tcar_gemm_seg& g = q.segs[s];
        if (!g.a || !g.b || g.k < 1 || (precise && !g.b_lo) || g.a_koff < 0) return TCAR_ERR_ARG;
        p.seg[s].nkb = (g.k + G_BK - 1) / G_BK;
        p.seg[s].a_mn = g.a_mn_major ? 1 : 0;
        p.seg[s].b_mn = g.b_mn_major ? 1 : 0;
        p.seg[s].a_koff = g.a_koff;
        kb_total += p.seg[s].nkb;
        int rc;
        // the tensor extent along K ends at a_koff + k, so the last K block is zero-filled beyond it
        if (g.a_mn_major) rc = make_map_f32(&maps.a[s], g.a, g.a_koff + g.k, M, g.lda, 32, G_BK, true);  // [K, M]
        else rc = make_map_f32(&maps.a[s], g.a, M, g.a_koff + g.k, g.lda, G_BK, G_BM);             // [M, K]
        if (rc) return rc;
        if (g.b_mn_major) rc = make_map_f32(&maps.b[s], g.b, g.k, N, g.ldb, 32, G_BK, true);       // tensor [K, N]
        else rc = make_map_f32(&maps.b[s], g.b, N, g.k, g.ldb, G_BK, p.bn);                 // tensor [N, K]
        if (rc) return rc;
        if (precise) {
            if (g.b_mn_major) rc = make_map_f32(&maps.blo[s], g.b_lo, g.k, N, g.ldb, 32, G_BK, true);
            else rc = make_map_f32(&maps.blo[s], g.b_lo, N, g.k, g.ldb, G_BK, p.bn);
            if (rc) return rc;
        } else {
            maps.blo[s] = maps.b[s];
        }
    }
    for (int s = nseg; s < G_MAX_SEG; ++s) {
        maps.a[s] = maps.a[0];
        maps.b[s] = maps.b[0];
        maps.blo[s] = maps.blo[0];
    }
    p.kb_total = kb_total;
    if (splits > kb_total) splits = kb_total;
    p.splits = splits;
    p.kb_per_split = (kb_total + splits - 1) / splits;
    p.bias = q.bias;
    p.act = q.act;
    p.accumulate = q.accumulate ? 1 : 0;
    p.mtiles = mt0;
    p.ntiles = (N + p.bn - 1) / p.bn;
    p.out = q.C;
    p.ld_out = q.ldc;
    p.out2 = q.C2;
    p.out2_row0 = q.c2_row0;
    if (q.C2 && (q.accumulate || q.c2_row0 < 0 || q.c2_row0 >= M)) return TCAR_ERR_ARG;
    if (splits > 1) {
        p.C = q.part;
        p.ldc = p.ntiles * p.bn;
        p.part_stride = (long long)p.mtiles * G_BM * p.ldc;
    } else {
        p.C = q.C;
        p.ldc = q.ldc;
        p.part_stride = 0;
    }
    return 0;
}

static long long* g_gemm_trace = nullptr;
extern "C" int tcar_debug_gemm_trace(long long* trace_buf) {
    g_gemm_trace = trace_buf;
    return 0;
}

extern "C" int tcar_gemm_tf32_group(const tcar_gemm_problem* probs, int nprob, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (nprob < 1 || nprob > G_MAX_PROB || !probs) return TCAR_ERR_ARG;
    static thread_local GemmGroup grp;      // ~11 KB of kernel parameters, rebuilt per call
    grp.nprob = nprob;
    grp.trace = g_gemm_trace;
    int ctas = 0, red = 0;
    bool any_split = false;
    for (int g = 0; g < nprob; ++g) {
        int rc = setup_problem(probs[g], grp.prm[g], grp.maps[g], ctas);
        if (rc) return rc;
        grp.cta_start[g] = ctas;
        grp.red_start[g] = red;
        ctas += grp.prm[g].mtiles * grp.prm[g].ntiles * grp.prm[g].splits;
        if (grp.prm[g].splits > 1) {
            red += (probs[g].M * probs[g].N + 255) / 256;
            any_split = true;
        }
    }
    for (int g = nprob; g <= G_MAX_PROB; ++g) {
        grp.cta_start[g] = ctas;
        grp.red_start[g] = red;
    }
    TCAR_SET_SMEM_ONCE(gemm_tf32_kernel, G_SMEM);
    launch_pdl(gemm_tf32_kernel, dim3(ctas), dim3(G_THREADS), G_SMEM, stream, grp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (any_split) {
        launch_pdl(gemm_reduce_splits_kernel, dim3(red), dim3(256), 0, stream, grp);
        e = cudaGetLastError();
    }
    return (int)e;
}

extern "C" int tcar_gemm_tf32(const tcar_gemm_seg* segs, int nseg, int M, int N, const float* bias, int act, float* C,
                              int ldc, int accumulate, int precise, int splits, float* part, void* stream_) {
    if (nseg < 1 || nseg > G_MAX_SEG || !segs) return TCAR_ERR_ARG;
    tcar_gemm_problem q = {};
    for (int s = 0; s < nseg; ++s) q.segs[s] = segs[s];
    q.nseg = nseg;
    q.M = M;
    q.N = N;
    q.bias = bias;
    q.act = act;
    q.C = C;
    q.ldc = ldc;
    q.accumulate = accumulate;
    q.precise = precise;
    q.splits = splits;
    q.part = part;
    return tcar_gemm_tf32_group(&q, 1, stream_);
}

extern "C" long long tcar_gemm_tf32_part_elems(int M, int N, int splits) {
    // upper bound over every tile width the launcher may pick (N rounded up to 256 columns)
    const long long mt = (M + G_BM - 1) / G_BM, nt = (N + 255) / 256;
    return (long long)splits * mt * G_BM * nt * 256;
}

extern "C" int tcar_prep_weights(const float* theta, const int32_t* table, int ntensors, float* hi, float* lo,
                                 void* stream_) {
    if (ntensors < 1) return TCAR_ERR_ARG;
    launch_pdl(prep_weights_kernel, dim3(dim3(74, ntensors)), dim3(256), 0, static_cast<cudaStream_t>(stream_), theta, table, hi, lo);
    return (int)cudaGetLastError();
}
