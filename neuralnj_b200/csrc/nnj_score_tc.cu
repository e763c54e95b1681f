// nnj_score_tc.cu — pair scoring on tcgen05 (precision bf16x3): the fused decode_gg kernel of the NJ loop.
//
// For every (pair, site) row the reference computes (model.py:90-99, 102-155)
//     x      = z*x_i + (1-z)*x_j                       (k_blend_planes, stored as bf16 hi/lo planes)
//     x_glob = sum_r alpha[pair,r] * X[r,site,:]   \
//     g      = W_g x_glob + b_g = sum_r alpha[pair,r] * (W_g X[r,site,:]) + b_g   (sum_r alpha = 1)
//                                                       -> UMMA 1: [128 pairs x slots] . [slots x (64 | 64)]  (B MN-major, TMA)
//     w      = sigmoid(g);  x' = (1-w)*x + w*x_glob
//     score += w2 . GELU(W_s x' + b_s) + b2             -> UMMA 2: [128 x 64] . W_s^T
// G = W_g X is kept per node next to X in the node planes, which removes one GEMM and one operand round trip per row.
// One CTA = 128 pairs x 64 sites.  Sites alternate between two pipelines, each with its own TMEM accumulators
// ([x_glob | g]: 128 columns, s: 64), operand buffers and a group of 8 epilogue warps (4 TMEM lane quarters x 2 column
// halves); two control warps issue TMA + tcgen05.mma, so the tensor core works on one site while the CUDA cores run
// the sigmoid / blend / GELU epilogue of the other.  With <= 64 pairs (every step after the first) the pairs are
// duplicated into rows 64..127 so that all four lane quarters - hence all 16 epilogue warps - have work.  Every product is split-bf16 (hi*hi + hi*lo + lo*hi, fp32
// accumulate); operands produced on the fly are written by the epilogue threads straight into SWIZZLE_128B shared
// memory (validated by nnj_tc_selftest).  Node state for UMMA 1 comes from site-major bf16 planes [B][C][S][128]
// [X | W_g X], indexed by PHYSICAL slot, so alpha is scattered to slot order and dead / free slots simply get weight 0.
#include "nnj_internal.h"
#include "nnj_tc.cuh"

namespace nnj {

constexpr int ST_THREADS = 576;            // 16 epilogue warps + 2 control warps
constexpr int ST_SITES = 64;               // sites per CTA
constexpr int ST_A0 = 0;                   // alpha operand   hi 16 KB | lo 16 KB
constexpr int ST_W = 32768;                // Ws_h, Ws_l: 2 x 8 KB
constexpr int ST_PIPE = 49152;             // per pipeline: [X|G]_h 16K | [X|G]_l 16K | A1_h 16K | A1_l 16K = 64 KB
constexpr int ST_PIPE_BYTES = 65536;
constexpr int ST_MISC = ST_PIPE + 2 * ST_PIPE_BYTES;   // biases (768 B) | part (4 KB) | barriers | tmem slot
constexpr int ST_SMEM = ST_MISC + 768 + 4096 + 256 + 1024;

// x planes of the listed pairs: xh/xl [B][pc][C][64] bf16.  grid (ceil(C/16), nc, B), 256 threads.
__global__ void __launch_bounds__(256) k_blend_planes(const float* __restrict__ X, const float* __restrict__ Y, size_t tree_stride,
                                                      const int32_t* __restrict__ slot_of, int slot_stride, int C,
                                                      const int32_t* __restrict__ pair_i, const int32_t* __restrict__ pair_j,
                                                      int pair_stride, int n0, const float* __restrict__ bh, uint2* __restrict__ xh,
                                                      uint2* __restrict__ xl, int pc) {
    const int b = blockIdx.z, n = blockIdx.y;
    const int c = blockIdx.x * 16 + (threadIdx.x >> 4), d4 = (threadIdx.x & 15) * 4;
    if (c >= C) return;
    const int li = pair_i[(size_t)b * pair_stride + n0 + n], lj = pair_j[(size_t)b * pair_stride + n0 + n];
    uint2 oh = make_uint2(0u, 0u), ol = make_uint2(0u, 0u);
    if (li >= 0) {
        const int32_t* so = slot_of + (size_t)b * slot_stride;
        const size_t oi = (size_t)b * tree_stride + ((size_t)so[li] * C + c) * D + d4, oj = (size_t)b * tree_stride + ((size_t)so[lj] * C + c) * D + d4;
        const float4 xi = ld4(X + oi), xj = ld4(X + oj), yi = ld4(Y + oi), yj = ld4(Y + oj);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bh + d4));
        float z, v0, v1, v2, v3;
        z = sigmoid_fast(yi.x - yj.x + b4.x); v0 = fmaf(z, xi.x - xj.x, xj.x);   // z x_i + (1-z) x_j
        z = sigmoid_fast(yi.y - yj.y + b4.y); v1 = fmaf(z, xi.y - xj.y, xj.y);
        z = sigmoid_fast(yi.z - yj.z + b4.z); v2 = fmaf(z, xi.z - xj.z, xj.z);
        z = sigmoid_fast(yi.w - yj.w + b4.w); v3 = fmaf(z, xi.w - xj.w, xj.w);
        split2(v0, v1, oh.x, ol.x);
        split2(v2, v3, oh.y, ol.y);
    }
    const size_t o = (((size_t)b * pc + n) * C + c) * 16 + (threadIdx.x & 15);   // uint2 = 4 bf16
    xh[o] = oh;
    xl[o] = ol;
}

// ------------------------------------------------------------------ fused pair blend + alpha partials (model.py:105-118)
// One CTA = up to 128 pairs x 64 sites.  For every site the 8 blend warps form x = z x_i + (1-z) x_j (z = sigmoid(Y_i - Y_j + b_h))
// straight from the fp32 node pool, store it (bf16 hi/lo) both into the x planes the pair-score kernel reads later and into a
// double-buffered SWIZZLE_128B A operand; a control warp streams the site's K' tile [slots x 64] in by TMA (3-deep ring) and issues
//     alpha_part[pair, slot] += x[pair, site, :] . K'[slot, site, :]          (split-bf16 UMMA, fp32 accumulator in TMEM)
// so the tensor core works on site s while the CUDA cores blend site s+1.  The accumulator is written once per CTA as the
// partial of this site group; k_alpha_softmax reduces the groups in a fixed order.  With <= 64 pairs four threads share a row.
constexpr int AT_THREADS = 288;            // 8 blend warps + 1 control warp
constexpr int AT_SITES = 64;
constexpr int AT_A = 0;                    // 2 x (hi 16 KB | lo 16 KB)
constexpr int AT_K = 65536;                // 3 x (hi 8 KB | lo 8 KB)
constexpr int AT_MISC = AT_K + 3 * 16384;  // pair slots (1 KB) | b_h (256 B) | barriers | tmem slot
constexpr int AT_SMEM = 1024 + AT_MISC + 1024 + 256 + 128;

struct AlphaTcArgs {
    const float* X; const float* Y; size_t tree_stride;
    const int32_t* slot_of; int slot_stride;
    const int32_t* pair_i; const int32_t* pair_j; int pair_stride; int n0; int nc;
    int C, S;
    const float* bh;
    uint4* xh; uint4* xl; int pc;            // x planes [B][pc][C][64] bf16 (output)
    float* alpha_part; int alpha_pairs; int nSG; int RP;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// blend NCOL columns of one pair row at one site and publish them (A operand + x planes)
template <int NCOL>
__device__ __forceinline__ void blend_site(const AlphaTcArgs& a, const float* __restrict__ xi_p, const float* __restrict__ xj_p,
                                           const float* __restrict__ yi_p, const float* __restrict__ yj_p, const float* __restrict__ s_bh,
                                           uint8_t* a_hi, uint8_t* a_lo, int row, int col0, uint4* __restrict__ gxh, uint4* __restrict__ gxl) {
#pragma unroll
    for (int e = 0; e < NCOL / 8; ++e) {
        float v[8];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int o = 8 * e + 4 * h2;
            const float4 b4 = *reinterpret_cast<const float4*>(s_bh + col0 + o);
            const float4 xa = ld4(xi_p + o), xb = ld4(xj_p + o), ya = ld4(yi_p + o), yb = ld4(yj_p + o);
            v[4 * h2 + 0] = fmaf(sigmoid_fast(ya.x - yb.x + b4.x), xa.x - xb.x, xb.x);   // z x_i + (1-z) x_j
            v[4 * h2 + 1] = fmaf(sigmoid_fast(ya.y - yb.y + b4.y), xa.y - xb.y, xb.y);
            v[4 * h2 + 2] = fmaf(sigmoid_fast(ya.z - yb.z + b4.z), xa.z - xb.z, xb.z);
            v[4 * h2 + 3] = fmaf(sigmoid_fast(ya.w - yb.w + b4.w), xa.w - xb.w, xb.w);
        }
        uint4 hh, ll;
        split2(v[0], v[1], hh.x, ll.x);
        split2(v[2], v[3], hh.y, ll.y);
        split2(v[4], v[5], hh.z, ll.z);
        split2(v[6], v[7], hh.w, ll.w);
        const int off = row * 128 + ((((col0 >> 3) + e) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = hh;
        *reinterpret_cast<uint4*>(a_lo + off) = ll;
        gxh[e] = hh;
        gxl[e] = ll;
    }
}

__global__ void __launch_bounds__(AT_THREADS, 2)
k_alpha_tc(const __grid_constant__ CUtensorMap mapKh, const __grid_constant__ CUtensorMap mapKl, const AlphaTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_align1024(smem_raw);
    int* s_pi = reinterpret_cast<int*>(sm + AT_MISC);      // physical slots of the pair rows (-1: no pair)
    int* s_pj = s_pi + 128;
    float* s_bh = reinterpret_cast<float*>(s_pj + 128);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bh + 64);   // a_ready[2], mma_done[2], k_full[3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, pt = blockIdx.y, sg = blockIdx.x;
    const int rows_here = min(128, a.nc - pt * 128);
    const int c_base = sg * AT_SITES;
    const int n_sites = min(AT_SITES, a.C - c_base);
    uint64_t *a_ready = bars, *mma_done = bars + 2, *k_full = bars + 4;

    if (tid == 0) {
        mbar_init(a_ready, 8); mbar_init(a_ready + 1, 8);
        mbar_init(mma_done, 1); mbar_init(mma_done + 1, 1);
        mbar_init(k_full, 1); mbar_init(k_full + 1, 1); mbar_init(k_full + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(tmem_slot, 64);
    for (int i = tid; i < 4096; i += AT_THREADS) reinterpret_cast<uint4*>(sm + AT_A)[i] = make_uint4(0u, 0u, 0u, 0u);   // rows without a pair stay 0
    if (tid < 128) {
        int pi = -1, pj = -1;
        if (tid < rows_here) {
            const size_t o = (size_t)b * a.pair_stride + a.n0 + pt * 128 + tid;
            const int li = a.pair_i[o], lj = a.pair_j[o];
            if (li >= 0) { const int32_t* so = a.slot_of + (size_t)b * a.slot_stride; pi = so[li]; pj = so[lj]; }
        }
        s_pi[tid] = pi; s_pj[tid] = pj;
    }
    if (tid < 64) s_bh[tid] = a.bh[tid];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ================= control warp: K' tiles by TMA, UMMA issue =================
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        if (elect_one()) {
            for (int s = 0; s < 2 && s < n_sites; ++s) {
                mbar_expect_tx(k_full + s, 16384);
                tma_load_4d(sm + AT_K + s * 16384, &mapKh, k_full + s, 0, c_base + s, 0, b);
                tma_load_4d(sm + AT_K + s * 16384 + 8192, &mapKl, k_full + s, 0, c_base + s, 0, b);
            }
        }
        __syncwarp();
        for (int s = 0; s < n_sites; ++s) {
            const int ab = s & 1, kb = s % 3;
            mbar_wait(a_ready + ab, (s >> 1) & 1);
            mbar_wait(k_full + kb, (s / 3) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ah = smem_u32(sm + AT_A + ab * 32768), kh = smem_u32(sm + AT_K + kb * 16384);
                umma_split_k64(tmem_base, ah, ah + 16384, kh, kh + 8192, idesc, s ? 1u : 0u);
                umma_commit(mma_done + ab);
            }
            __syncwarp();
            if (s + 2 < n_sites) {      // ring slot (s+2)%3 was last read by the UMMA of site s-1
                if (s >= 1) mbar_wait(mma_done + ((s - 1) & 1), ((s - 1) >> 1) & 1);
                if (elect_one()) {
                    const int kn = (s + 2) % 3;
                    mbar_expect_tx(k_full + kn, 16384);
                    tma_load_4d(sm + AT_K + kn * 16384, &mapKh, k_full + kn, 0, c_base + s + 2, 0, b);
                    tma_load_4d(sm + AT_K + kn * 16384 + 8192, &mapKl, k_full + kn, 0, c_base + s + 2, 0, b);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= blend warps =================
        const bool quad = rows_here <= 64;                 // four threads per pair row (16 columns each) instead of two (32)
        const int row = quad ? (tid & 63) : (tid & 127);
        const int col0 = quad ? (tid >> 6) * 16 : (tid >> 7) * 32;
        const int pi = s_pi[row], pj = s_pj[row];
        const size_t tb = (size_t)b * a.tree_stride;
        const float* xi_p = a.X + tb + (size_t)max(pi, 0) * a.C * D + col0;
        const float* xj_p = a.X + tb + (size_t)max(pj, 0) * a.C * D + col0;
        const float* yi_p = a.Y + tb + (size_t)max(pi, 0) * a.C * D + col0;
        const float* yj_p = a.Y + tb + (size_t)max(pj, 0) * a.C * D + col0;
        const size_t xrow = ((size_t)b * a.pc + pt * 128 + row) * a.C;    // x-plane row of this pair, in sites
        for (int s = 0; s < n_sites; ++s) {
            const int ab = s & 1;
            if (s >= 2) mbar_wait(mma_done + ab, ((s - 2) >> 1) & 1);      // the UMMA of site s-2 has released this operand buffer
            if (pi >= 0) {
                const size_t co = (size_t)(c_base + s) * D;
                uint8_t* a_hi = sm + AT_A + ab * 32768;
                const size_t go = ((xrow + c_base + s) * 64 + col0) >> 3;
                if (quad) blend_site<16>(a, xi_p + co, xj_p + co, yi_p + co, yj_p + co, s_bh, a_hi, a_hi + 16384, row, col0, a.xh + go, a.xl + go);
                else blend_site<32>(a, xi_p + co, xj_p + co, yi_p + co, yj_p + co, s_bh, a_hi, a_hi + 16384, row, col0, a.xh + go, a.xl + go);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready + ab);
        }
    }
    // ---- partial alpha of this site group: TMEM -> alpha_part[pair][sg][slot]
    mbar_wait(mma_done + ((n_sites - 1) & 1), ((n_sites - 1) >> 1) & 1);
    tc_fence_after();
    if (warp < 8) {
        const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + hf * 32, acc);
        if (row < rows_here) {
            float* o = a.alpha_part + (((size_t)b * a.alpha_pairs + pt * 128 + row) * a.nSG + sg) * a.RP + hf * 32;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (hf * 32 + 4 * k < a.RP)
                    st4(o + 4 * k, make_float4(__uint_as_float(acc[4 * k]), __uint_as_float(acc[4 * k + 1]), __uint_as_float(acc[4 * k + 2]), __uint_as_float(acc[4 * k + 3])));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

struct ScoreTcArgs {
    const uint4* xh; const uint4* xl;      // x planes [B][pc][C][64] bf16
    int pc;                                 // pairs capacity of the x planes
    const float* alpha; int RP;             // [B][PAIR_CHUNK][RP]
    int alpha_pairs;                        // PAIR_CHUNK stride of alpha / score_part
    const int32_t* slot_of; int slot_stride;
    const int32_t* pair_i; int pair_stride; int n0; int nc;
    int Rp, S, C;
    const uint4* wsh; const uint4* wsl;     // s_out.0 weight [64][64] as bf16 planes
    const float* bg; const float* bs; const float* w2; float b2;
    const uint8_t* mask;
    float* score_part; int nSG;
};

struct EpiCtx {
    uint64_t* pb; uint8_t* a1h; uint8_t* a1l; uint32_t t_d1; uint32_t t_s;
    const float* s_bias; int my_sites; int c_base; int p; int b;
};

// Epilogue of one warp over its sites.  NSUB 16-column sub-chunks per thread (2: one of two 32-column halves of a
// 128-pair tile; 1: one of four 16-column quarters when <= 64 pairs are duplicated into rows 64..127).
template <int NSUB>
__device__ __forceinline__ float score_epilogue(const ScoreTcArgs& a, const EpiCtx& e, int prow, int n, bool row_ok, int col0, bool dup) {
    const float* bgv = e.s_bias + col0;
    const float* bsv = e.s_bias + 64 + col0;
    const float* w2v = e.s_bias + 128 + col0;
    float score = 0.f;
    for (int i = 0; i < e.my_sites; ++i) {
        const uint32_t par = i & 1;
        const int c = e.c_base + 2 * i + e.p;
        uint4 xh4[2 * NSUB], xl4[2 * NSUB];
        if (row_ok) {
            const size_t o = ((((size_t)e.b * a.pc + n) * a.C + c) * 64 + col0) >> 3;
#pragma unroll
            for (int j = 0; j < 2 * NSUB; ++j) { xh4[j] = __ldg(a.xh + o + j); xl4[j] = __ldg(a.xl + o + j); }
        } else {
#pragma unroll
            for (int j = 0; j < 2 * NSUB; ++j) { xh4[j] = make_uint4(0u, 0u, 0u, 0u); xl4[j] = xh4[j]; }
        }
        const uint32_t* xhw = reinterpret_cast<const uint32_t*>(xh4);
        const uint32_t* xlw = reinterpret_cast<const uint32_t*>(xl4);
        // ---- gate: w = sigmoid(g + b_g), x' = (1-w) x + w x_glob  -> A1 operand of the s_out GEMM
        mbar_wait(e.pb + 1, par);
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            uint32_t g[16], xg[16];
            tmem_ld16_nw(e.t_d1 + 64 + col0 + sub * 16, g);
            tmem_ld16_nw(e.t_d1 + col0 + sub * 16, xg);
            tmem_ld_wait();
            float xp[16];
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const uint32_t wh = xhw[sub * 8 + (k >> 1)], wl = xlw[sub * 8 + (k >> 1)];
                const float x0 = bf_lo(wh) + bf_lo(wl), x1 = bf_hi(wh) + bf_hi(wl);
                const float w0 = sigmoid_fast(__uint_as_float(g[k]) + bgv[sub * 16 + k]);
                const float w1 = sigmoid_fast(__uint_as_float(g[k + 1]) + bgv[sub * 16 + k + 1]);
                xp[k] = fmaf(w0, __uint_as_float(xg[k]) - x0, x0);       // (1-w) x + w x_glob
                xp[k + 1] = fmaf(w1, __uint_as_float(xg[k + 1]) - x1, x1);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint4 hh, ll;
                split2(xp[8 * j + 0], xp[8 * j + 1], hh.x, ll.x);
                split2(xp[8 * j + 2], xp[8 * j + 3], hh.y, ll.y);
                split2(xp[8 * j + 4], xp[8 * j + 5], hh.z, ll.z);
                split2(xp[8 * j + 6], xp[8 * j + 7], hh.w, ll.w);
                const int off = prow * 128 + (((((col0 + sub * 16) >> 3) + j) ^ (prow & 7)) << 4);
                *reinterpret_cast<uint4*>(e.a1h + off) = hh;
                *reinterpret_cast<uint4*>(e.a1l + off) = ll;
                if (dup) {
                    *reinterpret_cast<uint4*>(e.a1h + off + 8192) = hh;
                    *reinterpret_cast<uint4*>(e.a1l + off + 8192) = ll;
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(e.pb + 2);
        // ---- score: w2 . GELU(s + b_s) (+ b2 once per row), masked site sum
        mbar_wait(e.pb + 3, par);
        tc_fence_after();
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            uint32_t sv[16];
            tmem_ld16_nw(e.t_s + col0 + sub * 16, sv);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                acc0 = fmaf(gelu_fast(__uint_as_float(sv[k]) + bsv[sub * 16 + k]), w2v[sub * 16 + k], acc0);
                acc1 = fmaf(gelu_fast(__uint_as_float(sv[k + 1]) + bsv[sub * 16 + k + 1]), w2v[sub * 16 + k + 1], acc1);
            }
        }
        tc_fence_before();
        const bool site_ok = !(a.mask && a.mask[(size_t)e.b * a.C + c]);
        if (site_ok) score += (acc0 + acc1) + (col0 == 0 ? a.b2 : 0.f);
    }
    return row_ok ? score : 0.f;
}

__global__ void __launch_bounds__(ST_THREADS, 1)
k_score_tc(const __grid_constant__ CUtensorMap mapXh, const __grid_constant__ CUtensorMap mapXl, const ScoreTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_align1024(smem_raw);
    float* s_bias = reinterpret_cast<float*>(sm + ST_MISC);            // bg[64] | bs[64] | w2[64]
    float* s_part = s_bias + 192;                                       // [2 pipelines][4 column slots][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 1024);        // per pipeline: bx_full, d1_done, a1_ready, s_done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, pt = blockIdx.y, sg = blockIdx.x;
    const int rows_here = min(128, a.nc - pt * 128);
    const bool dup = rows_here <= 64;                                   // pairs duplicated into rows 64..127: all 16 warps stay busy
    const int act_warps = dup ? (rows_here > 32 ? 8 : 4) : 2 * ((rows_here + 31) >> 5);   // per pipeline
    const int c_base = sg * ST_SITES;
    const int n_sites = min(ST_SITES, a.C - c_base);

    // ---- one-time setup
    if (tid == 0) {
        for (int p = 0; p < 2; ++p) {
            uint64_t* pb = bars + p * 4;
            mbar_init(pb + 0, 1); mbar_init(pb + 1, 1); mbar_init(pb + 2, act_warps); mbar_init(pb + 3, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 2048; i += ST_THREADS) reinterpret_cast<uint4*>(sm + ST_A0)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 1024; i += ST_THREADS) {                      // W_s -> swizzled K-major tiles (hi, lo)
        const int plane = i >> 9, rem = i & 511, row = rem >> 3, j = rem & 7;
        const uint4* src = plane == 0 ? a.wsh : a.wsl;
        *reinterpret_cast<uint4*>(sm + ST_W + plane * 8192 + row * 128 + ((j ^ (row & 7)) << 4)) = __ldg(src + rem);
    }
    for (int i = tid; i < 1024; i += ST_THREADS) s_part[i] = 0.f;
    if (tid < 64) { s_bias[tid] = a.bg[tid]; s_bias[64 + tid] = a.bs[tid]; s_bias[128 + tid] = a.w2[tid]; }
    __syncthreads();
    {   // alpha scattered to physical-slot order: A0[row][slot] (K-major, 64 slots = one 128 B swizzle row)
        const int32_t* so = a.slot_of + (size_t)b * a.slot_stride;
        for (int idx = tid; idx < rows_here * a.Rp; idx += ST_THREADS) {
            const int row = idx / a.Rp, r = idx - row * a.Rp;
            const float v = a.alpha[((size_t)b * a.alpha_pairs + pt * 128 + row) * a.RP + r];
            const int slot = so[r];
            const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
            const int off = row * 128 + (((slot >> 3) ^ (row & 7)) << 4) + (slot & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(sm + ST_A0 + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(sm + ST_A0 + 16384 + off) = l;
            if (dup) {
                *reinterpret_cast<__nv_bfloat16*>(sm + ST_A0 + off + 8192) = h;
                *reinterpret_cast<__nv_bfloat16*>(sm + ST_A0 + 16384 + off + 8192) = l;
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 16) {
        // ================= control warp of pipeline p: TMA + MMA issue (warp-uniform, one elected lane per action) =================
        const int p = warp - 16;
        uint64_t* pb = bars + p * 4;
        uint8_t* pipe = sm + ST_PIPE + p * ST_PIPE_BYTES;
        const uint32_t a0h = smem_u32(sm + ST_A0), a0l = a0h + 16384;
        const uint32_t wsh = smem_u32(sm + ST_W), wsl = wsh + 8192;
        const uint32_t bxh = smem_u32(pipe), bxl = bxh + 16384, a1h = bxh + 32768, a1l = bxh + 49152;
        const uint32_t t_d1 = tmem_base + p * 192, t_s = t_d1 + 128;
        const uint32_t id_s = umma_idesc_bf16(128, 64), id_d1 = umma_idesc_bf16(128, 128) | (1u << 16);
        const int ksteps = (a.S + 15) >> 4;
        const int my_sites = (n_sites - p + 1) >> 1;
        if (my_sites > 0 && elect_one()) {
            const int c = c_base + p;
            mbar_expect_tx(pb + 0, 32768);
            tma_load_3d(pipe, &mapXh, pb + 0, 0, 0, b * a.C + c);
            tma_load_3d(pipe + 8192, &mapXh, pb + 0, 64, 0, b * a.C + c);
            tma_load_3d(pipe + 16384, &mapXl, pb + 0, 0, 0, b * a.C + c);
            tma_load_3d(pipe + 24576, &mapXl, pb + 0, 64, 0, b * a.C + c);
        }
        __syncwarp();
        for (int i = 0; i < my_sites; ++i) {
            const uint32_t par = i & 1;
            mbar_wait(pb + 0, par);
            tc_fence_after();
            if (elect_one()) {      // [x_glob | g] = alpha . [X | W_g X]   (B MN-major: 16 slots = 2048 B per k-step, column blocks 8 KB apart)
                for (int k = 0; k < ksteps; ++k) {
                    umma_bf16(t_d1, umma_desc_k128(a0l + k * 32), umma_desc_lbo(bxh + k * 2048, 8192), id_d1, k ? 1u : 0u);
                    umma_bf16(t_d1, umma_desc_k128(a0h + k * 32), umma_desc_lbo(bxl + k * 2048, 8192), id_d1, 1u);
                    umma_bf16(t_d1, umma_desc_k128(a0h + k * 32), umma_desc_lbo(bxh + k * 2048, 8192), id_d1, 1u);
                }
                umma_commit(pb + 1);
            }
            __syncwarp();
            if (i + 1 < my_sites) {     // prefetch the next node tile as soon as UMMA 1 has retired (hides the HBM/L2 latency)
                mbar_wait(pb + 1, par);
                if (elect_one()) {
                    const int c = c_base + 2 * (i + 1) + p;
                    mbar_expect_tx(pb + 0, 32768);
                    tma_load_3d(pipe, &mapXh, pb + 0, 0, 0, b * a.C + c);
                    tma_load_3d(pipe + 8192, &mapXh, pb + 0, 64, 0, b * a.C + c);
                    tma_load_3d(pipe + 16384, &mapXl, pb + 0, 0, 0, b * a.C + c);
                    tma_load_3d(pipe + 24576, &mapXl, pb + 0, 64, 0, b * a.C + c);
                }
                __syncwarp();
            }
            mbar_wait(pb + 2, par);     // epilogue has consumed [x_glob | g] and written x' into the A1 operand
            tc_fence_after();
            if (elect_one()) {
                for (int k = 0; k < 4; ++k) {   // s = x' . W_s^T
                    umma_bf16(t_s, umma_desc_k128(a1l + k * 32), umma_desc_k128(wsh + k * 32), id_s, k ? 1u : 0u);
                    umma_bf16(t_s, umma_desc_k128(a1h + k * 32), umma_desc_k128(wsl + k * 32), id_s, 1u);
                    umma_bf16(t_s, umma_desc_k128(a1h + k * 32), umma_desc_k128(wsh + k * 32), id_s, 1u);
                }
                umma_commit(pb + 3);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue warps: pipeline p = warp / 8, TMEM lane quarter q, column half hf =================
        const int p = warp >> 3, q = warp & 3, hf = (warp >> 2) & 1;
        const bool active = dup ? ((q & 1) * 32 < rows_here) : (q * 32 < rows_here);
        if (active) {
            uint8_t* pipe = sm + ST_PIPE + p * ST_PIPE_BYTES;
            EpiCtx e;
            e.pb = bars + p * 4; e.a1h = pipe + 32768; e.a1l = pipe + 49152;
            e.t_d1 = tmem_base + ((uint32_t)(q * 32) << 16) + p * 192; e.t_s = e.t_d1 + 128;
            e.s_bias = s_bias; e.my_sites = (n_sites - p + 1) >> 1; e.c_base = c_base; e.p = p; e.b = b;
            const int prow = dup ? ((q & 1) * 32 + lane) : (q * 32 + lane);
            const int n = pt * 128 + prow;
            const bool row_ok = n < a.nc && a.pair_i[(size_t)b * a.pair_stride + a.n0 + n] >= 0;
            if (dup) {
                const int cq = (q >> 1) * 2 + hf;
                s_part[(p * 4 + cq) * 128 + prow] = score_epilogue<1>(a, e, prow, n, row_ok, cq * 16, true);
            } else {
                s_part[(p * 4 + hf) * 128 + prow] = score_epilogue<2>(a, e, prow, n, row_ok, hf * 32, false);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < rows_here) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_part[k * 128 + tid];
        a.score_part[((size_t)b * a.alpha_pairs + pt * 128 + tid) * a.nSG + sg] = s;
    }
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------ host side

// box {64, 64, 1} over site-major node planes [B*C][S][128]  ([X | W_g X] per slot)
static int make_tmap_nodes(CUtensorMap* map, const void* base, int S, int BC) {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        enc = reinterpret_cast<PFN>(p);
    }
    cuuint64_t gdim[3] = {128, (cuuint64_t)S, (cuuint64_t)BC};
    cuuint64_t gstr[2] = {256, (cuuint64_t)S * 256};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the node planes");
    return 0;
}

// K' planes [B][S][C][64] bf16 as a 4-D tensor (d, site, slot, tree); box = all slots of one site: [64 slots][64 d]
static int make_tmap_kprime(CUtensorMap* map, const void* base, int S, int C, int B) {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        enc = reinterpret_cast<PFN>(p);
    }
    cuuint64_t gdim[4] = {64, (cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t gstr[3] = {128, (cuuint64_t)C * 128, (cuuint64_t)S * C * 128};
    cuuint32_t box[4] = {64, 1, 64, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the K' planes");
    return 0;
}

// fused blend + alpha partials: writes the x planes of pairs [n0, n0+nc) and alpha_part[b][n][sg][slot], sg over ceil(C/64) site groups
int launch_alpha_tc(const Model* m, const float* X, const float* Y, size_t tree_stride, const int32_t* slot_of, int slot_stride, const int32_t* pair_i,
                    const int32_t* pair_j, int pair_stride, int n0, int nc, int S, int C, int B, const void* kp_h, const void* kp_l, void* xh,
                    void* xl, int pc, float* alpha_part, int alpha_pairs, int nSG, int RP, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_alpha_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        attr = true;
    }
    if (S > 64) return set_error(NNJ_ERR_INVALID, "alpha_tc: at most 63 taxa on the tensor-core path");
    if ((C + AT_SITES - 1) / AT_SITES > nSG) return set_error(NNJ_ERR_INVALID, "alpha_tc: partial buffer too small");
    CUtensorMap mh, ml;
    if (int e = make_tmap_kprime(&mh, kp_h, S, C, B)) return e;
    if (int e = make_tmap_kprime(&ml, kp_l, S, C, B)) return e;
    AlphaTcArgs a;
    a.X = X; a.Y = Y; a.tree_stride = tree_stride; a.slot_of = slot_of; a.slot_stride = slot_stride;
    a.pair_i = pair_i; a.pair_j = pair_j; a.pair_stride = pair_stride; a.n0 = n0; a.nc = nc; a.C = C; a.S = S; a.bh = m->nj.bh;
    a.xh = (uint4*)xh; a.xl = (uint4*)xl; a.pc = pc; a.alpha_part = alpha_part; a.alpha_pairs = alpha_pairs; a.nSG = nSG; a.RP = RP;
    prof_begin(KC_ALPHA, st);
    k_alpha_tc<<<dim3((C + AT_SITES - 1) / AT_SITES, (nc + 127) / 128, B), AT_THREADS, AT_SMEM, st>>>(mh, ml, a);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

int launch_blend_planes(const float* X, const float* Y, size_t tree_stride, const int32_t* slot_of, int slot_stride, int C, const int32_t* pair_i,
                        const int32_t* pair_j, int pair_stride, int n0, int nc, int B, const float* bh, void* xh, void* xl, int pc,
                        cudaStream_t st) {
    prof_begin(KC_BLEND, st);
    k_blend_planes<<<dim3((C + 15) / 16, nc, B), 256, 0, st>>>(X, Y, tree_stride, slot_of, slot_stride, C, pair_i, pair_j, pair_stride, n0, bh,
                                                               (uint2*)xh, (uint2*)xl, pc);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

int launch_score_tc(const Model* m, const void* xh, const void* xl, int pc, const void* nodes_h, const void* nodes_l, const float* alpha, int RP,
                    int alpha_pairs, const int32_t* slot_of, int slot_stride, const int32_t* pair_i, int pair_stride, int n0, int nc, int Rp,
                    int S, int C, int B, const uint8_t* mask, float* score_part, int nSG, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_score_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        attr = true;
    }
    if (S > 64) return set_error(NNJ_ERR_INVALID, "score_tc: at most 63 taxa on the tensor-core pair-score path");
    CUtensorMap mh, ml;
    if (int e = make_tmap_nodes(&mh, nodes_h, S, B * C)) return e;
    if (int e = make_tmap_nodes(&ml, nodes_l, S, B * C)) return e;
    ScoreTcArgs a;
    a.xh = (const uint4*)xh; a.xl = (const uint4*)xl; a.pc = pc;
    a.alpha = alpha; a.RP = RP; a.alpha_pairs = alpha_pairs;
    a.slot_of = slot_of; a.slot_stride = slot_stride; a.pair_i = pair_i; a.pair_stride = pair_stride; a.n0 = n0; a.nc = nc;
    a.Rp = Rp; a.S = S; a.C = C;
    a.wsh = (const uint4*)m->nj_bf.wsh; a.wsl = (const uint4*)m->nj_bf.wsl;
    a.bg = m->nj.bg; a.bs = m->nj.bs; a.w2 = m->nj.w2; a.b2 = m->nj.b2;
    a.mask = mask; a.score_part = score_part; a.nSG = nSG;
    prof_begin(KC_SCORE, st);
    k_score_tc<<<dim3((C + ST_SITES - 1) / ST_SITES, (nc + 127) / 128, B), ST_THREADS, ST_SMEM, st>>>(mh, ml, a);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

}  // namespace nnj
