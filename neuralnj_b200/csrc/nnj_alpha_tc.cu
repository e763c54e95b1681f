// nnj_alpha_tc.cu — pair blend + global-attention logits of the NJ loop on tcgen05 (precision bf16x3).
//
// For every listed pair n = (i, j) of a tree and every site c the reference forms (model.py:105-118)
//     z = sigmoid(W_h (x_i - x_j) + b_h) = sigmoid(Y_i - Y_j + b_h),      x = z x_i + (1 - z) x_j
//     alpha[n, r] = sum_{c,d} (W_q x + b_q)[c,d] K_r[c,d] = sum_{c,d} x[c,d] K'_r[c,d] + kappa_r        (K' = W_q^T K)
// One CTA = one tree x 64 sites x up to 4 tiles of 128 pairs.  Per site a producer warp streams the site's node tile
// (X, Y fp32 [slots x 64], K' bf16 hi/lo [slots x 64]) into a 4-deep shared-memory ring by TMA (SWIZZLE_128B, so that
// lanes reading different slots hit different banks); 16 blend warps form x for one pair row x 16 channels per thread,
// write it (fp32) to the x planes the pair-score kernel reads later, and store its bf16 hi/lo split straight into TENSOR
// MEMORY (tcgen05.st: row = TMEM lane, two bf16 per column) as the A operand; an issue warp runs
//     acc[tile][pair, slot] += x[pair, site, :] . K'[slot, site, :]        (tcgen05.mma, A from TMEM, 3-product split)
// with a double-buffered A operand, so the tensor core works on one (site, tile) while the CUDA cores blend the next.
// Shared memory carries only the node ring: the A operand never touches it.
// With <= 64 pairs (every step after the first) rows 64..127 would idle, so the tile is split by site parity instead:
// lanes 0..63 hold the pairs at site 2k, lanes 64..127 the same pairs at site 2k+1, and two UMMA groups (K' of site 2k ->
// accumulator 0, K' of site 2k+1 -> accumulator 1) leave the even-site sums in rows 0..63 of accumulator 0 and the
// odd-site sums in rows 64..127 of accumulator 1 (the other halves are ignored).
// The accumulators are written once per CTA as partials of this site group; k_alpha_softmax reduces them in fixed order.
#include "nnj_internal.h"
#include "nnj_tc.cuh"

namespace nnj {

constexpr int AV_BLEND_WARPS = 16;
constexpr int AV_THREADS = (AV_BLEND_WARPS + 2) * 32;   // + UMMA issue warp + TMA producer warp
constexpr int AV_SITES = 64;
constexpr int AV_NST = 4;                               // node-ring depth
constexpr int AV_STAGE = 49152;                         // X lo-half 8K | X hi-half 8K | Y 8K | Y 8K | K'_hi 8K | K'_lo 8K
constexpr int AV_MAXT = 4;                              // pair tiles per CTA
constexpr int AV_MISC = AV_NST * AV_STAGE;              // pair slots (2 x 512 int) | b_h | barriers | tmem slot
constexpr int AV_SMEM = 1024 + AV_MISC + 4096 + 256 + 256 + 64;
constexpr uint32_t AV_TM_A = 256;                       // TMEM columns: acc tiles [0,256) | A buffers 2 x (hi 32 | lo 32)

struct AlphaV3Args {
    const int32_t* slot_of; int slot_stride;
    const int32_t* pair_i; const int32_t* pair_j; int pair_stride; int n0; int nc;
    int C;
    const float* bh;
    float* xf; int pc;                        // x planes [B][pc][C][64] fp32 (output)
    float* alpha_part; int alpha_pairs; int nSG; int RP;
    int dup;                                  // nc <= 64: tile split by site parity, two partials per site group
};

__global__ void __launch_bounds__(AV_THREADS, 1)
k_alpha_v3(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapKh,
           const __grid_constant__ CUtensorMap mapKl, const AlphaV3Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_align1024(smem_raw);
    int* s_pi = reinterpret_cast<int*>(sm + AV_MISC);          // physical slots of the pair rows (-1: no pair)
    int* s_pj = s_pi + AV_MAXT * 128;
    float* s_bh = reinterpret_cast<float*>(s_pj + AV_MAXT * 128);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bh + 64);   // full[4], stage_free[4], a_ready[2], a_free[2], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    uint64_t *full = bars, *stage_free = bars + 4, *a_ready = bars + 8, *a_free = bars + 10, *done = bars + 12;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, sg = blockIdx.x;
    const int rows_total = a.nc - blockIdx.y * (AV_MAXT * 128);           // pair rows handled by this CTA
    const int rows_cta = min(rows_total, AV_MAXT * 128);
    const int row0 = blockIdx.y * (AV_MAXT * 128);
    const int NT = (rows_cta + 127) >> 7;
    const bool dup = a.dup != 0;                                          // site-parity split of a half-empty tile
    const int c_base = sg * AV_SITES;
    const int n_sites = min(AV_SITES, a.C - c_base);
    const int n_items = dup ? (n_sites + 1) >> 1 : n_sites * NT;

    if (tid == 0) {
        for (int i = 0; i < AV_NST; ++i) { mbar_init(full + i, 1); mbar_init(stage_free + i, (dup ? AV_BLEND_WARPS / 2 : AV_BLEND_WARPS) + 1); }
        mbar_init(a_ready, AV_BLEND_WARPS); mbar_init(a_ready + 1, AV_BLEND_WARPS);
        mbar_init(a_free, 1); mbar_init(a_free + 1, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == AV_BLEND_WARPS) tmem_alloc(tmem_slot, 512);
    for (int r = tid; r < AV_MAXT * 128; r += AV_THREADS) {
        int pi = -1, pj = -1;
        if (r < rows_cta) {
            const size_t o = (size_t)b * a.pair_stride + a.n0 + row0 + r;
            const int li = a.pair_i[o], lj = a.pair_j[o];
            if (li >= 0) { const int32_t* so = a.slot_of + (size_t)b * a.slot_stride; pi = so[li]; pj = so[lj]; }
        }
        s_pi[r] = pi; s_pj[r] = pj;
    }
    if (tid < 64) s_bh[tid] = a.bh[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == AV_BLEND_WARPS + 1) {
        // ================= producer warp: node tiles of one site per ring stage =================
        for (int s = 0; s < n_sites; ++s) {
            const int st = s % AV_NST;
            if (s >= AV_NST) mbar_wait(stage_free + st, ((s / AV_NST) - 1) & 1);
            if (elect_one()) {
                uint8_t* stg = sm + st * AV_STAGE;
                const int c = c_base + s;
                mbar_expect_tx(full + st, AV_STAGE);
                tma_load_4d(stg, &mapX, full + st, 0, c, 0, b);
                tma_load_4d(stg + 8192, &mapX, full + st, 32, c, 0, b);
                tma_load_4d(stg + 16384, &mapY, full + st, 0, c, 0, b);
                tma_load_4d(stg + 24576, &mapY, full + st, 32, c, 0, b);
                tma_load_4d(stg + 32768, &mapKh, full + st, 0, c, 0, b);
                tma_load_4d(stg + 40960, &mapKl, full + st, 0, c, 0, b);
            }
            __syncwarp();
        }
    } else if (warp == AV_BLEND_WARPS) {
        // ================= issue warp: acc (+)= A (tensor memory) . K'^T =================
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        if (!dup) {
            int k = 0;
            for (int s = 0; s < n_sites; ++s) {
                const int st = s % AV_NST;
                mbar_wait(full + st, (s / AV_NST) & 1);
                const uint32_t kh = smem_u32(sm + st * AV_STAGE + 32768), kl = kh + 8192;
                for (int t = 0; t < NT; ++t, ++k) {
                    const int buf = k & 1;
                    mbar_wait(a_ready + buf, (k >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ta = tmem_base + AV_TM_A + buf * 64, td = tmem_base + t * 64;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_bf16_ta(td, ta + 32 + kk * 8, umma_desc_k128(kh + kk * 32), idesc, (s | kk) ? 1u : 0u);   // small terms first
                            umma_bf16_ta(td, ta + kk * 8, umma_desc_k128(kl + kk * 32), idesc, 1u);
                            umma_bf16_ta(td, ta + kk * 8, umma_desc_k128(kh + kk * 32), idesc, 1u);
                        }
                        umma_commit(a_free + buf);
                        if (t == NT - 1) umma_commit(stage_free + st);
                    }
                    __syncwarp();
                }
            }
        } else {
            for (int k = 0; k < n_items; ++k) {
                const int buf = k & 1;
                mbar_wait(a_ready + buf, (k >> 1) & 1);
                for (int h = 0; h < 2; ++h) {
                    const int s = 2 * k + h;
                    if (s >= n_sites) break;
                    const int st = s % AV_NST;
                    mbar_wait(full + st, (s / AV_NST) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t kh = smem_u32(sm + st * AV_STAGE + 32768), kl = kh + 8192;
                        const uint32_t ta = tmem_base + AV_TM_A + buf * 64, td = tmem_base + h * 64;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_bf16_ta(td, ta + 32 + kk * 8, umma_desc_k128(kh + kk * 32), idesc, (k | kk) ? 1u : 0u);
                            umma_bf16_ta(td, ta + kk * 8, umma_desc_k128(kl + kk * 32), idesc, 1u);
                            umma_bf16_ta(td, ta + kk * 8, umma_desc_k128(kh + kk * 32), idesc, 1u);
                        }
                        umma_commit(stage_free + st);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(a_free + buf);
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else {
        // ================= blend warps: TMEM lane quarter q, channels [16 cg, 16 cg + 16) =================
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int half_off = (cg >> 1) * 8192, j0 = (cg & 1) * 4;
        float bh[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) bh[e] = s_bh[cg * 16 + e];
        for (int k = 0; k < n_items; ++k) {
            int s, row;                      // site (relative) and pair row of this thread for item k
            if (dup) { s = 2 * k + (q >> 1); row = (q & 1) * 32 + lane; }
            else { s = k / NT; row = (k - s * NT) * 128 + q * 32 + lane; }
            const bool site_ok = s < n_sites;
            const int st = s % AV_NST, buf = k & 1;
            const int pi = s_pi[row], pj = s_pj[row];
            if (site_ok) mbar_wait(full + st, (s / AV_NST) & 1);
            if (k >= 2) { mbar_wait(a_free + buf, ((k >> 1) - 1) & 1); tc_fence_after(); }
            uint32_t hh[8], ll[8];
            if (site_ok && pi >= 0) {
                const uint8_t* xb = sm + st * AV_STAGE + half_off;
                const uint8_t* xi_r = xb + pi * 128;
                const uint8_t* xj_r = xb + pj * 128;
                float4* xo = reinterpret_cast<float4*>(a.xf + (((size_t)b * a.pc + row0 + row) * a.C + c_base + s) * D + cg * 16);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ci = ((j0 + e) ^ (pi & 7)) << 4, cj = ((j0 + e) ^ (pj & 7)) << 4;
                    const float4 xa = *reinterpret_cast<const float4*>(xi_r + ci), xc = *reinterpret_cast<const float4*>(xj_r + cj);
                    const float4 ya = *reinterpret_cast<const float4*>(xi_r + 16384 + ci), yc = *reinterpret_cast<const float4*>(xj_r + 16384 + cj);
                    float4 v;
                    v.x = fmaf(sigmoid_fast(ya.x - yc.x + bh[4 * e + 0]), xa.x - xc.x, xc.x);   // z x_i + (1-z) x_j
                    v.y = fmaf(sigmoid_fast(ya.y - yc.y + bh[4 * e + 1]), xa.y - xc.y, xc.y);
                    v.z = fmaf(sigmoid_fast(ya.z - yc.z + bh[4 * e + 2]), xa.z - xc.z, xc.z);
                    v.w = fmaf(sigmoid_fast(ya.w - yc.w + bh[4 * e + 3]), xa.w - xc.w, xc.w);
                    split2(v.x, v.y, hh[2 * e], ll[2 * e]);
                    split2(v.z, v.w, hh[2 * e + 1], ll[2 * e + 1]);
                    xo[e] = v;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) { hh[e] = 0u; ll[e] = 0u; }
            }
            tmem_st8(lane_base + AV_TM_A + buf * 64 + cg * 8, hh);
            tmem_st8(lane_base + AV_TM_A + buf * 64 + 32 + cg * 8, ll);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(a_ready + buf);
                // this warp's last read of the stage: the last tile of the site (or, split by parity, its only item)
                if (site_ok && (dup || (k - s * NT) == NT - 1)) mbar_arrive(stage_free + st);
            }
        }
    }
    // ---- partial alpha of this site group: TMEM -> alpha_part[pair][partial][slot]
    mbar_wait(done, 0);
    tc_fence_after();
    if (warp < AV_BLEND_WARPS) {
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        if (cg * 16 < a.RP) {
            if (!dup) {
                for (int t = 0; t < NT; ++t) {
                    uint32_t acc[16];
                    tmem_ld16_nw(lane_base + t * 64 + cg * 16, acc);
                    tmem_ld_wait();
                    const int row = t * 128 + q * 32 + lane;
                    if (row < rows_cta) {
                        float* o = a.alpha_part + (((size_t)b * a.alpha_pairs + row0 + row) * a.nSG + sg) * a.RP + cg * 16;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (cg * 16 + 4 * e < a.RP)
                                st4(o + 4 * e, make_float4(__uint_as_float(acc[4 * e]), __uint_as_float(acc[4 * e + 1]), __uint_as_float(acc[4 * e + 2]), __uint_as_float(acc[4 * e + 3])));
                    }
                }
            } else {
                const int h = q >> 1, row = (q & 1) * 32 + lane;
                uint32_t acc[16];
                tmem_ld16_nw(lane_base + h * 64 + cg * 16, acc);
                tmem_ld_wait();
                const bool have = h < n_sites;             // the odd-site accumulator is never written when the group has one site
                if (row < rows_cta) {
                    float* o = a.alpha_part + (((size_t)b * a.alpha_pairs + row0 + row) * a.nSG + 2 * sg + h) * a.RP + cg * 16;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (cg * 16 + 4 * e < a.RP)
                            st4(o + 4 * e, have ? make_float4(__uint_as_float(acc[4 * e]), __uint_as_float(acc[4 * e + 1]), __uint_as_float(acc[4 * e + 2]), __uint_as_float(acc[4 * e + 3]))
                                                : make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == AV_BLEND_WARPS) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_enc get_enc() {
    static PFN_enc enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_enc>(p);
    }
    return enc;
}

// fp32 node pool [B][S][C][64] (tree stride given) as a 4-D tensor (d, site, slot, tree); box = 32 channels of all slots at one site
static int make_tmap_pool_f32(CUtensorMap* map, const float* base, size_t tree_stride, int S, int C, int B) {
    PFN_enc enc = get_enc();
    if (!enc) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4] = {64, (cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t gstr[3] = {256, (cuuint64_t)C * 256, (cuuint64_t)tree_stride * 4};
    cuuint32_t box[4] = {32, 1, 64, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the fp32 node pool");
    return 0;
}

// K' planes [B][S][C][64] bf16 as a 4-D tensor (d, site, slot, tree); box = all slots of one site: [64 slots][64 d]
static int make_tmap_kprime(CUtensorMap* map, const void* base, int S, int C, int B) {
    PFN_enc enc = get_enc();
    if (!enc) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4] = {64, (cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t gstr[3] = {128, (cuuint64_t)C * 128, (cuuint64_t)S * C * 128};
    cuuint32_t box[4] = {64, 1, 64, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled failed for the K' planes");
    return 0;
}

// writes the x planes of pairs [n0, n0+nc) and alpha_part[b][n][partial][slot]; returns the number of partials per pair in *n_part
int launch_alpha_tc(const Model* m, const float* X, const float* Y, size_t tree_stride, const int32_t* slot_of, int slot_stride, const int32_t* pair_i,
                    const int32_t* pair_j, int pair_stride, int n0, int nc, int S, int C, int B, const void* kp_h, const void* kp_l, float* xf,
                    int pc, float* alpha_part, int alpha_pairs, int nSG, int RP, int* n_part, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_alpha_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, AV_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        attr = true;
    }
    if (S > 64) return set_error(NNJ_ERR_INVALID, "alpha_tc: at most 63 taxa on the tensor-core path");
    const int groups = (C + AV_SITES - 1) / AV_SITES;
    const bool dup = nc <= 64;
    *n_part = dup ? 2 * groups : groups;
    if (*n_part > nSG) return set_error(NNJ_ERR_INVALID, "alpha_tc: partial buffer too small");
    CUtensorMap mx, my, mh, ml;
    if (int e = make_tmap_pool_f32(&mx, X, tree_stride, S, C, B)) return e;
    if (int e = make_tmap_pool_f32(&my, Y, tree_stride, S, C, B)) return e;
    if (int e = make_tmap_kprime(&mh, kp_h, S, C, B)) return e;
    if (int e = make_tmap_kprime(&ml, kp_l, S, C, B)) return e;
    AlphaV3Args a;
    a.slot_of = slot_of; a.slot_stride = slot_stride;
    a.pair_i = pair_i; a.pair_j = pair_j; a.pair_stride = pair_stride; a.n0 = n0; a.nc = nc; a.C = C; a.bh = m->nj.bh;
    a.dup = dup ? 1 : 0;
    a.xf = xf; a.pc = pc; a.alpha_part = alpha_part; a.alpha_pairs = alpha_pairs; a.nSG = nSG; a.RP = RP;
    prof_begin(KC_ALPHA, st);
    k_alpha_v3<<<dim3(groups, (nc + AV_MAXT * 128 - 1) / (AV_MAXT * 128), B), AV_THREADS, AV_SMEM, st>>>(mx, my, mh, ml, a);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

}  // namespace nnj
