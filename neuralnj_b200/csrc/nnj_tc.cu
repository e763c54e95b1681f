// nnj_tc.cu — tcgen05 / TMEM / TMA split-bf16 batched GEMM for sm_100a.
//
//   C[z] (M x N, fp32) = A[z] (M x K) * B[z]^T (N x K),   A = A_hi + A_lo, B = B_hi + B_lo   (bf16 planes)
//   computed as A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32 accumulation in TMEM (~16-bit mantissa operands,
//   the precision at which Argmax topologies stay identical to the fp32 reference).
//
// Used for the tied row attention (axial_attention.py:97,114): logits S = Q K^T (K = R*8) and ctx = P V
// (B = V^T, K = C).  One CTA per 128 x 128 output tile, K walked in 64-element (128 B, SWIZZLE_128B) chunks through
// a 3-stage TMA -> mbarrier -> tcgen05.mma pipeline:
//   warp 0   : TMA producer (one elected lane), 4 boxes per stage (A_hi, A_lo, B_hi, B_lo; 64 KB)
//   warp 1   : TMEM allocator + single-thread MMA issuer (12 x UMMA 128x128x16 per stage), tcgen05.commit
//   warps 2-5: epilogue, one TMEM lane quarter each: tcgen05.ld -> registers -> global (fp32, or softmax-ready)
// Every mbarrier wait is bounded (trap after ~20 s) so that a protocol bug surfaces as an error, not a hang.
//
// CTA-pair kernels (cta_group::2, clusters of two CTAs on one TPC; further down): k_tc_gemm2 (256 x 256 tiles, half the B tile per SM),
// k_tc_gemm2s (Q K^T with the row softmax in the epilogue - the default of the tied row attention), k_tc_gemm2w (P V in one pass over P).
// The single-CTA k_tc_gemm serves shapes without an even number of 128-row blocks, split-K launches, the 64 / 128-column tiles and the
// one-product mode (NNJ_PREC_BF16).
#include "nnj_internal.h"
#include "nnj_tc.cuh"

namespace nnj {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_PLANE_BYTES = TC_BM * TC_BK * 2;          // 16 KB: 128 rows x 128 B
constexpr int TC_STAGE_BYTES = 4 * TC_PLANE_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int TC_THREADS = 192;
constexpr int TC_EPI_LD = 36;                              // floats per row of an epilogue staging tile [32 rows][32 cols] (+4: conflict-free)
constexpr int TC_EPI_BYTES = 4 * 32 * TC_EPI_LD * 4;      // one tile per epilogue warp

// grid (N tiles * nsplit, M tiles, Z).  C row-major with leading dimension ldc, batch stride sC (elements).
// Split-K: split s = blockIdx.x % nsplit handles K chunks [s*cps, (s+1)*cps) and writes to C + s*split_stride.
struct TcGemmArgs {
    float* C; int M, N, K, ldc; size_t sC; int Z;
    int nsplit, chunks_per_split; size_t split_stride;
    const float* row_div = nullptr;     // MN-major instantiations: output row m of batch z is divided by row_div[z * M + m] (P V after k_tc_gemm2s)
};

// BMN: B is given MN-major, [Z][K][N] with N contiguous (e.g. V of the row attention, [key site][taxon*8+d]); its tiles are loaded
// as 64 x 64 boxes (64 N values = one 128-byte swizzle row per K index), the 64-column blocks of an N tile 8 KB apart.
//
// Persistent: grid = min(tiles, SMs); every role walks the same static tile list (groups of N tiles dealt out round robin, see below).  The TMA ring and its barriers run on across tiles, and the accumulator
// is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i drains while the tensor core works on tile i+1.
//
// ONE (precision NNJ_PREC_BF16, plain bf16 operands): only A_hi * B_hi is formed.  The lo planes are never fetched; their slots of
// a physical stage hold a second logical stage instead (ring depth 2 * NSTG at half the bytes per stage), one UMMA per k-step.
template <int BN, bool BMN, bool ONE>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
          const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcGemmArgs g) {
    constexpr int B_PLANE = BN * TC_BK * 2;
    constexpr int STAGE = 2 * TC_PLANE_BYTES + 2 * B_PLANE;
    constexpr int NSTG = BN == 256 ? 2 : TC_STAGES;    // 128 x 256 tiles: 96 KB per stage (a 128 x 256 x 16 UMMA runs at 75 % of the tensor peak, 128 x 128 at 60 %)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_align1024(smem_raw);
    constexpr int NL = ONE ? 2 * NSTG : NSTG;          // logical ring depth
    constexpr int LSTAGE = ONE ? STAGE / 2 : STAGE;    // bytes a logical stage receives
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + NSTG * STAGE);
    uint64_t* empty = full + NL;
    uint64_t* acc_full = empty + NL;        // [2]
    uint64_t* acc_free = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
    float* epi = reinterpret_cast<float*>(tiles + NSTG * STAGE + 256);     // [4 warps][32][TC_EPI_LD]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_chunks = (g.K + TC_BK - 1) / TC_BK;
    const int nt = ((g.N + BN - 1) / BN) * g.nsplit, mt = (g.M + TC_BM - 1) / TC_BM;
    const int n_tiles = g.Z * mt * nt;
    // static schedule: the nt N tiles of one 128-row A block form a group that runs back to back on one SM (the A block is re-read
    // from that SM's own L2 partition: with single tiles dealt out round robin, neighbouring CTAs on different dies each pulled
    // their own copy of P from DRAM); groups are dealt out round robin, so neighbouring CTAs work on neighbouring row blocks of the
    // same (tree, head) and share its B operand in L2.  (Contiguous tile ranges per CTA were tried: 148 different (tree, head)
    // working sets at once do not fit L2 and both GEMMs ran 1.5-1.9x slower.)
    // Measured (512 trees): P V (two N tiles, A = P streamed from DRAM) 64.7 -> 60.6 ms with groups; Q K^T (four N tiles, operands L2
    // resident) 56.0 -> 60.9 ms, so only the MN-major instantiation groups.  Small problems: single tiles, so that every SM gets work.
    const int gsz = (BMN && g.Z * mt >= (int)gridDim.x) ? nt : 1;
    const int n_groups = n_tiles / gsz;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NL; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_full + 1, 1);
        mbar_init(acc_free, 4); mbar_init(acc_free + 1, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);   // two 128 x BN fp32 accumulators
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Producer / issuer warps keep warp-uniform control flow and elect one lane per action, so descriptors and
    // addresses stay in uniform registers (a lane==0 branch makes ptxas emit a R2UR broadcast loop per MMA).
    if (warp == 0) {
        int gc = 0;     // chunks issued so far (ring position)
        for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t) {
            const int xi = t % nt, r0 = t / nt, z = r0 / mt;
            const int split = xi % g.nsplit, n0 = (xi / g.nsplit) * BN, m0 = (r0 - z * mt) * TC_BM;
            const int kc0 = split * g.chunks_per_split;
            const int nchunks = max(0, min(total_chunks - kc0, g.chunks_per_split));
            for (int kk = 0; kk < nchunks; ++kk, ++gc) {
                const int s = gc % NL, kc = kc0 + kk;
                if (gc >= NL) mbar_wait(&empty[s], ((gc / NL) - 1) & 1);
                // ONE: logical stage s lives in physical stage s / 2, in the hi (even s) or lo (odd s) plane slots
                uint8_t* st = ONE ? tiles + (s >> 1) * STAGE + (s & 1) * TC_PLANE_BYTES : tiles + s * STAGE;
                uint8_t* stb = ONE ? tiles + (s >> 1) * STAGE + 2 * TC_PLANE_BYTES + (s & 1) * B_PLANE : st + 2 * TC_PLANE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(&full[s], LSTAGE);
                    tma_load_3d(st, &mapAh, &full[s], kc * TC_BK, m0, z);
                    if (!ONE) tma_load_3d(st + TC_PLANE_BYTES, &mapAl, &full[s], kc * TC_BK, m0, z);
                    if (BMN) {
#pragma unroll
                        for (int nb = 0; nb < BN / 64; ++nb) {
                            tma_load_3d(stb + nb * 8192, &mapBh, &full[s], n0 + nb * 64, kc * TC_BK, z);
                            if (!ONE) tma_load_3d(stb + B_PLANE + nb * 8192, &mapBl, &full[s], n0 + nb * 64, kc * TC_BK, z);
                        }
                    } else {
                        tma_load_3d(stb, &mapBh, &full[s], kc * TC_BK, n0, z);
                        if (!ONE) tma_load_3d(stb + B_PLANE, &mapBl, &full[s], kc * TC_BK, n0, z);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        int gc = 0, ti = 0;
        for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t, ++ti) {
            const int split = (t % nt) % g.nsplit;
            // a ragged last N tile runs a narrower UMMA (N rounded up to 16): the tensor pipe time of an instruction is ~N/2 clk
            const int n_left = g.N - ((t % nt) / g.nsplit) * BN;
            const uint32_t idesc = umma_idesc_bf16(TC_BM, n_left >= BN ? BN : ((n_left + 15) & ~15)) | (BMN ? (1u << 16) : 0u);
            const int kc0 = split * g.chunks_per_split;
            const int nchunks = max(0, min(total_chunks - kc0, g.chunks_per_split));
            const int ab = ti & 1;
            const uint32_t t_acc = tmem_base + ab * BN;
            if (ti >= 2) { mbar_wait(acc_free + ab, ((ti >> 1) - 1) & 1); tc_fence_after(); }   // the epilogue has drained this accumulator
            for (int kk = 0; kk < nchunks; ++kk, ++gc) {
                const int s = gc % NL, kc = kc0 + kk;
                mbar_wait(&full[s], (gc / NL) & 1);
                tc_fence_after();
                const uint32_t a_hi = ONE ? smem_u32(tiles + (s >> 1) * STAGE + (s & 1) * TC_PLANE_BYTES) : smem_u32(tiles + s * STAGE);
                const uint32_t a_lo = a_hi + TC_PLANE_BYTES;
                const uint32_t b_hi = ONE ? smem_u32(tiles + (s >> 1) * STAGE + 2 * TC_PLANE_BYTES + (s & 1) * B_PLANE) : a_hi + 2 * TC_PLANE_BYTES;
                const uint32_t b_lo = b_hi + B_PLANE;
                const int kvalid = min(TC_BK, g.K - kc * TC_BK);
                const int ksteps = (kvalid + 15) / 16;
                if (elect_one()) {
                    // descriptor low words advance by plain adds: K-major 16 bf16 = 32 B (>> 4 = 2) per k-step; MN-major 16 K rows = 2 KB (128)
                    const uint32_t dah = umma_desc_lo(a_hi), dal = umma_desc_lo(a_lo);
                    const uint32_t dbh = BMN ? umma_desc_lo(b_hi, 8192) : umma_desc_lo(b_hi), dbl = BMN ? umma_desc_lo(b_lo, 8192) : umma_desc_lo(b_lo);
                    constexpr uint32_t BSTEP = BMN ? 128u : 2u;
                    if (ONE) {
                        if (kk == 0) umma_ss<false>(t_acc, dah, dbh, idesc); else umma_ss<true>(t_acc, dah, dbh, idesc);
#pragma unroll
                        for (int k = 1; k < 4; ++k)
                            if (k < ksteps) umma_ss<true>(t_acc, dah + k * 2, dbh + k * BSTEP, idesc);
                    } else {
                        if (kk == 0) umma_ss<false>(t_acc, dal, dbh, idesc); else umma_ss<true>(t_acc, dal, dbh, idesc);   // small terms first
                        umma_ss<true>(t_acc, dah, dbl, idesc);
                        umma_ss<true>(t_acc, dah, dbh, idesc);
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            if (k < ksteps) {
                                umma_ss<true>(t_acc, dal + k * 2, dbh + k * BSTEP, idesc);
                                umma_ss<true>(t_acc, dah + k * 2, dbl + k * BSTEP, idesc);
                                umma_ss<true>(t_acc, dah + k * 2, dbh + k * BSTEP, idesc);
                            }
                        }
                    }
                    umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
                    if (kk == nchunks - 1) umma_commit(acc_full + ab);   // accumulator complete
                }
                __syncwarp();
            }
            if (nchunks == 0 && elect_one()) mbar_arrive(acc_full + ab);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        int ti = 0;
        for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t, ++ti) {
            const int xi = t % nt, r0 = t / nt, z = r0 / mt;
            const int split = xi % g.nsplit, n0 = (xi / g.nsplit) * BN, m0 = (r0 - z * mt) * TC_BM;
            const int kc0 = split * g.chunks_per_split;
            const int nchunks = max(0, min(total_chunks - kc0, g.chunks_per_split));
            const int ab = ti & 1;
            mbar_wait(acc_full + ab, (ti >> 1) & 1);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            float* crow = g.C + (size_t)z * g.sC + (size_t)m * g.ldc + (size_t)split * g.split_stride;
            float* stg = epi + q * (32 * TC_EPI_LD);
            const bool vec_ok = (g.ldc & 3) == 0 && (g.split_stride & 3) == 0 && (g.sC & 3) == 0;
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; ++cb) {
                uint32_t v[32];
                if (nchunks > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + cb * 32, v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (BMN && g.row_div) {
                    const float rs = m < g.M ? 1.0f / g.row_div[(size_t)z * g.M + m] : 1.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * rs);
                }
                const int n = n0 + cb * 32;
                if (n + 31 < g.N && vec_ok) {
                    // TMEM hands a thread one ROW (32 columns): stored directly, every instruction would touch 32 rows x 16 B.  The tile goes
                    // through a padded shared-memory tile instead and leaves as 4 rows x 128 contiguous bytes per instruction.
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4(stg + lane * TC_EPI_LD + j * 4, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    __syncwarp();
                    const int rr = lane >> 3, c4 = (lane & 7) * 4;
                    float* cbase = g.C + (size_t)z * g.sC + (size_t)split * g.split_stride + (size_t)(m0 + q * 32) * g.ldc + n + c4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rr;
                        if (m0 + q * 32 + row < g.M) st4(cbase + (size_t)row * g.ldc, ld4(stg + row * TC_EPI_LD + c4));
                    }
                    __syncwarp();
                } else if (m < g.M) {
                    {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < g.N) crow[n + j] = __uint_as_float(v[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free + ab);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2), 256 x 256 tiles, 3-product split
// k_tc_gemm<256, .> moves 96 KB per k-chunk and SM (A 32 KB + B 64 KB, hi + lo) and sits at ~86 % of the chip's L2 -> SM throughput with
// the tensor pipe 55 % busy.  Here two CTAs of a cluster (one TPC) share a 256-row tile: each loads its own 128 rows of A and only HALF of
// the B tile (64 KB per chunk and SM), the rank-0 CTA issues 256 x 256 x 16 UMMAs that read both halves, and each CTA drains its own 128
// accumulator rows.  64 KB stages -> a 3-deep ring.  Barriers (same offsets in both CTAs):
//   full[s]     leader only: its producer expects the bytes of BOTH CTAs' loads (each TMA names the leader's barrier)
//   empty[s]    both: the leader's tcgen05.commit multicasts the arrival when the stage's MMAs have retired
//   acc_full[b] both: multicast commit after the last k-chunk of a tile
//   acc_free[b] leader only, 8 arrivals: the four epilogue warps of both CTAs (the peer's arrive remotely)
// Requires an even number of 128-row blocks; N tiles of 256 (a ragged last tile runs an UMMA of N = 16 ceil(n_left / 16), half per CTA).
constexpr int G2_STAGE = 4 * TC_PLANE_BYTES;      // A_hi, A_lo, B_hi (half), B_lo (half): 64 KB
constexpr int G2_NSTG = 3;
constexpr int G2_SMEM = G2_NSTG * G2_STAGE + 1024 + 256 + TC_EPI_BYTES;
template <bool BMN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
           const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcGemmArgs g) {
    constexpr int BN = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_align1024(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + G2_NSTG * G2_STAGE);
    uint64_t* empty = full + G2_NSTG;
    uint64_t* acc_full = empty + G2_NSTG;      // [2]
    uint64_t* acc_free = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
    float* epi = reinterpret_cast<float*>(tiles + G2_NSTG * G2_STAGE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int total_chunks = (g.K + TC_BK - 1) / TC_BK;
    const int nt = (g.N + BN - 1) / BN, mt2 = (g.M + 2 * TC_BM - 1) / (2 * TC_BM);
    const int n_tiles = g.Z * mt2 * nt;
    const int gsz = (BMN && g.Z * mt2 >= ncl) ? nt : 1;      // see k_tc_gemm: the N tiles of a row block back to back on one SM pair
    const int n_groups = n_tiles / gsz;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G2_NSTG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_full + 1, 1);
        mbar_init(acc_free, rank == 0 ? 5 : 4); mbar_init(acc_free + 1, rank == 0 ? 5 : 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int gc = 0;
        for (int grp = cid; grp < n_groups; grp += ncl) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t) {
            const int xi = t % nt, r0 = t / nt, z = r0 / mt2;
            const int n0 = xi * BN, m0 = (r0 - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            const int n_left = g.N - n0;
            const int n_umma = n_left >= BN ? BN : ((n_left + 15) & ~15);
            const int nb0 = n0 + (int)rank * (n_umma >> 1);           // this CTA's half of the B tile
            for (int kc = 0; kc < total_chunks; ++kc, ++gc) {
                const int s = gc % G2_NSTG;
                if (gc >= G2_NSTG) mbar_wait(&empty[s], ((gc / G2_NSTG) - 1) & 1);
                uint8_t* st = tiles + s * G2_STAGE;
                if (elect_one()) {
                    const uint32_t fb = mapa_u32(smem_u32(&full[s]), 0);
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * G2_STAGE);
                    tma_load_3d_2sm(st, &mapAh, fb, kc * TC_BK, m0, z);
                    tma_load_3d_2sm(st + TC_PLANE_BYTES, &mapAl, fb, kc * TC_BK, m0, z);
                    if (BMN) {
#pragma unroll
                        for (int nb = 0; nb < 2; ++nb) {
                            tma_load_3d_2sm(st + 2 * TC_PLANE_BYTES + nb * 8192, &mapBh, fb, nb0 + nb * 64, kc * TC_BK, z);
                            tma_load_3d_2sm(st + 3 * TC_PLANE_BYTES + nb * 8192, &mapBl, fb, nb0 + nb * 64, kc * TC_BK, z);
                        }
                    } else {
                        tma_load_3d_2sm(st + 2 * TC_PLANE_BYTES, &mapBh, fb, kc * TC_BK, nb0, z);
                        tma_load_3d_2sm(st + 3 * TC_PLANE_BYTES, &mapBl, fb, kc * TC_BK, nb0, z);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            int gc = 0, ti = 0;
            for (int grp = cid; grp < n_groups; grp += ncl) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t, ++ti) {
                const int n_left = g.N - (t % nt) * BN;
                const uint32_t idesc = umma_idesc_bf16(2 * TC_BM, n_left >= BN ? BN : ((n_left + 15) & ~15)) | (BMN ? (1u << 16) : 0u);
                const int ab = ti & 1;
                const uint32_t t_acc = tmem_base + ab * BN;
                if (ti >= 2) { mbar_wait(acc_free + ab, ((ti >> 1) - 1) & 1); tc_fence_after(); }
                for (int kc = 0; kc < total_chunks; ++kc, ++gc) {
                    const int s = gc % G2_NSTG;
                    mbar_wait(&full[s], (gc / G2_NSTG) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(tiles + s * G2_STAGE), a_lo = a_hi + TC_PLANE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * TC_PLANE_BYTES, b_lo = b_hi + TC_PLANE_BYTES;
                    const int kvalid = min(TC_BK, g.K - kc * TC_BK);
                    const int ksteps = (kvalid + 15) / 16;
                    if (elect_one()) {
                        const uint32_t dah = umma_desc_lo(a_hi), dal = umma_desc_lo(a_lo);
                        const uint32_t dbh = BMN ? umma_desc_lo(b_hi, 8192) : umma_desc_lo(b_hi), dbl = BMN ? umma_desc_lo(b_lo, 8192) : umma_desc_lo(b_lo);
                        constexpr uint32_t BSTEP = BMN ? 128u : 2u;
                        if (kc == 0) umma_ss2<false>(t_acc, dal, dbh, idesc); else umma_ss2<true>(t_acc, dal, dbh, idesc);   // small terms first
                        umma_ss2<true>(t_acc, dah, dbl, idesc);
                        umma_ss2<true>(t_acc, dah, dbh, idesc);
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            if (k < ksteps) {
                                umma_ss2<true>(t_acc, dal + k * 2, dbh + k * BSTEP, idesc);
                                umma_ss2<true>(t_acc, dah + k * 2, dbl + k * BSTEP, idesc);
                                umma_ss2<true>(t_acc, dah + k * 2, dbh + k * BSTEP, idesc);
                            }
                        }
                        umma_commit2_mc(&empty[s], 3);                              // frees the stage in both CTAs
                        if (kc == total_chunks - 1) umma_commit2_mc(acc_full + ab, 3);   // accumulator complete, both CTAs
                    }
                    __syncwarp();
                }
            }
        } else {      // relay: the peer's epilogue warps arrive locally (cheap CTA-scope release); one cluster-scope arrive per tile goes to the leader
            const uint32_t free0 = mapa_u32(smem_u32(acc_free), 0);
            int ti = 0;
            for (int grp = cid; grp < n_groups; grp += ncl) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t, ++ti) {
                const int ab = ti & 1;
                mbar_wait(acc_free + ab, (ti >> 1) & 1);
                if (elect_one()) mbar_arrive_cluster(free0 + ab * 8);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        int ti = 0;
        for (int grp = cid; grp < n_groups; grp += ncl) for (int t = grp * gsz; t < (grp + 1) * gsz; ++t, ++ti) {
            const int xi = t % nt, r0 = t / nt, z = r0 / mt2;
            const int n0 = xi * BN, m0 = (r0 - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            const int ab = ti & 1;
            mbar_wait(acc_full + ab, (ti >> 1) & 1);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            float* crow = g.C + (size_t)z * g.sC + (size_t)m * g.ldc;
            float* stg = epi + q * (32 * TC_EPI_LD);
            const bool vec_ok = (g.ldc & 3) == 0 && (g.sC & 3) == 0;
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; ++cb) {
                const int n = n0 + cb * 32;
                if (n >= g.N) break;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + cb * 32, v);
                if (BMN && g.row_div) {
                    const float rs = m < g.M ? 1.0f / g.row_div[(size_t)z * g.M + m] : 1.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * rs);
                }
                if (n + 31 < g.N && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4(stg + lane * TC_EPI_LD + j * 4, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    __syncwarp();
                    const int rr = lane >> 3, c4 = (lane & 7) * 4;
                    float* cbase = g.C + (size_t)z * g.sC + (size_t)(m0 + q * 32) * g.ldc + n + c4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rr;
                        if (m0 + q * 32 + row < g.M) st4(cbase + (size_t)row * g.ldc, ld4(stg + row * TC_EPI_LD + c4));
                    }
                    __syncwarp();
                } else if (m < g.M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n + j < g.N) crow[n + j] = __uint_as_float(v[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free + ab);      // local (CTA-scope release): the peer's relay warp forwards it to the leader
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // nobody leaves (or frees tensor memory) while the pair still works on shared state
    if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------ Q K^T with the row softmax in the epilogue (S never reaches HBM)
// A CTA pair owns a 256-query row block and walks its N / 256 key tiles TWICE:
//   pass A  one product (hi * hi, lo planes not fetched): every epilogue thread owns one query row (TMEM lane = row) and keeps the running maximum of
//           its row in a register.  Any shift within a few units of the true maximum yields the exact softmax, so bf16 operands suffice here.
//   pass B  the 3-product split; epilogue e = 2^((s - m) log2 e) (masked keys: s = -10000, axial_attention.py:81-82), e -> bf16 hi / lo straight to
//           the P planes (UNNORMALISED), row sums of e accumulated in the same register and written once per row: the P V epilogue divides by them.
// Accumulators stay double-buffered across all 2 N / 256 tiles, so both epilogues hide behind the next tile's MMAs.  Replaces the S round trip
// (4 B written + 4 B read per logit) and the separate softmax pass.  exp arguments are clamped at +80 (pass A's maximum is approximate).
constexpr int G2S_EPI_PITCH = 80;                           // bytes per staged row of 32 bf16 (64 B + 16: conflict-free 16-byte accesses)
constexpr int G2S_EPI_WARP = 2 * 32 * G2S_EPI_PITCH;        // hi + lo tile of one epilogue warp
constexpr int G2S_SMEM = G2_NSTG * G2_STAGE + 1024 + 256 + 4 * G2S_EPI_WARP;
struct TcSoftArgs {
    __nv_bfloat16 *Ph, *Pl;      // [Z][M][N] unnormalised probabilities (hi / lo planes)
    float* rowsum;               // [Z][M]
    const uint8_t* mask;         // [Z / heads][N] or null
    int heads;
};
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_tc_gemm2s(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
            const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcGemmArgs g, const TcSoftArgs sa) {
    constexpr int BN = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_align1024(smem_raw);
    // the ring is cut into 6 half slots of 32 KB ([A plane 16 KB][B half-tile plane 16 KB]), each with its own barrier pair: a one-product chunk
    // takes one half slot (hi planes), a 3-product chunk two (hi planes, then lo planes).  Position p in the ring -> slot p % 6, use p / 6.
    constexpr int NH = 2 * G2_NSTG, HSLOT = G2_STAGE / 2;
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + G2_NSTG * G2_STAGE);
    uint64_t* empty = full + NH;
    uint64_t* acc_full = empty + NH;           // [2]
    uint64_t* acc_free = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
    uint8_t* epi = tiles + G2_NSTG * G2_STAGE + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int total_chunks = (g.K + TC_BK - 1) / TC_BK;
    const int nt = g.N / BN, mt2 = g.M / (2 * TC_BM);       // host guarantees N % 256 == 0, M % 256 == 0
    const int n_items = g.Z * mt2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_full + 1, 1);
        mbar_init(acc_free, rank == 0 ? 5 : 4); mbar_init(acc_free + 1, rank == 0 ? 5 : 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int pos = 0;
        for (int it = cid; it < n_items; it += ncl) {
            const int z = it / mt2;
            const int m0 = (it - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            for (int t = 0; t < 2 * nt; ++t) {
                const bool three = t >= nt;
                const int nb0 = (three ? t - nt : t) * BN + (int)rank * (BN / 2);
                for (int kc = 0; kc < total_chunks; ++kc) {
                    for (int h = 0; h < (three ? 2 : 1); ++h, ++pos) {
                        const int s = pos % NH;
                        if (pos >= NH) mbar_wait(&empty[s], ((pos / NH) - 1) & 1);
                        uint8_t* st = tiles + s * HSLOT;
                        if (elect_one()) {
                            const uint32_t fb = mapa_u32(smem_u32(&full[s]), 0);
                            if (rank == 0) mbar_expect_tx(&full[s], 2 * HSLOT);
                            tma_load_3d_2sm(st, h ? &mapAl : &mapAh, fb, kc * TC_BK, m0, z);
                            tma_load_3d_2sm(st + TC_PLANE_BYTES, h ? &mapBl : &mapBh, fb, kc * TC_BK, nb0, z);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_bf16(2 * TC_BM, BN);
            int pos = 0, ti = 0;
            for (int it = cid; it < n_items; it += ncl) {
                for (int t = 0; t < 2 * nt; ++t, ++ti) {
                    const bool three = t >= nt;
                    const int ab = ti & 1;
                    const uint32_t t_acc = tmem_base + ab * BN;
                    if (ti >= 2) { mbar_wait(acc_free + ab, ((ti >> 1) - 1) & 1); tc_fence_after(); }
                    for (int kc = 0; kc < total_chunks; ++kc) {
                        const int s = pos % NH, s2 = (pos + 1) % NH;
                        mbar_wait(&full[s], (pos / NH) & 1);
                        if (three) mbar_wait(&full[s2], ((pos + 1) / NH) & 1);
                        pos += three ? 2 : 1;
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(tiles + s * HSLOT), b_hi = a_hi + TC_PLANE_BYTES;
                        const uint32_t a_lo = smem_u32(tiles + s2 * HSLOT), b_lo = a_lo + TC_PLANE_BYTES;
                        const int kvalid = min(TC_BK, g.K - kc * TC_BK);
                        const int ksteps = (kvalid + 15) / 16;
                        if (elect_one()) {
                            const uint32_t dah = umma_desc_lo(a_hi), dal = umma_desc_lo(a_lo), dbh = umma_desc_lo(b_hi), dbl = umma_desc_lo(b_lo);
                            if (three) {
                                if (kc == 0) umma_ss2<false>(t_acc, dal, dbh, idesc); else umma_ss2<true>(t_acc, dal, dbh, idesc);   // small terms first
                                umma_ss2<true>(t_acc, dah, dbl, idesc);
                                umma_ss2<true>(t_acc, dah, dbh, idesc);
#pragma unroll
                                for (int k = 1; k < 4; ++k) {
                                    if (k < ksteps) {
                                        umma_ss2<true>(t_acc, dal + k * 2, dbh + k * 2, idesc);
                                        umma_ss2<true>(t_acc, dah + k * 2, dbl + k * 2, idesc);
                                        umma_ss2<true>(t_acc, dah + k * 2, dbh + k * 2, idesc);
                                    }
                                }
                            } else {
                                if (kc == 0) umma_ss2<false>(t_acc, dah, dbh, idesc); else umma_ss2<true>(t_acc, dah, dbh, idesc);
#pragma unroll
                                for (int k = 1; k < 4; ++k)
                                    if (k < ksteps) umma_ss2<true>(t_acc, dah + k * 2, dbh + k * 2, idesc);
                            }
                            umma_commit2_mc(&empty[s], 3);
                            if (three) umma_commit2_mc(&empty[s2], 3);
                            if (kc == total_chunks - 1) umma_commit2_mc(acc_full + ab, 3);
                        }
                        __syncwarp();
                    }
                }
            }
        } else {      // relay (see k_tc_gemm2)
            const uint32_t free0 = mapa_u32(smem_u32(acc_free), 0);
            int ti = 0;
            for (int it = cid; it < n_items; it += ncl)
                for (int t = 0; t < 2 * nt; ++t, ++ti) {
                    const int ab = ti & 1;
                    mbar_wait(acc_free + ab, (ti >> 1) & 1);
                    if (elect_one()) mbar_arrive_cluster(free0 + ab * 8);
                    __syncwarp();
                }
        }
    } else {
        const int q = warp & 3;
        uint8_t* stg_h = epi + q * G2S_EPI_WARP;
        uint8_t* stg_l = stg_h + 32 * G2S_EPI_PITCH;
        constexpr float L2E = 1.4426950408889634f;
        int ti = 0;
        for (int it = cid; it < n_items; it += ncl) {
            const int z = it / mt2;
            const int m0 = (it - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            const int m = m0 + q * 32 + lane;                         // this thread's query row
            // padding mask of this tree as bits: lane l of the warp holds the 32 keys of column blocks l, l + 32, ... (N <= 4096), fetched once per
            // row block; a block's word reaches all lanes by shuffle (a global load per column block sat exposed in the epilogue: 15 % of the samples)
            uint32_t mbits[4] = {0u, 0u, 0u, 0u};
            if (sa.mask) {
                const uint8_t* mk = sa.mask + (size_t)(z / sa.heads) * g.N;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int c0 = (w * 32 + lane) * 32;
                    if (c0 < g.N) {
                        const uint4 k0 = *reinterpret_cast<const uint4*>(mk + c0), k1 = *reinterpret_cast<const uint4*>(mk + c0 + 16);
                        const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
                        uint32_t bits = 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) bits |= ((kw[j >> 2] >> (8 * (j & 3))) & 0xffu) ? (1u << j) : 0u;
                        mbits[w] = bits;
                    }
                }
            }
            float rmax = -INFINITY, rsum = 0.f, ms = 0.f;
            for (int t = 0; t < 2 * nt; ++t, ++ti) {
                const bool three = t >= nt;
                const int n0 = (three ? t - nt : t) * BN;
                if (t == nt) ms = -rmax * L2E;
                const int ab = ti & 1;
                mbar_wait(acc_full + ab, (ti >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int cb = 0; cb < BN / 32; ++cb) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + cb * 32, v);
                    const int n = n0 + cb * 32;
                    {
                        const int blk = n >> 5;
                        const uint32_t src = blk < 32 ? mbits[0] : (blk < 64 ? mbits[1] : (blk < 96 ? mbits[2] : mbits[3]));
                        const uint32_t bits = __shfl_sync(0xffffffffu, src, blk & 31);
                        if (bits) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if ((bits >> j) & 1u) v[j] = __float_as_uint(-10000.0f);
                        }
                    }
                    if (!three) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) rmax = fmaxf(rmax, __uint_as_float(v[j]));
                    } else {
                        uint32_t hh[16], ll[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float e0 = ex2_approx(fminf(fmaf(__uint_as_float(v[j]), L2E, ms), 115.0f));
                            const float e1 = ex2_approx(fminf(fmaf(__uint_as_float(v[j + 1]), L2E, ms), 115.0f));
                            rsum += e0 + e1;
                            split2(e0, e1, hh[j >> 1], ll[j >> 1]);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            *reinterpret_cast<uint4*>(stg_h + lane * G2S_EPI_PITCH + c * 16) = make_uint4(hh[4 * c], hh[4 * c + 1], hh[4 * c + 2], hh[4 * c + 3]);
                            *reinterpret_cast<uint4*>(stg_l + lane * G2S_EPI_PITCH + c * 16) = make_uint4(ll[4 * c], ll[4 * c + 1], ll[4 * c + 2], ll[4 * c + 3]);
                        }
                        __syncwarp();
                        // 8 rows x 64 contiguous bytes per store instruction and plane
                        const int rr = lane >> 2, c16 = (lane & 3) * 16;
                        const size_t gbase = ((size_t)z * g.M + (m0 + q * 32)) * g.N + n;
#pragma unroll
                        for (int i8 = 0; i8 < 4; ++i8) {
                            const int row = i8 * 8 + rr;
                            const size_t go = (gbase + (size_t)row * g.N) * 2 + c16;
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(sa.Ph) + go) = *reinterpret_cast<const uint4*>(stg_h + row * G2S_EPI_PITCH + c16);
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(sa.Pl) + go) = *reinterpret_cast<const uint4*>(stg_l + row * G2S_EPI_PITCH + c16);
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_free + ab);      // local (CTA-scope release): the peer's relay warp forwards it to the leader
            }
            sa.rowsum[(size_t)z * g.M + m] = rsum;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------ CTA-pair variant for ctx = P V with 256 < N <= 512: ONE pass over A
// With two N tiles per row block, P (the A operand, streamed from DRAM: 2 x 4.3 GB per 128 trees and layer) was fetched twice.  Here a
// CTA pair owns a 256-row block and ALL N columns: per k-step two UMMAs, N = 256 into accumulator columns [0, 256) and
// N2 = 16 ceil((N - 256) / 16) into [256, 256 + N2), so P is read once.  The accumulator is not double-buffered (TMEM: 256 + N2 columns);
// the TMA ring keeps filling during the epilogue.  B (MN-major) per CTA and plane: four 64-column blocks - blocks 0, 1 = this CTA's half
// of the first UMMA's columns, blocks 2, 3 = its half of the second's.  96 KB stages, 2-deep ring.
constexpr int G2W_STAGE = 6 * TC_PLANE_BYTES;     // A_hi, A_lo (16 KB each), B_hi, B_lo (32 KB each)
constexpr int G2W_NSTG = 2;
// The accumulator is not double-buffered, so the epilogue is serial with the MMAs: EIGHT epilogue warps (two per TMEM lane quarter, alternating
// 32-column blocks) drain it; their staging tiles are unpadded [32 rows][128 B] with the 16-byte chunk index XOR-swizzled by the row.
constexpr int G2W_THREADS = 320;
constexpr int G2W_EPI_BYTES = 8 * 32 * 128;
constexpr int G2W_SMEM = G2W_NSTG * G2W_STAGE + 1024 + 256 + G2W_EPI_BYTES;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2W_THREADS, 1)
k_tc_gemm2w(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
            const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcGemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_align1024(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + G2W_NSTG * G2W_STAGE);
    uint64_t* empty = full + G2W_NSTG;
    uint64_t* acc_full = empty + G2W_NSTG;     // [1]
    uint64_t* acc_free = acc_full + 1;         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);
    float* epi = reinterpret_cast<float*>(tiles + G2W_NSTG * G2W_STAGE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int total_chunks = (g.K + TC_BK - 1) / TC_BK;
    const int mt2 = (g.M + 2 * TC_BM - 1) / (2 * TC_BM);
    const int n_tiles = g.Z * mt2;
    const int n2 = (g.N - 256 + 15) & ~15;       // columns of the second UMMA (16 .. 256)
    const int h2 = n2 >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G2W_NSTG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_free, rank == 0 ? 9 : 8);      // eight local epilogue warps (+ the peer's relay in the leader)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int gc = 0;
        for (int t = cid; t < n_tiles; t += ncl) {
            const int z = t / mt2;
            const int m0 = (t - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            const int c1 = (int)rank * 128, c2 = 256 + (int)rank * h2;     // first column of this CTA's half of each UMMA
            for (int kc = 0; kc < total_chunks; ++kc, ++gc) {
                const int s = gc % G2W_NSTG;
                if (gc >= G2W_NSTG) mbar_wait(&empty[s], ((gc / G2W_NSTG) - 1) & 1);
                uint8_t* st = tiles + s * G2W_STAGE;
                if (elect_one()) {
                    const uint32_t fb = mapa_u32(smem_u32(&full[s]), 0);
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * G2W_STAGE);
                    tma_load_3d_2sm(st, &mapAh, fb, kc * TC_BK, m0, z);
                    tma_load_3d_2sm(st + TC_PLANE_BYTES, &mapAl, fb, kc * TC_BK, m0, z);
                    uint8_t *bh = st + 2 * TC_PLANE_BYTES, *bl = st + 4 * TC_PLANE_BYTES;
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        tma_load_3d_2sm(bh + nb * 8192, &mapBh, fb, c1 + nb * 64, kc * TC_BK, z);
                        tma_load_3d_2sm(bl + nb * 8192, &mapBl, fb, c1 + nb * 64, kc * TC_BK, z);
                        tma_load_3d_2sm(bh + (2 + nb) * 8192, &mapBh, fb, c2 + nb * 64, kc * TC_BK, z);
                        tma_load_3d_2sm(bl + (2 + nb) * 8192, &mapBl, fb, c2 + nb * 64, kc * TC_BK, z);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t id1 = umma_idesc_bf16(2 * TC_BM, 256) | (1u << 16), id2 = umma_idesc_bf16(2 * TC_BM, n2) | (1u << 16);
            int gc = 0, ti = 0;
            for (int t = cid; t < n_tiles; t += ncl, ++ti) {
                if (ti >= 1) { mbar_wait(acc_free, (ti - 1) & 1); tc_fence_after(); }
                for (int kc = 0; kc < total_chunks; ++kc, ++gc) {
                    const int s = gc % G2W_NSTG;
                    mbar_wait(&full[s], (gc / G2W_NSTG) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(tiles + s * G2W_STAGE), a_lo = a_hi + TC_PLANE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * TC_PLANE_BYTES, b_lo = a_hi + 4 * TC_PLANE_BYTES;
                    const int kvalid = min(TC_BK, g.K - kc * TC_BK);
                    const int ksteps = (kvalid + 15) / 16;
                    if (elect_one()) {
                        const uint32_t dah = umma_desc_lo(a_hi), dal = umma_desc_lo(a_lo);
                        const uint32_t dbh = umma_desc_lo(b_hi, 8192), dbl = umma_desc_lo(b_lo, 8192);
                        const uint32_t dbh2 = umma_desc_lo(b_hi + 16384, 8192), dbl2 = umma_desc_lo(b_lo + 16384, 8192);
                        const uint32_t t1 = tmem_base, t2 = tmem_base + 256;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < ksteps) {
                                if (kc == 0 && k == 0) { umma_ss2<false>(t1, dal, dbh, id1); umma_ss2<false>(t2, dal, dbh2, id2); }
                                else { umma_ss2<true>(t1, dal + k * 2, dbh + k * 128, id1); umma_ss2<true>(t2, dal + k * 2, dbh2 + k * 128, id2); }
                                umma_ss2<true>(t1, dah + k * 2, dbl + k * 128, id1);
                                umma_ss2<true>(t2, dah + k * 2, dbl2 + k * 128, id2);
                                umma_ss2<true>(t1, dah + k * 2, dbh + k * 128, id1);
                                umma_ss2<true>(t2, dah + k * 2, dbh2 + k * 128, id2);
                            }
                        }
                        umma_commit2_mc(&empty[s], 3);
                        if (kc == total_chunks - 1) umma_commit2_mc(acc_full, 3);
                    }
                    __syncwarp();
                }
            }
        } else {      // relay (see k_tc_gemm2)
            const uint32_t free0 = mapa_u32(smem_u32(acc_free), 0);
            int ti = 0;
            for (int t = cid; t < n_tiles; t += ncl, ++ti) {
                mbar_wait(acc_free, ti & 1);
                if (elect_one()) mbar_arrive_cluster(free0);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        int ti = 0;
        for (int t = cid; t < n_tiles; t += ncl, ++ti) {
            const int z = t / mt2;
            const int m0 = (t - z * mt2) * 2 * TC_BM + (int)rank * TC_BM;
            mbar_wait(acc_full, ti & 1);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            float* crow = g.C + (size_t)z * g.sC + (size_t)m * g.ldc;
            const int eh = (warp - 2) >> 2;                          // which of the quarter's two warps: column blocks eh, eh + 2, ...
            float* stg = epi + (warp - 2) * (32 * 32);
            const bool vec_ok = (g.ldc & 3) == 0 && (g.sC & 3) == 0;
#pragma unroll 1
            for (int cb = eh; cb < 16; cb += 2) {
                const int n = cb * 32;
                if (n >= g.N) break;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cb * 32, v);
                if (g.row_div) {
                    const float rs = m < g.M ? 1.0f / g.row_div[(size_t)z * g.M + m] : 1.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * rs);
                }
                if (n + 31 < g.N && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4(stg + lane * 32 + ((j ^ (lane & 7)) << 2), make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    __syncwarp();
                    const int rr = lane >> 3, c4 = (lane & 7) * 4;
                    float* cbase = g.C + (size_t)z * g.sC + (size_t)(m0 + q * 32) * g.ldc + n + c4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rr;
                        if (m0 + q * 32 + row < g.M) st4(cbase + (size_t)row * g.ldc, ld4(stg + row * 32 + ((((lane & 7) ^ (row & 7))) << 2)));
                    }
                    __syncwarp();
                } else if (m < g.M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n + j < g.N) crow[n + j] = __uint_as_float(v[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

// fp32 -> (hi, lo) bf16 planes: hi = bf16(x), lo = bf16(x - hi)
__global__ void k_split_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float v = x[i];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = NNJ_LO_BF16(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
}


// ------------------------------------------------------------------ operand staging helpers shared with nnj_score_tc.cu
// Unit-test kernel for the two operand forms the fused pair-score kernel relies on:
//   A [128 x 64] K-major, written into SWIZZLE_128B shared memory by threads (not TMA), and
//   B [64 (K) x 64 (N)] MN-major (N contiguous, one 128-byte row per k), also thread-written.
// D [128 x 64] = (A_hi + A_lo)(B_hi + B_lo) with the 3-product split.  One CTA, 128 threads.
template <int BN>
__global__ void __launch_bounds__(128, 1) k_tc_unit(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Dm) {
    // B [64 (K) x BN (N)] MN-major: BN/64 column blocks of [64 rows x 128 B], 8 KB apart (LBO) within each plane
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_align1024(smem_raw);
    uint8_t *a_hi = base, *a_lo = base + 16384, *b_hi = base + 32768, *b_lo = base + 32768 + 16384;
    uint64_t* done = reinterpret_cast<uint64_t*>(base + 65536);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {   // A row `tid`: 64 values -> 8 chunks of 8 bf16, chunk j stored at position j ^ (row & 7)
        const float* ar = A + (size_t)tid * 64;
        for (int j = 0; j < 8; ++j) {
            __align__(16) __nv_bfloat16 h[8], l[8];
            for (int e = 0; e < 8; ++e) { float v = ar[j * 8 + e]; h[e] = __float2bfloat16_rn(v); l[e] = NNJ_LO_BF16(__float2bfloat16_rn(v - __bfloat162float(h[e]))); }
            const int off = tid * 128 + ((j ^ (tid & 7)) << 4);
            *reinterpret_cast<uint4*>(a_hi + off) = *reinterpret_cast<uint4*>(h);
            *reinterpret_cast<uint4*>(a_lo + off) = *reinterpret_cast<uint4*>(l);
        }
        if (tid < 64) {   // B row k = tid: BN n-values
            const float* br = Bm + (size_t)tid * BN;
            for (int j = 0; j < BN / 8; ++j) {
                __align__(16) __nv_bfloat16 h[8], l[8];
                for (int e = 0; e < 8; ++e) { float v = br[j * 8 + e]; h[e] = __float2bfloat16_rn(v); l[e] = NNJ_LO_BF16(__float2bfloat16_rn(v - __bfloat162float(h[e]))); }
                const int off = (j >> 3) * 8192 + tid * 128 + (((j & 7) ^ (tid & 7)) << 4);
                *reinterpret_cast<uint4*>(b_hi + off) = *reinterpret_cast<uint4*>(h);
                *reinterpret_cast<uint4*>(b_lo + off) = *reinterpret_cast<uint4*>(l);
            }
        }
    }
    fence_async_smem();   // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, BN) | (1u << 16);   // B is MN-major
        if (elect_one()) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t dah = umma_desc_k128(smem_u32(a_hi) + k * 32), dal = umma_desc_k128(smem_u32(a_lo) + k * 32);
                const uint64_t dbh = umma_desc_lbo(smem_u32(b_hi) + k * 2048, 8192), dbl = umma_desc_lbo(smem_u32(b_lo) + k * 2048, 8192);
                umma_bf16(tmem_base, dal, dbh, idesc, k ? 1u : 0u);
                umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                umma_bf16(tmem_base, dah, dbh, idesc, 1u);
            }
            umma_commit(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
        for (int j = 0; j < 32; ++j) Dm[(size_t)tid * BN + cb * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN)); }
}


// Unit-test kernel for the A-from-TMEM operand form: A [128 x 64] is split into bf16 hi / lo by its row's thread and stored into
// tensor memory with tcgen05.st (row = TMEM lane, two bf16 per 32-bit column, K ascending), W [64 n x 64 k] is a thread-written
// K-major SWIZZLE_128B tile.  D [128 x 64] = A W^T with the 3-product split.  One CTA, 128 threads.
// TMEM columns: D [0,64) | A_hi [64,96) | A_lo [96,128).
__global__ void __launch_bounds__(128, 1) k_tc_unit_ta(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ Dm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_align1024(smem_raw);
    uint8_t *w_hi = base, *w_lo = base + 8192;
    uint64_t* done = reinterpret_cast<uint64_t*>(base + 16384);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid < 64) {
        const float* wr = W + (size_t)tid * 64;
        for (int j = 0; j < 8; ++j) {
            float v[8];
            for (int e = 0; e < 8; ++e) v[e] = wr[j * 8 + e];
            a_store8(w_hi, w_lo, tid, j, v);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    {
        const float* ar = A + (size_t)tid * 64;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 2; ++c) {           // 32 K values -> 16 columns per plane
            uint32_t hh[16], ll[16];
            for (int e = 0; e < 16; ++e) split2(ar[c * 32 + 2 * e], ar[c * 32 + 2 * e + 1], hh[e], ll[e]);
            tmem_st16(lane_base + 64 + c * 16, hh);
            tmem_st16(lane_base + 96 + c * 16, ll);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        if (elect_one()) {
            for (int k = 0; k < 4; ++k) {
                umma_bf16_ta(tmem_base, tmem_base + 96 + k * 8, umma_desc_k128(smem_u32(w_hi) + k * 32), idesc, k ? 1u : 0u);
                umma_bf16_ta(tmem_base, tmem_base + 64 + k * 8, umma_desc_k128(smem_u32(w_lo) + k * 32), idesc, 1u);
                umma_bf16_ta(tmem_base, tmem_base + 64 + k * 8, umma_desc_k128(smem_u32(w_hi) + k * 32), idesc, 1u);
            }
            umma_commit(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
        for (int j = 0; j < 32; ++j) Dm[(size_t)tid * 64 + cb * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// Micro-benchmark (not on any product path): cycles per tcgen05.mma for the shapes the kernels use.  variant bits:
//   [0..1] N: 0 = 64, 1 = 128, 2 = 256, 3 = 32      [2] A from tensor memory      [3] B MN-major      [4] two independent accumulators, alternating
// Issues 512 UMMAs (operands = whatever is in shared / tensor memory, zero-initialised) from one thread and times issue-to-retire
// with clock64.  Dm[0] = cycles per UMMA, Dm[1] = cycles the issuing thread spent in the issue loop per UMMA.
__global__ void __launch_bounds__(128, 1) k_tc_mma_bench(float* __restrict__ Dm, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_align1024(smem_raw);
    uint64_t* done = reinterpret_cast<uint64_t*>(base + 98304);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 98304 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    {   // zero the TMEM columns used as A operand
        uint32_t z[16];
        for (int e = 0; e < 16; ++e) z[e] = 0u;
        for (int c = 448; c < 512; c += 16) tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int nsel = variant & 3, N = nsel == 0 ? 64 : (nsel == 1 ? 128 : (nsel == 2 ? 256 : 32));
    const bool ta = variant & 4, bmn = variant & 8, two = variant & 16;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(128, N) | (bmn ? (1u << 16) : 0u);
        const uint32_t a_lo = umma_desc_lo(smem_u32(base)), b_lo = umma_desc_lo(smem_u32(base) + 32768, bmn ? 8192 : 0);
        const long long t0 = clock64();
        for (int i = 0; i < 128; ++i) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t td = tmem_base + ((two && (kk & 1)) ? 256 - (N > 128 ? 64 : 0) : 0);
                if (ta) umma_ts<true>(td, tmem_base + 448 + kk * 8, b_lo + (bmn ? kk * 128 : kk * 2), idesc);
                else umma_ss<true>(td, a_lo + kk * 2, b_lo + (bmn ? kk * 128 : kk * 2), idesc);
            }
        }
        const long long t1 = clock64();
        umma_commit(done);
        mbar_wait(done, 0);
        const long long t2 = clock64();
        Dm[0] = (float)(t2 - t0) / 512.0f;
        Dm[1] = (float)(t1 - t0) / 512.0f;
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int run_tc_unit(const float* A, const float* B, float* Dm, int bn, cudaStream_t st) {
    if (bn >= 2000 && bn < 2032) {   // UMMA throughput micro-benchmark
        cudaError_t e0 = cudaFuncSetAttribute(k_tc_mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e0 != cudaSuccess) return set_cuda_error(e0, __FILE__, __LINE__);
        k_tc_mma_bench<<<1, 128, 100 * 1024, st>>>(Dm, bn - 2000);
        ++g_launches;
        e0 = cudaGetLastError();
        if (e0 != cudaSuccess) return set_cuda_error(e0, __FILE__, __LINE__);
        return 0;
    }
    if (bn == 1064) {   // A from tensor memory, W [64 n][64 k] K-major
        cudaError_t e0 = cudaFuncSetAttribute(k_tc_unit_ta, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
        if (e0 != cudaSuccess) return set_cuda_error(e0, __FILE__, __LINE__);
        k_tc_unit_ta<<<1, 128, 20 * 1024, st>>>(A, B, Dm);
        ++g_launches;
        e0 = cudaGetLastError();
        if (e0 != cudaSuccess) return set_cuda_error(e0, __FILE__, __LINE__);
        return 0;
    }
    if (bn != 64 && bn != 128) return set_error(NNJ_ERR_INVALID, "tc_selftest: N must be 64, 128 or 1064 (A from TMEM)");
    cudaError_t e = cudaFuncSetAttribute(k_tc_unit<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_unit<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 68 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    if (bn == 64) k_tc_unit<64><<<1, 128, 68 * 1024, st>>>(A, B, Dm);
    else k_tc_unit<128><<<1, 128, 68 * 1024, st>>>(A, B, Dm);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// 3-D bf16 tensor [Z][rows][K] (K contiguous, row pitch ld elements, batch stride sz elements), box 64 x 128 x 1, SWIZZLE_128B.
int make_tmap_k_major(CUtensorMap* map, const void* base, int K, int rows, int Z, size_t ld, size_t sz, int box_rows) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((ld * 2) % 16 != 0 || (sz * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15))
        return set_error(NNJ_ERR_INVALID, "tensor-core path: operand rows must be 16-byte aligned (K and site count multiples of 8)");
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)Z};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)sz * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[96];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return set_error(NNJ_ERR_CUDA, msg);
    }
    return 0;
}

// 3-D bf16 tensor [Z][K][N] (N contiguous, pitch ld elements), box 64 (N) x 64 (K) x 1, SWIZZLE_128B: MN-major B operand tiles.
static int make_tmap_mn_major(CUtensorMap* map, const void* base, int N, int K, int Z, size_t ld, size_t sz) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((ld * 2) % 16 != 0 || (sz * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15))
        return set_error(NNJ_ERR_INVALID, "tensor-core path: operand rows must be 16-byte aligned");
    cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)K, (cuuint64_t)Z};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)sz * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled failed (MN-major operand)");
    return 0;
}

static int tc_sm_count() { return sm_count(); }    // of the current device (cached per device, nnj_api.cu)

// ---- CTA-pair GEMM (k_tc_gemm2).  NNJ_GEMM_2SM: bit 0 = the K-major instantiation (Q K^T), bit 1 = the MN-major one (P V), bit 2 = the
// one-pass form of P V for 256 < N <= 512 (k_tc_gemm2w); default all.
static int gemm2_mask() {
    static const int mk = [] { const char* v = getenv("NNJ_GEMM_2SM"); return v ? (atoi(v) & 7) : 7; }();
    return mk;
}
template <bool BMN>
static int gemm2_clusters() {      // CTA pairs that can be resident at once on the current device (cached per device)
    static std::atomic<int> cache[64];
    const int dev = current_device() & 63;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * tc_sm_count()); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = G2_SMEM;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, k_tc_gemm2<BMN>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = tc_sm_count() / 2; }
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// maps: A box 64 x 128 rows; B K-major box 64 x 128 rows (half an N tile) or MN-major 64 x 64
template <bool BMN>
static int launch_tc_gemm2(int cls, const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl, const TcGemmArgs& g,
                           cudaStream_t st) {
    static DevOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm2<BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        once.done();
    }
    const int n_pairs = ((g.N + 255) / 256) * ((g.M + 2 * TC_BM - 1) / (2 * TC_BM)) * g.Z;
    const int ncl = gemm2_clusters<BMN>();
    const dim3 grid(2 * (n_pairs < ncl ? n_pairs : ncl));
    prof_begin(cls, st);
    k_tc_gemm2<BMN><<<grid, TC_THREADS, G2_SMEM, st>>>(mAh, mAl, mBh, mBl, g);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}
// one pass over A for 256 < N <= 512 (NNJ_GEMM_2SM bit 2 = off switch: value 3 keeps the two-tile form)
static int launch_tc_gemm2w(int cls, const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl, const TcGemmArgs& g,
                            cudaStream_t st) {
    static DevOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm2w, cudaFuncAttributeMaxDynamicSharedMemorySize, G2W_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        once.done();
    }
    const int n_pairs = ((g.M + 2 * TC_BM - 1) / (2 * TC_BM)) * g.Z;
    const int ncl = gemm2_clusters<true>();
    const dim3 grid(2 * (n_pairs < ncl ? n_pairs : ncl));
    prof_begin(cls, st);
    k_tc_gemm2w<<<grid, G2W_THREADS, G2W_SMEM, st>>>(mAh, mAl, mBh, mBl, g);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}
static bool gemm2_ok(int bit, int M, int N, int bn, int nsplit, int products) {
    return (gemm2_mask() >> bit & 1) && bn == 256 && nsplit == 1 && products == 3 && M % (2 * TC_BM) == 0 && N >= 256;
}

// S = Q K^T with the softmax in the epilogue (k_tc_gemm2s).  NNJ_ROW_FUSED=0 keeps the three-kernel form (GEMM, softmax pass, GEMM).
bool row_qk_softmax_ok(int C, int products) {
    static const int on = [] { const char* v = getenv("NNJ_ROW_FUSED"); return v ? atoi(v) : 1; }();
    return on && products == 3 && (gemm2_mask() & 1) && C >= 256 && C % 256 == 0 && C <= 4096;
}
int launch_row_qk_softmax(int cls, const void* Qh, const void* Ql, const void* Kh, const void* Kl, void* Ph, void* Pl, float* rowsum, const uint8_t* mask,
                          int heads, int Z, int C, int K, cudaStream_t st) {
    static DevOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm2s, cudaFuncAttributeMaxDynamicSharedMemorySize, G2S_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        once.done();
    }
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int e = make_tmap_k_major(&mAh, Qh, K, C, Z, K, (size_t)C * K, TC_BM)) return e;
    if (int e = make_tmap_k_major(&mAl, Ql, K, C, Z, K, (size_t)C * K, TC_BM)) return e;
    if (int e = make_tmap_k_major(&mBh, Kh, K, C, Z, K, (size_t)C * K, 128)) return e;
    if (int e = make_tmap_k_major(&mBl, Kl, K, C, Z, K, (size_t)C * K, 128)) return e;
    TcGemmArgs g{nullptr, C, C, K, C, (size_t)C * C, Z, 1, (K + TC_BK - 1) / TC_BK, 0};
    TcSoftArgs sa{(__nv_bfloat16*)Ph, (__nv_bfloat16*)Pl, rowsum, mask, heads};
    const int n_items = (C / (2 * TC_BM)) * Z;
    const int ncl = gemm2_clusters<false>();
    const dim3 grid(2 * (n_items < ncl ? n_items : ncl));
    prof_begin(cls, st);
    k_tc_gemm2s<<<grid, TC_THREADS, G2S_SMEM, st>>>(mAh, mAl, mBh, mBl, g, sa);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

// C[z] = (Ah+Al)[z] * (Bh+Bl)[z] with B MN-major: B planes [Z][K][N] (N contiguous, pitch ldb).  128 x 128 tiles.
int launch_tc_gemm_bmn(int cls, const void* Ah, const void* Al, const void* Bh, const void* Bl, float* Cm, int Z, int M, int N, int K, size_t lda,
                       size_t sA, size_t ldb, size_t sB, int ldc, size_t sC, cudaStream_t st, int products, const float* row_div) {
    static DevOnce once;      // per device, not per process
    constexpr int SMEM128 = TC_STAGES * (4 * TC_PLANE_BYTES) + 1024 + 256 + TC_EPI_BYTES;
    constexpr int SMEM256 = 2 * (6 * TC_PLANE_BYTES) + 1024 + 256 + TC_EPI_BYTES;
    const bool one = products == 1;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<256, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM256);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM256);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        once.done();
    }
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int e = make_tmap_k_major(&mAh, Ah, K, M, Z, lda, sA, TC_BM)) return e;
    if (int e = make_tmap_k_major(&mAl, Al, K, M, Z, lda, sA, TC_BM)) return e;
    if (int e = make_tmap_mn_major(&mBh, Bh, N, K, Z, ldb, sB)) return e;
    if (int e = make_tmap_mn_major(&mBl, Bl, N, K, Z, ldb, sB)) return e;
    TcGemmArgs g{Cm, M, N, K, ldc, sC, Z, 1, (K + TC_BK - 1) / TC_BK, 0};
    g.row_div = row_div;
    const int bn = N > 128 ? 256 : 128;         // 128 x 256 tiles: a UMMA with N = 256 runs at 75 % of the tensor peak, N = 128 at 60 %
    if (gemm2_ok(1, M, N, bn, 1, products) && (gemm2_mask() & 4) && N > 256 && N <= 512) return launch_tc_gemm2w(cls, mAh, mAl, mBh, mBl, g, st);
    if (gemm2_ok(1, M, N, bn, 1, products)) return launch_tc_gemm2<true>(cls, mAh, mAl, mBh, mBl, g, st);
    const int n_tiles = ((N + bn - 1) / bn) * ((M + TC_BM - 1) / TC_BM) * Z;
    const dim3 grid(n_tiles < tc_sm_count() ? n_tiles : tc_sm_count());
    prof_begin(cls, st);
    if (bn == 256) { if (one) k_tc_gemm<256, true, true><<<grid, TC_THREADS, SMEM256, st>>>(mAh, mAl, mBh, mBl, g); else k_tc_gemm<256, true, false><<<grid, TC_THREADS, SMEM256, st>>>(mAh, mAl, mBh, mBl, g); }
    else { if (one) k_tc_gemm<128, true, true><<<grid, TC_THREADS, SMEM128, st>>>(mAh, mAl, mBh, mBl, g); else k_tc_gemm<128, true, false><<<grid, TC_THREADS, SMEM128, st>>>(mAh, mAl, mBh, mBl, g); }
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

// C[z] = (Ah+Al)[z] * (Bh+Bl)[z]^T.  A planes [Z][M][K] (pitch lda, batch stride sA), B planes [Z][N][K].
// bn = 128 or 64 (N tile); nsplit > 1 splits K into nsplit ranges of chunks_per_split 64-element chunks, split s writing
// its partial tile at C + s*split_stride.
int launch_tc_gemm_ex(int cls, const void* Ah, const void* Al, const void* Bh, const void* Bl, float* Cm, int Z, int M, int N, int K, size_t lda,
                      size_t sA, size_t ldb, size_t sB, int ldc, size_t sC, int bn, int nsplit, int chunks_per_split, size_t split_stride,
                      cudaStream_t st, int products) {
    static DevOnce once;      // per device, not per process
    constexpr int SMEM128 = TC_STAGES * (4 * TC_PLANE_BYTES) + 1024 + 256 + TC_EPI_BYTES;
    constexpr int SMEM64 = TC_STAGES * (3 * TC_PLANE_BYTES) + 1024 + 256 + TC_EPI_BYTES;
    constexpr int SMEM256 = 2 * (6 * TC_PLANE_BYTES) + 1024 + 256 + TC_EPI_BYTES;
    const bool one = products == 1;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM64);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM256);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM128);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_gemm<256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM256);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        once.done();
    }
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int e = make_tmap_k_major(&mAh, Ah, K, M, Z, lda, sA, TC_BM)) return e;
    if (int e = make_tmap_k_major(&mAl, Al, K, M, Z, lda, sA, TC_BM)) return e;
    if (int e = make_tmap_k_major(&mBh, Bh, K, N, Z, ldb, sB, bn)) return e;
    if (int e = make_tmap_k_major(&mBl, Bl, K, N, Z, ldb, sB, bn)) return e;
    TcGemmArgs g{Cm, M, N, K, ldc, sC, Z, nsplit, chunks_per_split, split_stride};
    if (gemm2_ok(0, M, N, bn, nsplit, products) && N % 256 == 0) {
        CUtensorMap mBh2, mBl2;      // half an N tile per CTA of the pair
        if (int e = make_tmap_k_major(&mBh2, Bh, K, N, Z, ldb, sB, 128)) return e;
        if (int e = make_tmap_k_major(&mBl2, Bl, K, N, Z, ldb, sB, 128)) return e;
        return launch_tc_gemm2<false>(cls, mAh, mAl, mBh2, mBl2, g, st);
    }
    const int n_tiles = ((N + bn - 1) / bn) * nsplit * ((M + TC_BM - 1) / TC_BM) * Z;
    const dim3 grid(n_tiles < tc_sm_count() ? n_tiles : tc_sm_count());
    prof_begin(cls, st);
    if (one && bn == 64) return set_error(NNJ_ERR_INVALID, "tc_gemm: the one-product form has no 64-column tile");
    if (bn == 256) { if (one) k_tc_gemm<256, false, true><<<grid, TC_THREADS, SMEM256, st>>>(mAh, mAl, mBh, mBl, g); else k_tc_gemm<256, false, false><<<grid, TC_THREADS, SMEM256, st>>>(mAh, mAl, mBh, mBl, g); }
    else if (bn == 128) { if (one) k_tc_gemm<128, false, true><<<grid, TC_THREADS, SMEM128, st>>>(mAh, mAl, mBh, mBl, g); else k_tc_gemm<128, false, false><<<grid, TC_THREADS, SMEM128, st>>>(mAh, mAl, mBh, mBl, g); }
    else k_tc_gemm<64, false, false><<<grid, TC_THREADS, SMEM64, st>>>(mAh, mAl, mBh, mBl, g);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

int launch_tc_gemm(int cls, const void* Ah, const void* Al, const void* Bh, const void* Bl, float* Cm, int Z, int M, int N, int K, size_t lda,
                   size_t sA, size_t ldb, size_t sB, int ldc, size_t sC, cudaStream_t st, int products) {
    const int bn = (N >= 256 && N % 256 == 0) ? 256 : 128;     // 128 x 256 tiles when they divide N (row-attention logits)
    return launch_tc_gemm_ex(cls, Ah, Al, Bh, Bl, Cm, Z, M, N, K, lda, sA, ldb, sB, ldc, sC, bn, 1, (K + TC_BK - 1) / TC_BK, 0, st, products);
}

// Stand-alone building block (also the unit-test entry): fp32 A [Z][M][K], B [Z][N][K] -> C [Z][M][N].
// ws must hold 2*(Z*M*K + Z*N*K) bf16 values (+256 B).
int run_gemm_split_bf16(const float* A, const float* B, float* Cm, int Z, int M, int N, int K, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t nA = (size_t)Z * M * K, nB = (size_t)Z * N * K;
    if (K % 8 != 0) return set_error(NNJ_ERR_INVALID, "gemm_split_bf16: K must be a multiple of 8");
    if (ws_bytes < 2 * (nA + nB) * 2 + 1024) return set_error(NNJ_ERR_WORKSPACE, "gemm_split_bf16: workspace too small");
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    auto take = [&](size_t elems) { void* r = p; p += (elems * 2 + 255) & ~(size_t)255; return (__nv_bfloat16*)r; };
    __nv_bfloat16 *Ah = take(nA), *Al = take(nA), *Bh = take(nB), *Bl = take(nB);
    prof_begin(KC_MISC, st);
    k_split_bf16<<<1184, 256, 0, st>>>(A, Ah, Al, nA);
    ++g_launches;
    prof_end(st);
    prof_begin(KC_MISC, st);
    k_split_bf16<<<1184, 256, 0, st>>>(B, Bh, Bl, nB);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return launch_tc_gemm(KC_MISC, Ah, Al, Bh, Bl, Cm, Z, M, N, K, K, (size_t)M * K, K, (size_t)N * K, N, (size_t)M * N, st);
}

}  // namespace nnj
