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
// Every mbarrier wait is bounded (trap after ~2 s) so that a protocol bug surfaces as an error, not a hang.
#include <cuda.h>
#include <cuda_bf16.h>

#include "nnj_internal.h"

namespace nnj {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_PLANE_BYTES = TC_BM * TC_BK * 2;          // 16 KB: 128 rows x 128 B
constexpr int TC_STAGE_BYTES = 4 * TC_PLANE_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 1.9 GHz: protocol error, fail loudly instead of hanging
    }
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024 B (8 rows x 128 B),
// [46,48) version = 1 (sm_100), [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// grid (N tiles, M tiles, Z).  C row-major with leading dimension ldc, batch stride sC (elements).
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
          const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, float* __restrict__ Cm, int M, int N,
          int K, int ldc, size_t sC) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* accum_done = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM, z = blockIdx.z;
    const int nchunks = (K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 128 fp32 columns for the 128 x 128 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = kc % TC_STAGES;
                if (kc >= TC_STAGES) mbar_wait(&empty[s], ((kc / TC_STAGES) - 1) & 1);
                uint8_t* st = tiles + s * TC_STAGE_BYTES;
                mbar_expect_tx(&full[s], TC_STAGE_BYTES);
                tma_load_3d(st, &mapAh, &full[s], kc * TC_BK, m0, z);
                tma_load_3d(st + TC_PLANE_BYTES, &mapAl, &full[s], kc * TC_BK, m0, z);
                tma_load_3d(st + 2 * TC_PLANE_BYTES, &mapBh, &full[s], kc * TC_BK, n0, z);
                tma_load_3d(st + 3 * TC_PLANE_BYTES, &mapBl, &full[s], kc * TC_BK, n0, z);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = kc % TC_STAGES;
                mbar_wait(&full[s], (kc / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(tiles + s * TC_STAGE_BYTES), a_lo = a_hi + TC_PLANE_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_PLANE_BYTES, b_lo = a_hi + 3 * TC_PLANE_BYTES;
                const int kvalid = min(TC_BK, K - kc * TC_BK);
                const int ksteps = (kvalid + 15) / 16;
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t ko = k * 32;   // 16 bf16 = 32 B inside the 128 B swizzle row
                    const uint64_t dah = umma_desc_k128(a_hi + ko), dal = umma_desc_k128(a_lo + ko);
                    const uint64_t dbh = umma_desc_k128(b_hi + ko), dbl = umma_desc_k128(b_lo + ko);
                    umma_bf16(tmem_base, dal, dbh, idesc, (kc | k) ? 1u : 0u);   // small terms first
                    umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                    umma_bf16(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
            }
            umma_commit(accum_done);               // accumulator complete
        }
    } else {
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        mbar_wait(accum_done, 0);
        tc_fence_after();
        const int m = m0 + q * 32 + lane;
        float* crow = Cm + (size_t)z * sC + (size_t)m * ldc;
#pragma unroll 1
        for (int cb = 0; cb < TC_BN / 32; ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cb * 32, v);
            const int n = n0 + cb * 32;
            if (m < M) {
                if (n + 31 < N && (ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4(crow + n + j * 4, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n + j < N) crow[n + j] = __uint_as_float(v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

// fp32 -> (hi, lo) bf16 planes: hi = bf16(x), lo = bf16(x - hi)
__global__ void k_split_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float v = x[i];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
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
int make_tmap_k_major(CUtensorMap* map, const void* base, int K, int rows, int Z, size_t ld, size_t sz) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return set_error(NNJ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((ld * 2) % 16 != 0 || (sz * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15))
        return set_error(NNJ_ERR_INVALID, "tensor-core path: operand rows must be 16-byte aligned (K and site count multiples of 8)");
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)Z};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)sz * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM, 1};
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

// C[z] = (Ah+Al)[z] * (Bh+Bl)[z]^T.  A planes [Z][M][K] (pitch lda, batch stride sA), B planes [Z][N][K].
int launch_tc_gemm(int cls, const void* Ah, const void* Al, const void* Bh, const void* Bl, float* Cm, int Z, int M, int N, int K, size_t lda,
                   size_t sA, size_t ldb, size_t sB, int ldc, size_t sC, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        attr = true;
    }
    CUtensorMap mAh, mAl, mBh, mBl;
    if (int e = make_tmap_k_major(&mAh, Ah, K, M, Z, lda, sA)) return e;
    if (int e = make_tmap_k_major(&mAl, Al, K, M, Z, lda, sA)) return e;
    if (int e = make_tmap_k_major(&mBh, Bh, K, N, Z, ldb, sB)) return e;
    if (int e = make_tmap_k_major(&mBl, Bl, K, N, Z, ldb, sB)) return e;
    prof_begin(cls, st);
    k_tc_gemm<<<dim3((N + TC_BN - 1) / TC_BN, (M + TC_BM - 1) / TC_BM, Z), TC_THREADS, TC_SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, Cm, M, N, K, ldc, sC);
    ++g_launches;
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
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
