// nnj_llh.cu — tree log-likelihood under GTR+I+G4 and its maximisation over branch lengths.
//
// Replaces, for Search / branch_optimize scoring, the reference's native binding raxmlpy.compute_llh / optimize_brlen
// (RAxMLpy/cpp/raxmlpy.cpp:1749-1872 -> raxml-ng TreeInfo::loglh / optimize_branches; call sites environment.py:365-379,625-670,
// finetune_rl_search.py:401-411).  raxml-ng / libpll are git-cloned by the reference's setup.py and are not available: the
// algorithm is the published one (Felsenstein pruning; Yang's discrete gamma, class means; invariant-site mixture; Newton-Raphson
// per branch on the eigen-space sumtable, branches swept depth-first with the conditional likelihood vectors refreshed on the way),
// restated on the CPU in oracle/llh_oracle.py - parity against raxml-ng itself is UNPINNED.
//
// One CTA per tree, threads over alignment patterns, everything in fp64 (log-likelihoods of ~1e4 with gains of ~1e-3 per pass).
// Conditional likelihood vectors (CLVs) [node][pattern][4 rate classes][4 states] live in the caller's workspace: D[v] = subtree
// below v, U[v] = everything else, seen from the parent end of the branch above v.  A thread owns its patterns through the whole
// traversal, so the only block-wide steps are the 4 x 4 transition matrices of an operation (64 threads, then a barrier) and
// the reductions of the Newton iteration (fixed order: results are bit-reproducible).
// Node numbering: leaves 0..R-1, inner node R+k = k-th join (children[k]), root = the last; brlen[v] = branch above v.
#include "nnj_internal.h"

namespace nnj {

constexpr int LLH_THREADS = 512;
constexpr int LLH_MODEL = 64;      // lam 4 | U 16 | Uinv 16 | freqs 4 | class rates 4 | p_inv | alpha | flags (1 GTR rates free, 2 +G, 4 +I) | pad | 6 GTR rates | pad 10
constexpr double RATE_LO = 1e-3, RATE_HI = 1e3, ALPHA_LO = 0.02, ALPHA_HI = 100.0, PINV_HI = 0.99, GOLD = 0.3819660112501051;   // as in oracle/llh_oracle.py
constexpr double BRLEN_MIN = 1e-6, BRLEN_MAX = 100.0;
enum { OP_OPT = 0, OP_UP = 1, OP_DOWN = 2, OP_SWAPROOT = 3 };

struct LlhArgs {
    const uint8_t* tips;        // [B][R][L] 4-bit state masks
    const double* weights;      // [B][L] pattern multiplicities
    const int32_t* children;    // [B][R-1][2]
    const int32_t* ops;         // [B][n_ops][4] (optimise only)
    double* brlen;              // [B][2R-2] in / out
    const double* model;        // [B][LLH_MODEL]
    double *D, *U, *S, *inv;    // workspace: [B][R-1][L][16], [B][2R-2][L][16], [B][L][16], [B][L]
    double* llh;                // [B][2]: before, after
    int B, R, L, n_ops, max_passes, optimise;     // optimise: 0 evaluate, 1 branch lengths, 2 branch lengths + model parameters
    double eps;
    double lh_eps; int max_rounds, golden_iters;  // optimise = 2: rounds of (free parameters one at a time, branch sweeps) until a round gains < lh_eps
    double* model_out;                            // [B][LLH_MODEL] optimised model (optimise = 2)
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// P(t)[k][a][b] = sum_j U[a][j] exp(lam_j r_k t) Uinv[j][b] for two branch lengths at once (threads 0..127), then a barrier
__device__ __forceinline__ void pmat2(double* Pm, const double* mdl, double t0, double t1) {
    const int i = threadIdx.x;
    if (i < 128) {
        const int w = i >> 6, k = (i >> 4) & 3, a = (i >> 2) & 3, b = i & 3;
        const double t = w ? t1 : t0;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) s += mdl[4 + a * 4 + j] * exp(mdl[j] * mdl[40 + k] * t) * mdl[20 + j * 4 + b];
        Pm[i] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void clv_load(const LlhArgs& a, const double* Dt, const uint8_t* tips, int v, int s, double (&x)[16]) {
    if (v < a.R) {
        const int m = tips[(size_t)v * a.L + s];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) x[k * 4 + c] = (double)((m >> c) & 1);
    } else {
        const double2* p = reinterpret_cast<const double2*>(Dt + ((size_t)(v - a.R) * a.L + s) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const double2 d = p[i]; x[2 * i] = d.x; x[2 * i + 1] = d.y; }
    }
}
__device__ __forceinline__ void vec_load(const double* p, double (&x)[16]) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 d = q[i]; x[2 * i] = d.x; x[2 * i + 1] = d.y; }
}
__device__ __forceinline__ void vec_store(double* p, const double (&x)[16]) {
    double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = make_double2(x[2 * i], x[2 * i + 1]);
}
// y[k][a] = sum_b P[k][a][b] x[k][b]
__device__ __forceinline__ void pm_apply(const double* P, const double (&x)[16], double (&y)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) s = fma(P[k * 16 + c * 4 + b], x[k * 4 + b], s);
            y[k * 4 + c] = s;
        }
}

// block-wide sums of two values in a fixed order; every thread returns with the totals.  red: 2 * 32 + 2 doubles.
__device__ __forceinline__ void block_sum2(double& v0, double& v1, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[warp] = v0; red[32 + warp] = v1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < LLH_THREADS / 32; ++w) { s0 += red[w]; s1 += red[32 + w]; }
        red[64] = s0; red[65] = s1;
    }
    __syncthreads();
    v0 = red[64]; v1 = red[65];
    __syncthreads();
}

// ---- model parameters on the device (optimise = 2): one thread rebuilds the eigen-system and the class rates of its tree
// regularised lower incomplete gamma P(a, x): series / continued fraction (Lentz), as gammp below on the host
__device__ double gammp_dev(double a, double x) {
    if (x <= 0.0) return 0.0;
    const double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 2000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
        return sum * exp(-x + a * log(x) - gln);
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
}
// class i (0..3) boundary quantile of Gamma(alpha, 1) at (i + 1) / 4, then P(alpha + 1, .) - what gamma_rates() does on the host
__device__ double gamma_cdf1_dev(double alpha, int i) {
    const double target = (double)(i + 1) * 0.25, gln = lgamma(alpha);
    double lo = 0.0, hi = alpha + 1.0;
    while (gammp_dev(alpha, hi) < target) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        const double f = gammp_dev(alpha, x) - target;
        if (f < 0.0) lo = x; else hi = x;
        const double pdf = exp((alpha - 1.0) * log(x) - x - gln);
        double xn = pdf > 0.0 ? x - f / pdf : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        const bool done = fabs(xn - x) <= 1e-15 * x || hi - lo <= 1e-300;
        x = xn;
        if (done) break;
    }
    return gammp_dev(alpha + 1.0, x);
}
// mdl: raw parameters (rates [48..53], freqs [36..39], alpha [45]) -> eigenvalues [0..3], U [4..19], U^-1 [20..35], class rates [40..43].
// Threads 0 (eigen-system: cyclic Jacobi on the symmetrised rate matrix) and 1..3 (class boundaries) work, then a barrier.
__device__ void rebuild_model(double* mdl, double* scratch /* 4 doubles */, bool eigen, bool gamma) {
    const int tid = threadIdx.x;
    if (tid == 0 && eigen) {
        const double* pi = mdl + 36;
        double r[4][4] = {{0, mdl[48], mdl[49], mdl[50]}, {mdl[48], 0, mdl[51], mdl[52]}, {mdl[49], mdl[51], 0, mdl[53]}, {mdl[50], mdl[52], mdl[53], 0}};
        double Q[4][4], S[4][4], V[4][4];
        double norm = 0.0;
        for (int a = 0; a < 4; ++a) {
            double d = 0.0;
            for (int c = 0; c < 4; ++c) { Q[a][c] = r[a][c] * pi[c]; if (c != a) d += Q[a][c]; }
            Q[a][a] = -d;
            norm += pi[a] * d;
        }
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 4; ++c) { S[a][c] = sqrt(pi[a]) * Q[a][c] / (norm * sqrt(pi[c])); V[a][c] = a == c ? 1.0 : 0.0; }
        for (int a = 0; a < 4; ++a)
            for (int c = a + 1; c < 4; ++c) { const double m = 0.5 * (S[a][c] + S[c][a]); S[a][c] = m; S[c][a] = m; }
        double diag2 = 0.0;
        for (int p = 0; p < 4; ++p) diag2 += S[p][p] * S[p][p];
        for (int sweep = 0; sweep < 16; ++sweep) {           // quadratic convergence: 5-6 sweeps reach the rounding level of a 4 x 4 matrix
            double off = 0.0;
            for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += S[p][q] * S[p][q];
            if (off < 1e-31 * diag2) break;
            for (int p = 0; p < 4; ++p)
                for (int q = p + 1; q < 4; ++q) {
                    if (fabs(S[p][q]) < 1e-300) continue;
                    const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                    const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                    for (int k = 0; k < 4; ++k) { const double skp = S[k][p], skq = S[k][q]; S[k][p] = cs * skp - sn * skq; S[k][q] = sn * skp + cs * skq; }
                    for (int k = 0; k < 4; ++k) { const double spk = S[p][k], sqk = S[q][k]; S[p][k] = cs * spk - sn * sqk; S[q][k] = sn * spk + cs * sqk; }
                    for (int k = 0; k < 4; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
                }
        }
        for (int j = 0; j < 4; ++j) mdl[j] = S[j][j];
        for (int a = 0; a < 4; ++a)
            for (int j = 0; j < 4; ++j) { mdl[4 + a * 4 + j] = V[a][j] / sqrt(pi[a]); mdl[20 + j * 4 + a] = V[a][j] * sqrt(pi[a]); }
    }
    if (gamma && tid >= 1 && tid <= 3) scratch[tid] = gamma_cdf1_dev(mdl[45], tid - 1);
    __syncthreads();
    if (gamma && tid == 0) {
        scratch[0] = 0.0;
        const double c4 = 1.0;
        mdl[40] = 4.0 * (scratch[1] - 0.0); mdl[41] = 4.0 * (scratch[2] - scratch[1]); mdl[42] = 4.0 * (scratch[3] - scratch[2]); mdl[43] = 4.0 * (c4 - scratch[3]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LLH_THREADS) k_llh(const LlhArgs a) {
    extern __shared__ double sm_d[];
    double* mdl = sm_d;                 // [48]
    double* Pm = mdl + LLH_MODEL;       // [2][4][16]
    double* red = Pm + 128;             // [66]
    double* ek = red + 66;              // [3][16]: exp(rate t), rate, rate^2 of the Newton iteration
    double* t = ek + 48;                // [2R-2]
    const int b = blockIdx.x, tid = threadIdx.x, R = a.R, L = a.L;
    const uint8_t* tips = a.tips + (size_t)b * R * L;
    const double* wts = a.weights + (size_t)b * L;
    const int32_t* ch = a.children + (size_t)b * (R - 1) * 2;
    double* Dt = a.D + (size_t)b * (R - 1) * L * 16;
    double* Ut = a.U + (size_t)b * (2 * R - 2) * L * 16;
    double* St = a.S + (size_t)b * L * 16;
    double* invt = a.inv + (size_t)b * L;
    if (tid < LLH_MODEL) mdl[tid] = a.model[(size_t)b * LLH_MODEL + tid];
    for (int v = tid; v < 2 * R - 2; v += LLH_THREADS) t[v] = clampd(a.brlen[(size_t)b * (2 * R - 2) + v], 0.0, BRLEN_MAX);
    __syncthreads();
    const int c1 = ch[(R - 2) * 2], c2 = ch[(R - 2) * 2 + 1];
    if (tid == 0) { const double t0 = clampd(t[c1] + t[c2], a.optimise ? BRLEN_MIN : 0.0, BRLEN_MAX); t[c1] = t0; t[c2] = t0; }   // the root branch, stored at both ends
    __syncthreads();
    if (a.optimise)
        for (int v = tid; v < 2 * R - 2; v += LLH_THREADS) t[v] = clampd(t[v], BRLEN_MIN, BRLEN_MAX);
    __syncthreads();
    double pinv = mdl[44], wk = (1.0 - pinv) * 0.25;            // re-read after every change of the model (optimise = 2)
    // invariant-site term: pi_x where every tip of the pattern is compatible with exactly one state x
    for (int s = tid; s < L; s += LLH_THREADS) {
        int m = 15;
        for (int v = 0; v < R; ++v) m &= tips[(size_t)v * L + s];
        invt[s] = (m != 0 && (m & (m - 1)) == 0) ? mdl[36 + (31 - __clz(m))] : 0.0;      // times p_inv where it is used
    }
    // ---- subtree vectors, children before parents (join order); the virtual root itself is never needed
    auto combine = [&](int x, int y, double* dst_base /* [L][16] */, const double* xsrc_U /* non-null: take U[x] instead of D[x] */) {
        pmat2(Pm, mdl, t[x], t[y]);
        for (int s = tid; s < L; s += LLH_THREADS) {
            double vx[16], vy[16], px[16], py[16];
            if (xsrc_U) vec_load(xsrc_U + (size_t)s * 16, vx); else clv_load(a, Dt, tips, x, s, vx);
            clv_load(a, Dt, tips, y, s, vy);
            pm_apply(Pm, vx, px);
            pm_apply(Pm + 64, vy, py);
#pragma unroll
            for (int i = 0; i < 16; ++i) px[i] *= py[i];
            vec_store(dst_base + (size_t)s * 16, px);
        }
        __syncthreads();      // Pm is rewritten by the next operation
    };
    auto down_pass = [&]() { for (int k = 0; k < R - 2; ++k) combine(ch[2 * k], ch[2 * k + 1], Dt + (size_t)k * L * 16, nullptr); };
    down_pass();
    auto root_loglik = [&]() {
        pmat2(Pm, mdl, t[c1], 0.0);
        double acc = 0.0, dummy = 0.0;
        for (int s = tid; s < L; s += LLH_THREADS) {
            double v1[16], v2[16], p1[16];
            clv_load(a, Dt, tips, c1, s, v1);
            clv_load(a, Dt, tips, c2, s, v2);
            pm_apply(Pm, v1, p1);
            double l = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) l = fma(mdl[36 + c] * p1[k * 4 + c], v2[k * 4 + c], l);
            acc += wts[s] * log(wk * l + pinv * invt[s]);
        }
        block_sum2(acc, dummy, red);
        return acc;
    };
    double ll = root_loglik();
    const double ll0 = ll;
    const int32_t* ops = a.ops + (size_t)b * a.n_ops * 4;
    // branch-length sweeps from the current subtree vectors (which must match t[] and the model) until a sweep gains < eps
    auto sweeps = [&]() {
        for (int pass = 0; pass < a.max_passes; ++pass) {
            // U[c1] = D[c2]
            for (int s = tid; s < L; s += LLH_THREADS) {
                double v2[16];
                clv_load(a, Dt, tips, c2, s, v2);
                vec_store(Ut + ((size_t)c1 * L + s) * 16, v2);
            }
            for (int oi = 0; oi < a.n_ops; ++oi) {
                const int op = ops[oi * 4], v = ops[oi * 4 + 1], x = ops[oi * 4 + 2], y = ops[oi * 4 + 3];
                if (op == OP_UP) {
                    combine(x, y, Ut + (size_t)v * L * 16, Ut + (size_t)x * L * 16);
                } else if (op == OP_DOWN) {
                    combine(x, y, Dt + (size_t)(v - R) * L * 16, nullptr);
                } else if (op == OP_SWAPROOT) {
                    for (int s = tid; s < L; s += LLH_THREADS) {
                        double v1[16];
                        clv_load(a, Dt, tips, x, s, v1);
                        vec_store(Ut + ((size_t)v * L + s) * 16, v1);
                    }
                } else {
                    // ---- Newton-Raphson on the branch above v.  Sumtable S[s][k][j] = (sum_a pi_a U[a] Umat[a][j]) (sum_b Uinv[j][b] D[b])
                    for (int s = tid; s < L; s += LLH_THREADS) {
                        double u[16], d[16], sv[16];
                        vec_load(Ut + ((size_t)v * L + s) * 16, u);
                        clv_load(a, Dt, tips, v, s, d);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                double A = 0.0, Bm = 0.0;
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    A = fma(mdl[36 + c] * u[k * 4 + c], mdl[4 + c * 4 + j], A);
                                    Bm = fma(mdl[20 + j * 4 + c], d[k * 4 + c], Bm);
                                }
                                sv[k * 4 + j] = A * Bm;
                            }
                        vec_store(St + (size_t)s * 16, sv);
                    }
                    double tv = t[v], lo = BRLEN_MIN, hi = BRLEN_MAX;
                    for (int it = 0; it < 32; ++it) {
                        if (tid < 16) {
                            const double rate = mdl[tid & 3] * mdl[40 + (tid >> 2)];
                            ek[tid] = exp(rate * tv); ek[16 + tid] = rate; ek[32 + tid] = rate * rate;
                        }
                        __syncthreads();
                        double f1 = 0.0, f2 = 0.0;
                        for (int s = tid; s < L; s += LLH_THREADS) {
                            double sv[16];
                            vec_load(St + (size_t)s * 16, sv);
                            double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const double se = sv[i] * ek[i];
                                l0 += se; l1 = fma(se, ek[16 + i], l1); l2 = fma(se, ek[32 + i], l2);
                            }
                            const double L0 = wk * l0 + pinv * invt[s], g = wk * l1 / L0;
                            f1 = fma(wts[s], g, f1);
                            f2 = fma(wts[s], wk * l2 / L0 - g * g, f2);
                        }
                        block_sum2(f1, f2, red);
                        if (f1 > 0.0) lo = tv; else hi = tv;
                        double tn = f2 < 0.0 ? tv - f1 / f2 : (f1 > 0.0 ? tv * 4.0 : tv * 0.25);
                        if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
                        const bool done = fabs(tn - tv) < 1e-9 + 1e-7 * tv;
                        tv = tn;
                        if (done) break;
                    }
                    if (tid == 0) { t[v] = tv; if (v == c1) t[c2] = tv; }
                    __syncthreads();
                }
            }
            const double nl = root_loglik();
            const double gain = nl - ll;
            ll = nl;
            if (gain < a.eps) break;
        }
    };
    if (a.optimise) {
        sweeps();
        if (a.optimise == 2) {
            // ---- model parameters: rounds of (every free parameter by golden section, all else fixed; branch sweeps) until a round
            //      gains < lh_eps - the loop of oracle/llh_oracle.py optimize_all, one tree per CTA, no host round trips
            const int flags = (int)mdl[46];
            auto objective = [&](int prm, double x) {            // set parameter prm to x (log scale for rates / alpha), rebuild, evaluate
                if (tid == 0) {
                    if (prm < 5) mdl[48 + prm] = exp(x); else if (prm == 5) mdl[45] = exp(x); else mdl[44] = x;
                }
                __syncthreads();
                rebuild_model(mdl, ek, prm < 5, prm == 5);
                pinv = mdl[44]; wk = (1.0 - pinv) * 0.25;
                down_pass();
                return root_loglik();
            };
            for (int round = 0; round < a.max_rounds; ++round) {
                const double start = ll;
                for (int prm = 0; prm < 7; ++prm) {
                    if (prm < 5 ? !(flags & 1) : (prm == 5 ? !(flags & 2) : !(flags & 4))) continue;
                    double lo = prm < 5 ? log(RATE_LO) : (prm == 5 ? log(ALPHA_LO) : 0.0), hi = prm < 5 ? log(RATE_HI) : (prm == 5 ? log(ALPHA_HI) : PINV_HI);
                    double x1 = lo + GOLD * (hi - lo), x2 = hi - GOLD * (hi - lo);
                    double f1 = objective(prm, x1), f2 = objective(prm, x2);
                    for (int it = 0; it < a.golden_iters; ++it) {
                        if (f1 < f2) { lo = x1; x1 = x2; f1 = f2; x2 = hi - GOLD * (hi - lo); f2 = objective(prm, x2); }
                        else { hi = x2; x2 = x1; f2 = f1; x1 = lo + GOLD * (hi - lo); f1 = objective(prm, x1); }
                    }
                    ll = objective(prm, f1 > f2 ? x1 : x2);
                }
                sweeps();
                if (ll - start < a.lh_eps) break;
            }
            if (tid < LLH_MODEL) a.model_out[(size_t)b * LLH_MODEL + tid] = mdl[tid];
        }
        __syncthreads();
        if (tid == 0) { const double h = 0.5 * t[c1]; t[c1] = h; t[c2] = h; }     // the rooted form splits the root branch evenly
        __syncthreads();
        for (int v = tid; v < 2 * R - 2; v += LLH_THREADS) a.brlen[(size_t)b * (2 * R - 2) + v] = t[v];
    }
    if (tid == 0) { a.llh[2 * b] = ll0; a.llh[2 * b + 1] = ll; }
}

// ---- host side
static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
struct LlhLayout { size_t tipsz, D, U, S, inv, small, total; };
static LlhLayout llh_layout(int B, int R, int L, int n_ops) {
    LlhLayout l;
    l.D = up256((size_t)B * (R - 1) * L * 16 * 8);
    l.U = up256((size_t)B * (2 * R - 2) * L * 16 * 8);
    l.S = up256((size_t)B * L * 16 * 8);
    l.inv = up256((size_t)B * L * 8);
    // staged host arrays: children, ops, brlen, model, llh
    l.small = up256((size_t)B * (R - 1) * 2 * 4) + up256((size_t)B * n_ops * 4 * 4) + up256((size_t)B * (2 * R - 2) * 8) + up256((size_t)B * LLH_MODEL * 8) +
              up256((size_t)B * 2 * 8);
    l.tipsz = 0;
    l.total = l.D + l.U + l.S + l.inv + l.small + 256;
    return l;
}

size_t llh_ws_bytes(int B, int R, int L) { return llh_layout(B, R, L, 5 * R).total; }

// depth-first schedule of oracle/llh_oracle.py build_ops (same order: the two implementations visit the branches identically)
static int build_ops(const int32_t* ch, int R, std::vector<int32_t>& ops) {
    struct Item { int what, v, x, y; };   // 0 enter(optimise = x), 1 up, 2 down
    auto is_inner = [&](int v) { return v >= R; };
    std::vector<Item> st;
    auto run = [&](int v0, int optimise) {
        st.push_back({0, v0, optimise, 0});
        while (!st.empty()) {
            const Item it = st.back(); st.pop_back();
            if (it.what == 0) {
                if (it.x) { ops.insert(ops.end(), {OP_OPT, it.v, 0, 0}); }
                if (is_inner(it.v)) {
                    const int a = ch[2 * (it.v - R)], b = ch[2 * (it.v - R) + 1];
                    st.push_back({2, it.v, a, b});
                    st.push_back({0, b, 1, 0});
                    st.push_back({1, b, it.v, a});
                    st.push_back({0, a, 1, 0});
                    st.push_back({1, a, it.v, b});
                }
            } else if (it.what == 1) {
                ops.insert(ops.end(), {OP_UP, it.v, it.x, it.y});
            } else {
                ops.insert(ops.end(), {OP_DOWN, it.v, it.x, it.y});
            }
        }
    };
    const int c1 = ch[2 * (R - 2)], c2 = ch[2 * (R - 2) + 1];
    run(c1, 1);
    ops.insert(ops.end(), {OP_SWAPROOT, c2, c1, 0});
    run(c2, 0);
    return (int)(ops.size() / 4);
}

static int check_tree(const int32_t* ch, int R) {
    std::vector<char> used(2 * R - 1, 0);
    for (int k = 0; k < R - 1; ++k)
        for (int e = 0; e < 2; ++e) {
            const int c = ch[2 * k + e];
            if (c < 0 || c >= R + k || used[c]) return -1;      // children are created before their parent, every node has one parent
            used[c] = 1;
        }
    return 0;
}

int run_llh(const uint8_t* tips, const double* weights, const int32_t* children_h, double* brlen_h, double* model_h, int B, int R, int L,
            int optimise, int max_passes, double eps, double lh_eps, int max_rounds, int golden_iters, double* llh_h, void* ws, size_t ws_bytes,
            cudaStream_t st) {
    std::vector<int32_t> ops_all;
    int n_ops = 0;
    for (int b = 0; b < B; ++b) {
        if (check_tree(children_h + (size_t)b * (R - 1) * 2, R)) return set_error(NNJ_ERR_INVALID, "llh: children is not a binary tree in join order");
        if (optimise) {
            std::vector<int32_t> ops;
            const int n = build_ops(children_h + (size_t)b * (R - 1) * 2, R, ops);
            if (b == 0) n_ops = n;
            if (n != n_ops) return set_error(NNJ_ERR_INVALID, "llh: internal error (schedule length differs between trees)");
            ops_all.insert(ops_all.end(), ops.begin(), ops.end());
        }
    }
    const LlhLayout lay = llh_layout(B, R, L, n_ops);
    if (ws_bytes < lay.total) return set_error(NNJ_ERR_WORKSPACE, "llh: workspace too small (nnj_llh_workspace_bytes)");
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    LlhArgs a;
    a.tips = tips; a.weights = weights;
    a.D = reinterpret_cast<double*>(p); p += lay.D;
    a.U = reinterpret_cast<double*>(p); p += lay.U;
    a.S = reinterpret_cast<double*>(p); p += lay.S;
    a.inv = reinterpret_cast<double*>(p); p += lay.inv;
    int32_t* d_ch = reinterpret_cast<int32_t*>(p); p += up256((size_t)B * (R - 1) * 2 * 4);
    int32_t* d_ops = reinterpret_cast<int32_t*>(p); p += up256((size_t)B * n_ops * 4 * 4);
    double* d_br = reinterpret_cast<double*>(p); p += up256((size_t)B * (2 * R - 2) * 8);
    double* d_md = reinterpret_cast<double*>(p); p += up256((size_t)B * LLH_MODEL * 8);
    double* d_ll = reinterpret_cast<double*>(p);
    cudaError_t e = cudaMemcpyAsync(d_ch, children_h, (size_t)B * (R - 1) * 2 * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_ops) e = cudaMemcpyAsync(d_ops, ops_all.data(), (size_t)B * n_ops * 4 * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_br, brlen_h, (size_t)B * (2 * R - 2) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_md, model_h, (size_t)B * LLH_MODEL * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    a.children = d_ch; a.ops = d_ops; a.brlen = d_br; a.model = d_md; a.llh = d_ll;
    a.B = B; a.R = R; a.L = L; a.n_ops = n_ops; a.max_passes = max_passes; a.optimise = optimise; a.eps = eps;
    a.lh_eps = lh_eps; a.max_rounds = max_rounds; a.golden_iters = golden_iters; a.model_out = d_md;
    const size_t smem = (LLH_MODEL + 128 + 66 + 48 + 2 * (size_t)R) * sizeof(double);
    prof_begin(KC_LLH, st);
    k_llh<<<B, LLH_THREADS, smem, st>>>(a);
    ++g_launches;
    prof_end(st);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(llh_h, d_ll, (size_t)B * 2 * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && optimise) e = cudaMemcpyAsync(brlen_h, d_br, (size_t)B * (2 * R - 2) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && optimise == 2) e = cudaMemcpyAsync(model_h, d_md, (size_t)B * LLH_MODEL * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // the small results are host values: the call returns them
    if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
    return 0;
}

// ---- discrete gamma (Yang 1994): mean rate of each of ncat equal-probability classes of Gamma(alpha, alpha)
static double gammp(double a, double x) {     // regularised lower incomplete gamma P(a, x): series / continued fraction (Lentz)
    if (x <= 0.0) return 0.0;
    const double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 2000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
        return sum * exp(-x + a * log(x) - gln);
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
}

int gamma_rates(double alpha, int ncat, double* out) {
    if (!(alpha > 0.0) || ncat < 1 || ncat > 64) return set_error(NNJ_ERR_INVALID, "gamma_rates: alpha must be > 0 and 1 <= ncat <= 64");
    std::vector<double> cdf(ncat + 1, 0.0);
    cdf[ncat] = 1.0;
    for (int i = 1; i < ncat; ++i) {
        // quantile x of Gamma(alpha, 1) at i / ncat: Newton on P(alpha, x) (monotone, density known) kept inside a bracket, a bisection
        // step whenever Newton leaves it
        const double target = (double)i / ncat, gln = lgamma(alpha);
        double lo = 0.0, hi = alpha + 1.0;
        while (gammp(alpha, hi) < target) hi *= 2.0;
        double x = 0.5 * (lo + hi);
        for (int it = 0; it < 200; ++it) {
            const double f = gammp(alpha, x) - target;
            if (f < 0.0) lo = x; else hi = x;
            const double pdf = exp((alpha - 1.0) * log(x) - x - gln);
            double xn = pdf > 0.0 ? x - f / pdf : 0.5 * (lo + hi);
            if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
            const bool done = fabs(xn - x) <= 1e-15 * x || hi - lo <= 1e-300;
            x = xn;
            if (done) break;
        }
        cdf[i] = gammp(alpha + 1.0, x);
    }
    for (int i = 0; i < ncat; ++i) out[i] = ncat * (cdf[i + 1] - cdf[i]);
    return 0;
}

}  // namespace nnj
