// Per-trajectory ellipsoid calculus in float64:
//   ellipsoid_step   body of onestep_reachability after the ssm(...) call (gp_reachability.py:75-88 point branch,
//                    :102-156 set branch), with compute_remainder_overapproximations (utils.py:108-144),
//                    ellipsoid_from_rectangle (utils_ellipsoid.py:197-233) and both sum_two_ellipsoids
//                    (utils_ellipsoid.py:63-94) fused; optional GP-input transform (gp_reachability_casadi.py:60-98).
//                    In the fused path it also finishes the GP posterior: fixed-order reduction of the mean /
//                    Jacobian / |L^-1 k*|^2 partials written by the K* and contraction kernels.
//                    Two phases per block of 32 trajectories: (A) the partial sums, one warp per (output dimension,
//                    quantity) with the trajectory index across the lanes (coalesced 256-byte rows, eight loads in
//                    flight per lane), results parked in shared memory; (B) the algebra, one trajectory per lane of
//                    warp 0, everything in registers.
//   remainder, sum_two, from_rectangle, safety_distance : the batched leaves behind the Python mirrors of
//                    utils.py / utils_ellipsoid.py / gp_reachability.py:215-250.
// The reference takes max eig(Q (I + K^T K)) with a general eigen-solver (utils.py:133-134); Q B is similar to the
// symmetric C^T Q C with B = C C^T, so a Jacobi iteration on that gives the same spectrum (n_s = 2: closed form).
#include <math.h>

#include "segp_internal.cuh"

namespace segp {

// Small state dimensions are fully unrolled into registers; larger ones keep rolled loops over local-memory arrays
// (the ellipsoid algebra is O(n_s^3) per trajectory-step against O(n_s N^2) for the GP, so it never matters).
__host__ __device__ constexpr int unroll_factor(int n) { return n <= 4 ? 32 : 1; }
// innermost loops of the generic instance: four iterations in flight (their loads hit local memory; rolled one by one
// every iteration waits out the full load latency)
__host__ __device__ constexpr int unroll_inner(int n) { return n <= 4 ? 32 : 4; }

// Jacobi rotation annihilating m_pr: t = sgn(alpha) beta / (|alpha| + sqrt(alpha^2 + beta^2)), alpha = (m_rr - m_pp) / 2,
// beta = m_pr (the textbook t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) without forming theta = alpha / beta).
__device__ __forceinline__ void jacobi_cs(double app, double arr, double apr, double& cs, double& sn) {
    const double alpha = 0.5 * (arr - app);
    const double h2 = fma(alpha, alpha, apr * apr);
    const double hyp = h2 * rsqrt_fast(h2);
    double t = (alpha >= 0.0 ? apr : -apr) * rcp_fast(fabs(alpha) + hyp);
    if (apr == 0.0) t = 0.0;   // nothing to annihilate (also alpha == beta == 0, where the line above is NaN)
    cs = rsqrt_fast(fma(t, t, 1.0));
    sn = t * cs;
}
template <int N>
__device__ __forceinline__ void jacobi_apply(double (&m)[N][N], int p, int r, double cs, double sn) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double akp = m[k][p], akr = m[k][r];
        m[k][p] = cs * akp - sn * akr;
        m[k][r] = sn * akp + cs * akr;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double apk = m[p][k], ark = m[r][k];
        m[p][k] = cs * apk - sn * ark;
        m[r][k] = sn * apk + cs * ark;
    }
}

// M = C^T Q C with B = I + K^T K = C C^T: symmetric, same spectrum as Q B.  Q symmetric positive semi-definite (n x n),
// K (n_u x n)
template <int N, int NU>
__device__ __forceinline__ void build_similar(const double (&q)[N][N], const double (&kfb)[NU][N], int n, int nu,
                                              double (&m)[N][N]) {
    constexpr int UF = unroll_factor(N), UI = unroll_inner(N);
    // loop bounds: the compile-time capacity for the register-resident instances (fully unrolled), the run-time size
    // for the generic one (rolled loops over local-memory arrays: 16^3 -> n^3 work)
    const int NL = (N > 4) ? n : N, NUL = (N > 4) ? nu : NU;
    double bm[N][N];
#pragma unroll UF
    for (int i = 0; i < NL; ++i)
#pragma unroll UF
        for (int j = 0; j < NL; ++j) {
            double acc = (i == j) ? 1.0 : 0.0;
            if (i < n && j < n) {
#pragma unroll UF
                for (int u = 0; u < NUL; ++u)
                    if (u < nu) acc = fma(kfb[u][i], kfb[u][j], acc);
            }
            bm[i][j] = acc;
        }
    // Cholesky B = C C^T (B >= I, always positive definite), in place, lower
#pragma unroll UF
    for (int j = 0; j < NL; ++j) {
        if (j < n) {
            double d = bm[j][j];
#pragma unroll UI
            for (int k = 0; k < NL; ++k)
                if (k < j) d = fma(-bm[j][k], bm[j][k], d);
            d = sqrt(d);
            bm[j][j] = d;
            const double inv = 1.0 / d;
#pragma unroll UF
            for (int i = 0; i < NL; ++i) {
                if (i > j && i < n) {
                    double s = bm[i][j];
#pragma unroll UI
                    for (int k = 0; k < NL; ++k)
                        if (k < j) s = fma(-bm[i][k], bm[j][k], s);
                    bm[i][j] = s * inv;
                }
            }
        }
    }
    // M = C^T Q C ; first T = Q C (T[i][j] = sum_{k>=j} Q[i][k] C[k][j])
    double t[N][N];
#pragma unroll UF
    for (int i = 0; i < NL; ++i)
#pragma unroll UF
        for (int j = 0; j < NL; ++j) {
            double acc = 0.0;
#pragma unroll UI
            for (int k = 0; k < NL; ++k)
                if (k >= j && k < n && i < n) acc = fma(0.5 * (q[i][k] + q[k][i]), bm[k][j], acc);
            t[i][j] = acc;
        }
#pragma unroll UF
    for (int i = 0; i < NL; ++i)
#pragma unroll UF
        for (int j = 0; j < NL; ++j) {
            double acc = 0.0;
            if (j >= i) {
#pragma unroll UI
                for (int k = 0; k < NL; ++k)
                    if (k >= i && k < n) acc = fma(bm[k][i], t[k][j], acc);
            }
            m[i][j] = acc;
        }
#pragma unroll UF
    for (int i = 0; i < NL; ++i)
#pragma unroll UF
        for (int j = 0; j < NL; ++j)
            if (j < i) m[i][j] = m[j][i];
}

// largest eigenvalue of Q (I + K^T K)
template <int N, int NU>
__device__ __forceinline__ double lambda_max_qb(const double (&q)[N][N], const double (&kfb)[NU][N], int n, int nu) {
    constexpr int UF = unroll_factor(N), UI = unroll_inner(N);
    const int NL = (N > 4) ? n : N;
    double m[N][N];
    build_similar<N, NU>(q, kfb, n, nu, m);
    if constexpr (N == 2) {
        // closed form: the larger root of the 2 x 2 characteristic polynomial (stable: both terms are >= 0)
        const double hm = 0.5 * (m[0][0] + m[N - 1][N - 1]), hd = 0.5 * (m[0][0] - m[N - 1][N - 1]);
        return hm + sqrt(fma(hd, hd, m[0][N - 1] * m[0][N - 1]));
    }
    if constexpr (N == 4) {
        // Jacobi in the round-robin ordering: the two rotations of a round touch disjoint index pairs, so their
        // parameters are computed side by side (two independent dependent chains in flight) and a sweep is three
        // rounds deep instead of six rotations; padded rows / columns (n < 4) are zero and rotate as no-ops.
        for (int sweep = 0; sweep < 30; ++sweep) {
            double off = 0.0, dg = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                dg = fma(m[i][i], m[i][i], dg);
#pragma unroll
                for (int j = 0; j < N; ++j)
                    if (j > i) off = fma(m[i][j], m[i][j], off);
            }
            if (!(off > 1e-26 * dg)) break;   // relative off-diagonal norm < 1e-13; also exits on NaN / all-zero
#pragma unroll
            for (int round = 0; round < 3; ++round) {
                const int p0 = 0, r0 = round + 1;
                const int p1 = round == 0 ? 2 : 1, r1 = round == 2 ? 2 : 3;
                double c0, s0, c1, s1;
                jacobi_cs(m[p0][p0], m[r0][r0], m[p0][r0], c0, s0);
                jacobi_cs(m[p1][p1], m[r1][r1], m[p1][r1], c1, s1);
                jacobi_apply<N>(m, p0, r0, c0, s0);
                jacobi_apply<N>(m, p1, r1, c1, s1);
            }
        }
    } else {
    // cyclic Jacobi on the symmetric M (eigenvalues only)
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dg = 0.0;
#pragma unroll UF
        for (int i = 0; i < NL; ++i) {
            dg = fma(m[i][i], m[i][i], dg);
#pragma unroll UF
            for (int j = 0; j < NL; ++j)
                if (j > i) off = fma(m[i][j], m[i][j], off);
        }
        if (!(off > 1e-26 * dg)) break;   // relative off-diagonal norm < 1e-13; also exits on NaN / all-zero
#pragma unroll UF
        for (int p = 0; p < NL; ++p) {
#pragma unroll UF
            for (int r = 0; r < NL; ++r) {
                if (r > p && r < n) {
                    const double apq = m[p][r];
                    if (apq != 0.0) {
                        double cs, sn;
                        jacobi_cs(m[p][p], m[r][r], apq, cs, sn);
#pragma unroll UI
                        for (int k = 0; k < NL; ++k) {
                            const double akp = m[k][p], akr = m[k][r];
                            m[k][p] = cs * akp - sn * akr;
                            m[k][r] = sn * akp + cs * akr;
                        }
#pragma unroll UI
                        for (int k = 0; k < NL; ++k) {
                            const double apk = m[p][k], ark = m[r][k];
                            m[p][k] = cs * apk - sn * ark;
                            m[r][k] = sn * apk + cs * ark;
                        }
                    }
                }
            }
        }
    }
    }
    double lam = m[0][0];
#pragma unroll UF
    for (int i = 1; i < NL; ++i)
        if (i < n) lam = fmax(lam, m[i][i]);
    return lam;
}

constexpr int ELL_TB = 32;      // trajectories per block: one per lane
constexpr int ELL_WARPS = 8;    // warps of phase A (and of the cooperative Jacobi)
static_assert(2 * ELL_WARPS >= SEGP_MAX_NS, "one warp per disjoint rotation of a round");

// Cooperative Jacobi for the generic instance (n_s > 4): the matrices of the block's 32 trajectories sit in shared
// memory (element (i, j) of lane l at sm[(i n + j) 32 + l]) and the ceil(n / 2) disjoint rotations of a round of the
// round-robin ordering run on one warp each (n <= 16 -> at most 8 pairs = ELL_WARPS); lane = trajectory throughout.
// One thread per trajectory with the matrix in local memory is a dependent chain of ~360 rotations x 2 n element
// updates per trajectory (C5: 1.8 ms for 8192 trajectories, 5 % of a step).  A lane whose matrix has converged
// rotates by the identity (exact no-op), so a trajectory's result does not depend on its neighbours.  On return
// (block-synchronised) the eigenvalues are on the diagonal.
__device__ __forceinline__ void jacobi_coop(double* __restrict__ sm, int* __restrict__ s_flag, int n, int warp, int lane) {
    const int np = n + (n & 1);      // players of the tournament (a dummy for odd n)
    const int rounds = np - 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        if (warp == 0) {
            double off = 0.0, dg = 0.0;
            for (int i = 0; i < n; ++i) {
                const double d = sm[(i * n + i) * 32 + lane];
                dg = fma(d, d, dg);
                for (int j = i + 1; j < n; ++j) {
                    const double o = sm[(i * n + j) * 32 + lane];
                    off = fma(o, o, off);
                }
            }
            s_flag[lane] = !(off > 1e-26 * dg);   // relative off-diagonal norm < 1e-13; also NaN / all-zero
        }
        __syncthreads();
        const bool done = s_flag[lane] != 0;
        if (__syncthreads_and(done ? 1 : 0)) break;
        for (int r = 0; r < rounds; ++r) {
            // pair of this warp in round r (circle method: player np - 1 stays, the others rotate)
            int p = -1, q = -1;
            if (warp < np / 2) {
                const int x = warp == 0 ? np - 1 : (r + warp) % (np - 1);
                const int y = warp == 0 ? r : (r - warp + (np - 1)) % (np - 1);
                p = min(x, y);
                q = max(x, y);
                if (q >= n) p = -1;      // the dummy's partner rests
            }
            double cs = 1.0, sn = 0.0;
            if (p >= 0) {
                if (!done) jacobi_cs(sm[(p * n + p) * 32 + lane], sm[(q * n + q) * 32 + lane], sm[(p * n + q) * 32 + lane], cs, sn);
#pragma unroll 4
                for (int k = 0; k < n; ++k) {     // columns p, q (disjoint from every other pair's)
                    const double akp = sm[(k * n + p) * 32 + lane], akq = sm[(k * n + q) * 32 + lane];
                    sm[(k * n + p) * 32 + lane] = cs * akp - sn * akq;
                    sm[(k * n + q) * 32 + lane] = sn * akp + cs * akq;
                }
            }
            __syncthreads();
            if (p >= 0) {
#pragma unroll 4
                for (int k = 0; k < n; ++k) {     // rows p, q
                    const double apk = sm[(p * n + k) * 32 + lane], aqk = sm[(q * n + k) * 32 + lane];
                    sm[(p * n + k) * 32 + lane] = cs * apk - sn * aqk;
                    sm[(q * n + k) * 32 + lane] = sn * apk + cs * aqk;
                }
            }
            __syncthreads();
        }
    }
}

// =========================================================================================== ellipsoid_step
// NS/NU are compile-time capacities; n_s/n_u the run-time sizes (equal for the specialised instances).

template <int NS, int NU>
__global__ void __launch_bounds__(ELL_TB * ELL_WARPS) ellipsoid_step_kernel(const StepArgs a) {
    constexpr int UF = unroll_factor(NS), UI = unroll_inner(NS);
    extern __shared__ double ell_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long b = a.b0 + (long)blockIdx.x * ELL_TB + lane;
    const bool live = b < a.n_batch;
    const int n_s = a.n_s, n_u = a.n_u, n_in = a.n_in;
    const int dim = n_in + n_u;
    const bool fused = a.mu_part != nullptr;
    double* s_mu = ell_smem;                  // [n_s][32]
    double* s_qf = s_mu + n_s * ELL_TB;       // [n_s][32]  |L^-1 k*|^2
    double* s_e2 = s_qf + n_s * ELL_TB;       // [n_s][32]  error-model variance of it (int8 contraction)
    double* s_jac = s_e2 + n_s * ELL_TB;      // [n_s][dim][32]  final Jacobian entries (set branch only)

    // ---- phase A (fused path): fixed-order sums of the partials, one warp per (d, quantity); a lane's loads do not
    // depend on its running sum, so eight are in flight at a time
    if (fused) {
        const int per_d = 3 + (a.q != nullptr ? dim : 0);
        for (int item = warp; item < n_s * per_d; item += ELL_WARPS) {
            const int d = item / per_d, kind = item - d * per_d;
            double acc = 0.0;
            if (live) {
                if (kind == 0) {
#pragma unroll 8
                    for (int s = 0; s < a.nsplit; ++s) acc += a.mu_part[((long)s * n_s + d) * a.b_cap + b];
                } else if (kind == 1) {
#pragma unroll 8
                    for (int i = 0; i < a.nblk; ++i) acc += a.qpart[((long)d * a.nblk + i) * a.b_cap + b];
                } else if (kind == 2) {
                    if (a.epart != nullptr) {
                        float e2 = 0.f;
#pragma unroll 8
                        for (int i = 0; i < a.nblk; ++i) e2 += a.epart[((long)d * a.nblk + i) * a.b_cap + b];
                        acc = (double)e2;
                    }
                } else {
                    const int j = kind - 3;
                    double sum = 0.0;
#pragma unroll 8
                    for (int s = 0; s < a.nsplit; ++s) sum += a.jac_part[(((long)s * n_s + d) * dim + j) * a.b_cap + b];
                    acc = -sum * a.invls[d * dim + j];
                    if (a.jac2_part != nullptr) {   // composite kernels: additive terms, final form
                        double add = 0.0;
#pragma unroll 8
                        for (int s = 0; s < a.nsplit; ++s)
                            add += a.jac2_part[(((long)s * n_s + d) * dim + j) * a.b_cap + b];
                        acc += add;
                    }
                }
            }
            double* dst = kind == 0 ? s_mu + d * ELL_TB
                                    : kind == 1 ? s_qf + d * ELL_TB
                                                : kind == 2 ? s_e2 + d * ELL_TB : s_jac + (d * dim + kind - 3) * ELL_TB;
            dst[lane] = acc;
        }
        __syncthreads();
    }
    // ---- generic instance, set branch of the reachability step: lambda_max(Q (I + K^T K)) depends on the state only,
    // not on the GP, and is found by all warps together (jacobi_coop) before warp 0 goes on alone
    double* s_m = s_jac + n_s * dim * ELL_TB;                           // [n_s n_s][32]
    int* s_flag = reinterpret_cast<int*>(s_m + n_s * n_s * ELL_TB);     // [32]
    double q[NS][NS], kfb[NU][NS];   // shape matrix and feedback gain of the set branch
    bool coop = false;
    if constexpr (NS > 4) {
        coop = a.q != nullptr && a.sp->prop_mode == SEGP_PROP_ELLIPSOID;
        if (coop) {
            if (warp == 0) {
                double m[NS][NS];
                for (int i = 0; i < n_s; ++i)
                    for (int j = 0; j < n_s; ++j) q[i][j] = live ? a.q[b * a.q_stride + i * n_s + j] : 0.0;
                for (int i = 0; i < n_u; ++i)
                    for (int j = 0; j < n_s; ++j) kfb[i][j] = live ? a.kfb[b * a.kfb_stride + i * n_s + j] : 0.0;
                build_similar<NS, NU>(q, kfb, n_s, n_u, m);
                for (int i = 0; i < n_s; ++i)
                    for (int j = 0; j < n_s; ++j) s_m[(i * n_s + j) * ELL_TB + lane] = m[i][j];
            }
            __syncthreads();
            jacobi_coop(s_m, s_flag, n_s, warp, lane);
        }
    }
    if (warp != 0 || !live) return;

    // ---- phase B: one trajectory per lane.  Loop bounds: compile-time capacities for the register-resident instances,
    // run-time sizes for the generic one (see lambda_max_qb)
    const int NSL = (NS > 4) ? n_s : NS, NUL = (NS > 4) ? n_u : NU, NDL = (NS > 4) ? dim : NS + NU;
    const StepParams* __restrict__ sp = a.sp;
    int32_t status = 0;
    double mu[NS], var[NS];
#pragma unroll UF
    for (int d = 0; d < NSL; ++d) {
        mu[d] = 0.0;
        var[d] = 0.0;
        if (d < n_s) {
            if (fused) {
                mu[d] = s_mu[d * ELL_TB + lane];
                var[d] = (a.kss != nullptr ? a.kss[(long)d * a.b_cap + b] : a.gp_var[d]) - s_qf[d * ELL_TB + lane];
                // int8 contraction: a-posteriori error estimate against the variance
                if (a.epart != nullptr && a.guard_gs * s_e2[d * ELL_TB + lane] > var[d] * var[d])
                    status |= SEGP_STATUS_LOW_PRECISION;
            } else {
                mu[d] = a.mu_d[b * n_s + d];
                var[d] = a.var_d[b * n_s + d];
            }
            if (a.var_out != nullptr) a.var_out[b * a.var_out_stride + d] = var[d];
            if (!(var[d] > 0.0)) status |= SEGP_STATUS_BAD_VARIANCE;
        }
    }

    double p[NS], u[NU];
#pragma unroll UF
    for (int i = 0; i < NSL; ++i) p[i] = (i < n_s) ? a.p[b * a.p_stride + i] : 0.0;
#pragma unroll UF
    for (int i = 0; i < NUL; ++i) u[i] = (i < n_u) ? a.kff[b * a.kff_stride + i] : 0.0;

    // p_lin = mu + A p + B k_ff   (gp_reachability.py:82-83, :115)
    double p1[NS];
#pragma unroll UF
    for (int i = 0; i < NSL; ++i) {
        double acc = mu[i];
        if (i < n_s) {
#pragma unroll UF
            for (int k = 0; k < NSL; ++k)
                if (k < n_s) acc = fma(sp->a[i * n_s + k], p[k], acc);
#pragma unroll UF
            for (int k = 0; k < NUL; ++k)
                if (k < n_u) acc = fma(sp->b[i * n_u + k], u[k], acc);
        }
        p1[i] = acc;
    }
    const double c = sp->c_safety;
    const int mode = sp->prop_mode;   // 0 ellipsoid reachability, 1 Taylor, 2 mean-equivalent Gaussian propagation
    double q1[NS][NS];

    if (a.q == nullptr) {
        // ---- point branch (gp_reachability.py:65-88): Q1 = diag(n_s (c sigma_d)^2);
        //      Gaussian propagation (uncertainty_propagation_casadi.py:52-57, 256-261): Sigma1 = diag(sigma_d^2)
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j) q1[i][j] = 0.0;
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
            if (i < n_s) {
                if (mode != SEGP_PROP_ELLIPSOID) {
                    q1[i][i] = var[i];
                } else {
                    const double bound = c * sqrt(var[i]);
                    if (!(bound > 0.0)) status |= SEGP_STATUS_ZERO_BOUND;
                    q1[i][i] = (double)n_s * bound * bound;
                }
            }
    } else {
        // ---- set branch (gp_reachability.py:89-156)
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j) q[i][j] = (i < n_s && j < n_s) ? a.q[b * a.q_stride + i * n_s + j] : 0.0;
#pragma unroll UF
        for (int i = 0; i < NUL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j)
                kfb[i][j] = (i < n_u && j < n_s) ? a.kfb[b * a.kfb_stride + i * n_s + j] : 0.0;

        // H = A + A_mu + (B_mu + B) K_fb ; A_mu = J[:, :n_in] (T), B_mu = J[:, n_in:]
        double h[NS][NS];
#pragma unroll UF
        for (int d = 0; d < NSL; ++d) {
            if (d < n_s) {
                // Jacobian row d: jrow[j], j < dim
                double jrow[NS + NU];
#pragma unroll UF
                for (int j = 0; j < NDL; ++j) {
                    jrow[j] = 0.0;
                    if (j < dim && mode != SEGP_PROP_MEAN_EQUIVALENT)   // mean equivalent: no linearisation term
                        jrow[j] = fused ? s_jac[(d * dim + j) * ELL_TB + lane] : a.jac_d[(b * n_s + d) * dim + j];
                }
#pragma unroll UF
                for (int j = 0; j < NSL; ++j) {
                    if (j < n_s) {
                        double acc = sp->a[d * n_s + j];
                        if (sp->has_t) {
                            for (int i = 0; i < n_in; ++i) {
                                double ji = 0.0;
#pragma unroll UF
                                for (int jj = 0; jj < NDL; ++jj)
                                    if (jj == i) ji = jrow[jj];
                                acc = fma(ji, sp->t[i * n_s + j], acc);
                            }
                        } else {
                            acc += jrow[j];
                        }
#pragma unroll UF
                        for (int k = 0; k < NUL; ++k) {
                            if (k < n_u) {
                                double jb = 0.0;
#pragma unroll UF
                                for (int jj = 0; jj < NDL; ++jj)
                                    if (jj == n_in + k) jb = jrow[jj];
                                acc = fma(jb + sp->b[d * n_u + k], kfb[k][j], acc);
                            }
                        }
                        h[d][j] = acc;
                    } else {
                        h[d][j] = 0.0;
                    }
                }
            } else {
#pragma unroll UF
                for (int j = 0; j < NSL; ++j) h[d][j] = 0.0;
            }
        }
        // Q0 = H Q H^T
        double hq[NS][NS];
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j) {
                double acc = 0.0;
#pragma unroll UI
                for (int k = 0; k < NSL; ++k) acc = fma(h[i][k], q[k][j], acc);
                hq[i][j] = acc;
            }
        double tr0 = 0.0;
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j) {
                double acc = 0.0;
#pragma unroll UI
                for (int k = 0; k < NSL; ++k) acc = fma(hq[i][k], h[j][k], acc);
                q1[i][j] = acc;
                if (i == j) tr0 += acc;
            }
        if (mode != SEGP_PROP_ELLIPSOID) {
            // Gaussian propagation: [A B I] Sigma_all [A B I]^T of uncertainty_propagation_casadi.py:59-87 (Taylor)
            // and :263-283 (mean equivalent) collapses to H Sigma H^T + diag(sigma^2) with the H above
            // (H = A + B K for the mean-equivalent variant): with F = [I; K], Sigma_z = F Sigma F^T and
            // Sigma_zg = Sigma_z J^T, so every block carries the factor F Sigma F^T.
#pragma unroll UF
            for (int i = 0; i < NSL; ++i)
                if (i < n_s) q1[i][i] += var[i];
        } else {
        // remainder boxes (utils.py:129-142)
        double r2;
        if constexpr (NS > 4) {      // found by the whole block above: the eigenvalues are on the diagonal
            r2 = s_m[lane];
            for (int i = 1; i < n_s; ++i) r2 = fmax(r2, s_m[(i * n_s + i) * ELL_TB + lane]);
        } else {
            r2 = lambda_max_qb<NS, NU>(q, kfb, n_s, n_u);
        }
        const double r1 = sqrt(r2);
        double tr_sig = 0.0, tr_mu = 0.0;
        double d_sig[NS], d_mu[NS];
#pragma unroll UF
        for (int i = 0; i < NSL; ++i) {
            d_sig[i] = d_mu[i] = 0.0;
            if (i < n_s) {
                const double ub_mu = sp->l_mu[i] * r2;
                const double ub_sig = sp->l_sigma[i] * r1;
                const double bs = c * (sqrt(var[i]) + ub_sig);
                if (!(bs > 0.0) || !(ub_mu > 0.0)) status |= SEGP_STATUS_ZERO_BOUND;
                d_sig[i] = (double)n_s * bs * bs;
                d_mu[i] = (double)n_s * ub_mu * ub_mu;
                tr_sig += d_sig[i];
                tr_mu += d_mu[i];
            }
        }
        // (0, Q_L) = Q_sigma (+) Q_mu ; (p1, Q1) = (0, Q_L) (+) (p0, Q0)     (utils_ellipsoid.py:88-94)
        const double c1 = sqrt(tr_sig / tr_mu);
        double tr_l = 0.0;
#pragma unroll UF
        for (int i = 0; i < NSL; ++i) {
            d_sig[i] = (1.0 + 1.0 / c1) * d_sig[i] + (1.0 + c1) * d_mu[i];
            tr_l += d_sig[i];
        }
        const double c2 = sqrt(tr_l / tr0);
        const double f_l = 1.0 + 1.0 / c2, f_0 = 1.0 + c2;
#pragma unroll UF
        for (int i = 0; i < NSL; ++i)
#pragma unroll UF
            for (int j = 0; j < NSL; ++j) q1[i][j] = f_0 * q1[i][j] + ((i == j) ? f_l * d_sig[i] : 0.0);
        }
    }

    bool finite = true;
#pragma unroll UF
    for (int i = 0; i < NSL; ++i) {
        if (i < n_s) {
            a.p_out[b * a.p_out_stride + i] = p1[i];
            finite = finite && isfinite(p1[i]);
#pragma unroll UF
            for (int j = 0; j < NSL; ++j)
                if (j < n_s) {
                    a.q_out[b * a.q_out_stride + i * n_s + j] = q1[i][j];
                    finite = finite && isfinite(q1[i][j]);
                }
        }
    }
    if (!finite) status |= SEGP_STATUS_NONFINITE;
    if (a.status != nullptr && status != 0) atomicOr(&a.status[b], status);
}

int launch_ellipsoid_step(const StepArgs& a, cudaStream_t st) {
    // shared memory: phase A, n_s (3 + dim) rows of 32 doubles (C4: 8 KB; the largest supported model: 108 KB); the
    // generic instance adds the n_s x n_s matrices of the cooperative Jacobi (n_s = 16: 64 KB) and 32 flags
    constexpr int kMaxSmem = SEGP_MAX_NS * (3 + MAX_D + SEGP_MAX_NS) * ELL_TB * (int)sizeof(double) + 128;
    static bool attr_set[64] = {};
    if (first_call_on_device(attr_set)) {
        SEGP_CUDA_CHECK(cudaFuncSetAttribute(ellipsoid_step_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        SEGP_CUDA_CHECK(cudaFuncSetAttribute(ellipsoid_step_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        SEGP_CUDA_CHECK(cudaFuncSetAttribute(ellipsoid_step_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        SEGP_CUDA_CHECK(cudaFuncSetAttribute(ellipsoid_step_kernel<SEGP_MAX_NS, SEGP_MAX_NU>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    }
    if (a.n_batch <= a.b0) return SEGP_OK;
    const int threads = ELL_TB * ELL_WARPS;
    const unsigned grid = (unsigned)((a.n_batch - a.b0 + ELL_TB - 1) / ELL_TB);
    const bool generic = !((a.n_s == 2 && a.n_u == 1 && a.n_in <= 2) || (a.n_s <= 4 && a.n_u <= 2 && a.n_in <= 4));
    const size_t smem = (size_t)a.n_s * (3 + a.n_in + a.n_u + (generic ? a.n_s : 0)) * ELL_TB * sizeof(double) +
                        (generic ? 128 : 0);
    // the specialised instances hold a Jacobian row in NS + NU registers: a lifting input transform (n_in > n_s) does
    // not fit and takes the generic instance
    if (a.n_s == 2 && a.n_u == 1 && a.n_in <= 2)
        ellipsoid_step_kernel<2, 1><<<grid, threads, smem, st>>>(a);
    else if (a.n_s == 4 && a.n_u == 1 && a.n_in <= 4)
        ellipsoid_step_kernel<4, 1><<<grid, threads, smem, st>>>(a);
    else if (a.n_s <= 4 && a.n_u <= 2 && a.n_in <= 4)
        ellipsoid_step_kernel<4, 2><<<grid, threads, smem, st>>>(a);
    else
        ellipsoid_step_kernel<SEGP_MAX_NS, SEGP_MAX_NU><<<grid, threads, smem, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== batched leaves
template <int NS, int NU>
__global__ void remainder_kernel(long n_batch, int n_s, int n_u, const double* __restrict__ q,
                                 const double* __restrict__ kfb, long kfb_stride, const StepParams* __restrict__ sp,
                                 double* __restrict__ u_mu, double* __restrict__ u_sigma) {
    constexpr int UF = unroll_factor(NS);
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    const int NSL = (NS > 4) ? n_s : NS, NUL = (NS > 4) ? n_u : NU;
    double qm[NS][NS], k[NU][NS];
#pragma unroll UF
    for (int i = 0; i < NSL; ++i)
#pragma unroll UF
        for (int j = 0; j < NSL; ++j) qm[i][j] = (i < n_s && j < n_s) ? q[b * n_s * n_s + i * n_s + j] : 0.0;
#pragma unroll UF
    for (int i = 0; i < NUL; ++i)
#pragma unroll UF
        for (int j = 0; j < NSL; ++j) k[i][j] = (i < n_u && j < n_s) ? kfb[b * kfb_stride + i * n_s + j] : 0.0;
    const double r2 = lambda_max_qb<NS, NU>(qm, k, n_s, n_u);
    const double r1 = sqrt(r2);
    for (int i = 0; i < n_s; ++i) {
        u_mu[b * n_s + i] = sp->l_mu[i] * r2;
        u_sigma[b * n_s + i] = sp->l_sigma[i] * r1;
    }
}

int launch_remainder(long n_batch, int n_s, int n_u, const double* q, const double* kfb, long kfb_stride,
                     const StepParams* sp, double* u_mu, double* u_sigma, cudaStream_t st) {
    const int threads = 64;
    const unsigned grid = (unsigned)((n_batch + threads - 1) / threads);
    if (n_s <= 4 && n_u <= 2)
        remainder_kernel<4, 2><<<grid, threads, 0, st>>>(n_batch, n_s, n_u, q, kfb, kfb_stride, sp, u_mu, u_sigma);
    else
        remainder_kernel<SEGP_MAX_NS, SEGP_MAX_NU>
            <<<grid, threads, 0, st>>>(n_batch, n_s, n_u, q, kfb, kfb_stride, sp, u_mu, u_sigma);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// utils_ellipsoid.py:63-94 with c = sqrt(tr Q1 / tr Q2)
__global__ void sum_two_kernel(long n_batch, int n, const double* __restrict__ p1, const double* __restrict__ q1,
                               const double* __restrict__ p2, const double* __restrict__ q2, double* __restrict__ p,
                               double* __restrict__ q) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    const double* a1 = q1 + b * n * n;
    const double* a2 = q2 + b * n * n;
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < n; ++i) {
        t1 += a1[i * n + i];
        t2 += a2[i * n + i];
    }
    const double c = sqrt(t1 / t2);
    const double f1 = 1.0 + 1.0 / c, f2 = 1.0 + c;
    for (int i = 0; i < n * n; ++i) q[b * n * n + i] = f1 * a1[i] + f2 * a2[i];
    for (int i = 0; i < n; ++i) p[b * n + i] = p1[b * n + i] + p2[b * n + i];
}

int launch_sum_two(long n_batch, int n, const double* p1, const double* q1, const double* p2, const double* q2,
                   double* p, double* q, cudaStream_t st) {
    sum_two_kernel<<<(unsigned)((n_batch + 127) / 128), 128, 0, st>>>(n_batch, n, p1, q1, p2, q2, p, q);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// utils_ellipsoid.py:197-233
__global__ void from_rectangle_kernel(long n_batch, int n, const double* __restrict__ ub, double* __restrict__ q,
                                      int32_t* __restrict__ status) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    int32_t st = 0;
    for (int i = 0; i < n; ++i) {
        const double v = ub[b * n + i];
        if (!(v > 0.0)) st |= SEGP_STATUS_ZERO_BOUND;
        for (int j = 0; j < n; ++j) q[b * n * n + i * n + j] = (i == j) ? (double)n * v * v : 0.0;
    }
    if (status != nullptr) status[b] = st;
}

int launch_from_rectangle(long n_batch, int n, const double* ub, double* q, int32_t* status, cudaStream_t st) {
    from_rectangle_kernel<<<(unsigned)((n_batch + 127) / 128), 128, 0, st>>>(n_batch, n, ub, q, status);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// gp_reachability.py:215-250: d = h_mat p + c sqrt(diag(h_mat Q h_mat^T)) - h_vec ; one thread per (item, constraint)
__global__ void safety_distance_kernel(long n_items, int n_s, int m, const double* __restrict__ p,
                                       const double* __restrict__ q, const double* __restrict__ hmat,
                                       const double* __restrict__ hvec, double c, double* __restrict__ dist) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items * m) return;
    const long b = idx / m;
    const int r = (int)(idx % m);
    const double* hr = hmat + (long)r * n_s;
    const double* qb = q + b * n_s * n_s;
    double dc = 0.0, ds = 0.0;
    for (int i = 0; i < n_s; ++i) {
        dc = fma(hr[i], p[b * n_s + i], dc);
        double acc = 0.0;
        for (int j = 0; j < n_s; ++j) acc = fma(qb[i * n_s + j], hr[j], acc);
        ds = fma(hr[i], acc, ds);
    }
    dist[idx] = dc + c * sqrt(ds) - hvec[r];
}

int launch_safety_distance(long n_items, int n_s, int m, const double* p, const double* q, const double* hmat,
                           const double* hvec, double c, double* dist, cudaStream_t st) {
    const long total = n_items * m;
    safety_distance_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(n_items, n_s, m, p, q, hmat, hvec, c,
                                                                           dist);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

}  // namespace segp
