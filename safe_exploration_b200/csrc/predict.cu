// GP predictive posterior kernels (float64):
//   kstar_mean_jac : K*_d[i,b] = k_d(z_b, x_i) written once per step in the MMA-operand layout, fused with the
//                    N-length reductions for the mean and the closed-form mean Jacobian
//                    (reference ssm_gpy/gp_models_utils_casadi.py:17-70, 177-197, 275-280).
//   tri_sumsq      : |L_d^-1 K*_d[:,b]|^2 as a block-triangular W x K* contraction on the FP64 tensor pipe
//                    (DMMA m8n8k4), operands streamed by TMA bulk copies through an mbarrier ring, only the
//                    per-column sums of squares leave the kernel (gp_models_utils_casadi.py:190-193 /
//                    GPy predict_noiseless).
//   finalize       : fixed-order reduction of the partials for segp_predict.
//
// Why float64 and not tcgen05/bf16: sigma^2 = k** - |L^-1 k*|^2 cancels 3.5-4 digits on the benchmark models and
// the mean cancels ~5; an fp32-product pipeline gives 2e-3 relative variance error (measured)
// against a 1e-4 gate.  tcgen05 has no f64 kind, so the dense contraction runs on the DMMA pipe.
#include <math.h>

#include "segp_internal.cuh"

namespace segp {

// =========================================================================================== kstar_mean_jac
constexpr int KS_THREADS = 128;

// Composite kernels  k(z, x_i) = lp_i stat_i + ll_i  with  lp_i = sum_j a_j z_j x_ij,  ll_i = sum_j v_j z_j x_ij,
// stat_i = s_f^2 phi(|s (z - x_i)|)  (_k_lin_rbf / _k_lin_mat52, gp_models_utils_casadi.py:73-129).  Same outputs as
// the stationary path (kernel block in the DMMA operand layout, mean partial) plus
//   jac_part  [j] = sum_i beta_i lp_i g_i s_j (z_j - x_ij)          (finalised as -s_j * ., like the stationary part)
//   jac2_part [j] = a_j sum_i beta_i stat_i x_ij  +  v_j (X^T beta)_j          (added as is)
//   kss           = s_f^2 sum_j a_j z_j^2 + sum_j v_j z_j^2                      (prior variance k(z, z))
// i.e. d k / d z_j = a_j x_ij stat_i - lp_i g_i s_j^2 (z_j - x_ij) + v_j x_ij.  Small models only (journal configs:
// N of a few hundred): one training point per iteration, raw inputs read through the cache.
template <int DM>
__device__ __forceinline__ void kstar_composite(const KstarArgs& a, const double (&z)[DM], int dim, int d, int split,
                                                long b, bool active) {
    if (!active) return;
    const int kern = a.kern[d];
    const double var = a.var[d];
    const double sqrt5 = 2.23606797749978969641;
    const double* __restrict__ pl = a.plin + d * dim;
    const double* __restrict__ lv = a.lin + d * dim;
    double zs[DM], za[DM], zv[DM], jac[DM], jac2[DM];
    double kss = 0.0, kss_lin = 0.0;
#pragma unroll
    for (int j = 0; j < DM; ++j) {
        zs[j] = za[j] = zv[j] = jac[j] = jac2[j] = 0.0;
        if (j < dim) {
            zs[j] = z[j] * a.invls[d * dim + j];
            za[j] = z[j] * pl[j];
            zv[j] = z[j] * lv[j];
            kss = fma(za[j], z[j], kss);
            kss_lin = fma(zv[j], z[j], kss_lin);
        }
    }
    double mu = 0.0;
    const int ngroups = a.n_pad / 4;
    const int g0 = split * a.groups_per_split;
    const int g1 = min(g0 + a.groups_per_split, ngroups);
    for (int g = g0; g < g1; ++g) {
        double kv[4];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const int row = g * 4 + qd;
            const double* __restrict__ xsr = a.xs + ((long)d * a.n_pad + row) * dim;
            const double* __restrict__ xr = a.xraw + (long)row * dim;
            double diff[DM];
            double r2 = 0.0, lp = 0.0, ll = 0.0;
#pragma unroll
            for (int j = 0; j < DM; ++j) {
                diff[j] = 0.0;
                if (j < dim) {
                    diff[j] = zs[j] - xsr[j];
                    r2 = fma(diff[j], diff[j], r2);
                    lp = fma(za[j], xr[j], lp);
                    ll = fma(zv[j], xr[j], ll);
                }
            }
            double stat, gg;
            if (kern == SEGP_KERN_LIN_RBF) {
                stat = var * exp(-0.5 * r2);
                gg = stat;
            } else {
                const double rr = sqrt(r2);
                const double e = var * exp(-sqrt5 * rr);
                stat = (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * e;
                gg = (5.0 / 3.0) * (1.0 + sqrt5 * rr) * e;
            }
            double kval = fma(lp, stat, ll);
            if (row >= a.n_train) kval = 0.0;
            const double bt = a.beta[(long)d * a.n_pad + row];   // zero on padded rows
            mu = fma(bt, kval, mu);
            const double w = bt * lp * gg, w2 = bt * stat;
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (j < dim) {
                    jac[j] = fma(w, diff[j], jac[j]);
                    jac2[j] = fma(w2, xr[j], jac2[j]);
                }
            kv[qd] = kval;
        }
        double2* dst = reinterpret_cast<double2*>(a.ks + (((long)d * ngroups + g) * a.b_cap + b) * 4);
        dst[0] = make_double2(kv[0], kv[1]);
        dst[1] = make_double2(kv[2], kv[3]);
    }
    const int n_s = gridDim.y;
    a.mu_part[((long)split * n_s + d) * a.b_cap + b] = mu;
#pragma unroll
    for (int j = 0; j < DM; ++j)
        if (j < dim) {
            a.jac_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = jac[j];
            a.jac2_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] =
                pl[j] * jac2[j] + (split == 0 ? lv[j] * a.xtb[d * dim + j] : 0.0);
        }
    if (split == 0) a.kss[(long)d * a.b_cap + b] = fma(var, kss, kss_lin);
}

template <int D_T>
__global__ void __launch_bounds__(KS_THREADS) kstar_mean_jac_kernel(const KstarArgs a) {
    constexpr int DM = D_T > 0 ? D_T : MAX_D;
    const int dim = D_T > 0 ? D_T : a.dim;
    const int d = blockIdx.y;
    const int split = blockIdx.z;
    const long b = (long)blockIdx.x * KS_THREADS + threadIdx.x;
    const bool active = b < a.n_batch;

    __shared__ double s_x[KS_THREADS * DM];
    __shared__ double s_beta[KS_THREADS];

    double zs[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) zs[j] = 0.0;
    if (active) {
        if (a.z != nullptr) {
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (j < dim) zs[j] = a.z[b * dim + j];
        } else {
            const double* p = a.p + b * a.p_stride;
            const double* u = a.kff + b * a.kff_stride;
            if (a.sp != nullptr && a.sp->has_t) {
                for (int i = 0; i < a.n_in; ++i) {
                    double acc = 0.0;
                    for (int k = 0; k < a.n_s_state; ++k) acc += a.sp->t[i * a.n_s_state + k] * p[k];
#pragma unroll
                    for (int j = 0; j < DM; ++j)
                        if (j == i) zs[j] = acc;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DM; ++j)
                    if (j < a.n_in) zs[j] = p[j];
            }
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (j >= a.n_in && j < dim) zs[j] = u[j - a.n_in];
        }
    }
    const int kern = a.kern[d];
    const double var = a.var[d];
    if (kern_is_composite(kern)) {   // block-uniform branch: the composite kernels take their own (simple) path
        kstar_composite<DM>(a, zs, dim, d, split, b, active);
        return;
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < DM; ++j)
            if (j < dim) zs[j] *= a.invls[d * dim + j];
    }

    double mu = 0.0;
    double jac[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) jac[j] = 0.0;

    const int ngroups = a.n_pad / 4;
    const int g0 = split * a.groups_per_split;
    const int g1 = min(g0 + a.groups_per_split, ngroups);
    const double sqrt5 = 2.23606797749978969641;

    for (int gbase = g0; gbase < g1; gbase += KS_THREADS / 4) {
        const int row0 = gbase * 4;
        const int nrows = min(KS_THREADS, (g1 - gbase) * 4);
        __syncthreads();
        const double* src = a.xs + ((long)d * a.n_pad + row0) * dim;
        for (int idx = threadIdx.x; idx < nrows * dim; idx += KS_THREADS) s_x[idx] = src[idx];
        if (threadIdx.x < nrows) s_beta[threadIdx.x] = a.beta[(long)d * a.n_pad + row0 + threadIdx.x];
        __syncthreads();
        if (!active) continue;
        for (int r = 0; r < nrows; r += 4) {
            double kv[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const double* xr = s_x + (r + qd) * dim;
                double diff[DM];
                double r2 = 0.0;
#pragma unroll
                for (int j = 0; j < DM; ++j) {
                    diff[j] = 0.0;
                    if (j < dim) {
                        diff[j] = zs[j] - xr[j];
                        r2 = fma(diff[j], diff[j], r2);
                    }
                }
                double kval, g;
                if (kern == SEGP_KERN_RBF) {
                    kval = var * exp(-0.5 * r2);
                    g = kval;
                } else {
                    const double rr = sqrt(r2);
                    const double e = var * exp(-sqrt5 * rr);
                    kval = (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * e;
                    g = (5.0 / 3.0) * (1.0 + sqrt5 * rr) * e;
                }
                if (row0 + r + qd >= a.n_train) kval = 0.0;   // padded rows (beta is zero there)
                const double bt = s_beta[r + qd];
                mu = fma(bt, kval, mu);
                const double w = bt * g;
#pragma unroll
                for (int j = 0; j < DM; ++j)
                    if (j < dim) jac[j] = fma(w, diff[j], jac[j]);
                kv[qd] = kval;
            }
            double2* dst = reinterpret_cast<double2*>(
                a.ks + (((long)d * ngroups + (row0 + r) / 4) * a.b_cap + b) * 4);
            dst[0] = make_double2(kv[0], kv[1]);
            dst[1] = make_double2(kv[2], kv[3]);
        }
    }
    if (active) {
        const int n_s = gridDim.y;
        a.mu_part[((long)split * n_s + d) * a.b_cap + b] = mu;
#pragma unroll
        for (int j = 0; j < DM; ++j)
            if (j < dim) {
                a.jac_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = jac[j];
                if (a.jac2_part != nullptr) a.jac2_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = 0.0;
            }
        if (a.kss != nullptr && split == 0) a.kss[(long)d * a.b_cap + b] = var;
    }
}

int launch_kstar(const KstarArgs& a, int n_s, int nsplit, cudaStream_t st) {
    dim3 grid((unsigned)((a.n_batch + KS_THREADS - 1) / KS_THREADS), (unsigned)n_s, (unsigned)nsplit);
    dim3 block(KS_THREADS);
    switch (a.dim) {
#define SEGP_KS_CASE(D) \
    case D:             \
        kstar_mean_jac_kernel<D><<<grid, block, 0, st>>>(a); \
        break;
        SEGP_KS_CASE(2)
        SEGP_KS_CASE(3)
        SEGP_KS_CASE(4)
        SEGP_KS_CASE(5)
        SEGP_KS_CASE(6)
        SEGP_KS_CASE(7)
        SEGP_KS_CASE(8)
        SEGP_KS_CASE(13)
#undef SEGP_KS_CASE
        default:
            kstar_mean_jac_kernel<0><<<grid, block, 0, st>>>(a);
    }
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== tri_sumsq
// CTA tile: 128 rows of V = W K* (one block row I of W) x 128 trajectories, K loop over block columns J <= I.
// 8 consumer warps (2 x 4, warp tile 64 x 32 = 8 x 4 DMMA tiles, 64 fp64 accumulators per thread) + 1 producer
// warp.  Operand tiles are stored in global memory already in the fragment order
//   A (W):   [k/4][row 0..127][k%4]      -> a warp's 8x4 fragment is 256 contiguous bytes
//   B (K*):  [k/4][col 0..127][k%4]      -> same
// so a k-chunk of 16 is one 16 KB bulk copy for A and four 4 KB bulk copies for B, and every fragment load
// is a conflict-free LDS.64.
constexpr int TRI_STAGES = 5;
constexpr int TRI_STAGE_DOUBLES = 2 * KC * TILE;   // A + B = 4096 doubles = 32 KB
constexpr int TRI_CONSUMER_WARPS = 8;
constexpr int TRI_THREADS = (TRI_CONSUMER_WARPS + 1) * 32;
constexpr size_t TRI_SMEM = (size_t)TRI_STAGES * TRI_STAGE_DOUBLES * 8 + 2 * TILE * 8 + 2 * TRI_STAGES * 8 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(TRI_THREADS, 1) tri_sumsq_kernel(const TriArgs a) {
    // ---- tile decode: (d, panel group) outer, block row I descending (heavy first), panel inner
    const int tiles_per_group = a.group * a.nblk;
    const int npg = (a.npanels + a.group - 1) / a.group;
    const long gid = blockIdx.x / tiles_per_group;
    const int r = blockIdx.x % tiles_per_group;
    const int d = (int)(gid / npg);
    const int pg = (int)(gid % npg);
    const int bi = a.nblk - 1 - r / a.group;
    const int panel = pg * a.group + r % a.group;
    if (panel >= a.npanels) return;
    const long b0 = (long)panel * TILE;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* s_col = stages + (size_t)TRI_STAGES * TRI_STAGE_DOUBLES;   // [2][128]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_col + 2 * TILE);
    uint64_t* empty_bar = full_bar + TRI_STAGES;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TRI_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], TRI_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nchunks = (bi + 1) * (TILE / KC);
    const long ngroups = (long)a.nblk * (TILE / 4);

    if (warp == TRI_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer (one elected lane)
        if (lane == 0) {
            const double* wbase = a.wt + ((long)d * a.ntri + (long)bi * (bi + 1) / 2) * (TILE * TILE);
            const double* kbase = a.ks + ((long)d * ngroups * a.b_cap + b0) * 4;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % TRI_STAGES;
                const int use = c / TRI_STAGES;
                if (use > 0) mbar_wait(&empty_bar[s], (uint32_t)((use - 1) & 1));
                double* dst = stages + (size_t)s * TRI_STAGE_DOUBLES;
                mbar_expect_tx(&full_bar[s], TRI_STAGE_DOUBLES * 8);
                // A: chunk c of the block row = 16 KB contiguous (blocks J are consecutive, 8 chunks each)
                bulk_g2s(dst, wbase + (long)c * (KC * TILE), KC * TILE * 8, &full_bar[s]);
                // B: 4 k-groups of 128 columns x 4
                const long g = (long)c * (KC / 4);
#pragma unroll
                for (int kg = 0; kg < KC / 4; ++kg)
                    bulk_g2s(dst + KC * TILE + kg * (TILE * 4), kbase + (g + kg) * a.b_cap * 4, TILE * 4 * 8,
                             &full_bar[s]);
            }
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int row0 = (warp >> 2) * 64;
        const int col0 = (warp & 3) * 32;
        const int frag = (lane >> 2) * 4 + (lane & 3);
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int c = 0; c < nchunks; ++c) {
            const int s = c % TRI_STAGES;
            mbar_wait(&full_bar[s], (uint32_t)((c / TRI_STAGES) & 1));
            const double* as = stages + (size_t)s * TRI_STAGE_DOUBLES;
            const double* bs = as + KC * TILE;
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                double af[8], bf[4];
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) af[mt] = as[(kk * TILE + row0 + mt * 8) * 4 + frag];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) bf[nt] = bs[(kk * TILE + col0 + nt * 8) * 4 + frag];
#pragma unroll
                for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        // ---- epilogue: column sums of squares of the 64 x 32 warp tile
        // thread holds rows mt*8 + lane/4, columns nt*8 + (lane%4)*2 + {0,1}
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double sacc = 0.0;
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) sacc = fma(acc[mt][nt][i], acc[mt][nt][i], sacc);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
                if (lane < 4) s_col[(warp >> 2) * TILE + col0 + nt * 8 + lane * 2 + i] = sacc;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < TILE)
        a.qpart[((long)d * a.nblk + bi) * a.b_cap + b0 + threadIdx.x] = s_col[threadIdx.x] + s_col[TILE + threadIdx.x];
}

int tri_sumsq_init() {
    SEGP_CUDA_CHECK(cudaFuncSetAttribute(tri_sumsq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_SMEM));
    return SEGP_OK;
}

int launch_tri_sumsq(const TriArgs& a, int n_s, cudaStream_t st) {
    const int npg = (a.npanels + a.group - 1) / a.group;
    const long nblocks = (long)n_s * npg * a.group * a.nblk;
    if (nblocks <= 0 || nblocks > 2147483647L) {
        set_error("tri_sumsq: grid of %ld tiles out of range", nblocks);
        return SEGP_ERR_INVALID;
    }
    tri_sumsq_kernel<<<(unsigned)nblocks, TRI_THREADS, TRI_SMEM, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== finalize (predict)
__global__ void finalize_predict_kernel(const FinalizeArgs a) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_batch) return;
    int32_t status = 0;
    for (int d = 0; d < a.n_s; ++d) {
        double mu = 0.0;
#pragma unroll 8
        for (int s = 0; s < a.nsplit; ++s) mu += a.mu_part[((long)s * a.n_s + d) * a.b_cap + b];
        double qf = 0.0;
#pragma unroll 8
        for (int i = 0; i < a.nblk; ++i) qf += a.qpart[((long)d * a.nblk + i) * a.b_cap + b];
        a.mu[b * a.n_s + d] = mu;
        const double var = (a.kss != nullptr ? a.kss[(long)d * a.b_cap + b] : a.gp_var[d]) - qf;
        a.var[b * a.n_s + d] = var;
        if (!(var > 0.0)) status |= SEGP_STATUS_BAD_VARIANCE;
        if (a.epart != nullptr) {
            float e2 = 0.f;
            for (int i = 0; i < a.nblk; ++i) e2 += a.epart[((long)d * a.nblk + i) * a.b_cap + b];
            if (a.guard_gs * (double)e2 > var * var) status |= SEGP_STATUS_LOW_PRECISION;
        }
        if (a.jac != nullptr) {
            for (int j = 0; j < a.dim; ++j) {
                double acc = 0.0, add = 0.0;
                for (int s = 0; s < a.nsplit; ++s) {
                    const long idx = (((long)s * a.n_s + d) * a.dim + j) * a.b_cap + b;
                    acc += a.jac_part[idx];
                    if (a.jac2_part != nullptr) add += a.jac2_part[idx];
                }
                a.jac[(b * a.n_s + d) * a.dim + j] = fma(-acc, a.invls[d * a.dim + j], add);
            }
        }
    }
    if (a.status != nullptr) a.status[b] = status;
}

int launch_finalize_predict(const FinalizeArgs& a, cudaStream_t st) {
    const int threads = 128;
    finalize_predict_kernel<<<(unsigned)((a.n_batch + threads - 1) / threads), threads, 0, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

}  // namespace segp
