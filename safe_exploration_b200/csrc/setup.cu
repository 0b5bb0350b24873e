// Model setup on the device, all float64 (runs once per model update, off the rollout loop):
//   kmat        K_d = k_d(X,X) + noise_d I                    (ssm_gpy/gaussian_process.py:247 GPRegression)
//   potrf_lower blocked right-looking Cholesky, 64-wide panels, trailing update on the FP64 tensor pipe (DMMA)
//   trtri_lower W_d = L_d^-1 by recursive doubling: [[W11,0],[-W22 L21 W11, W22]]   (GPy posterior.woodbury_inv
//               is W^T W; ssm_gpy/gaussian_process.py:258-259, 406-409 pdinv)
//   solve_beta  beta_d = W^T (W y_d)                          (posterior.woodbury_vector, :261)
//   pack_w      W -> 128x128 tiles in DMMA fragment order for tri_sumsq
//   logdet      2 sum log L_ii                                (information_gain, :621-634)
// The factorisation must be float64: cond(K + noise I) ~ N s_f^2 / noise ~ 1e6..1e8 at the benchmark sizes.
#include <math.h>

#include <algorithm>

#include "segp_internal.cuh"

namespace segp {

// =========================================================================================== kmat
__global__ void kmat_kernel(double* __restrict__ k, const double* __restrict__ xs, int kern, double var, double noise,
                            int n_train, int n_pad, int dim, const double* __restrict__ xraw,
                            const double* __restrict__ plin, const double* __restrict__ lin, int row0) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = row0 + blockIdx.y;   // row of the kernel matrix; row blockIdx.y of the output buffer
    if (j >= n_pad) return;
    const bool composite = kern_is_composite(kern);
    double v;
    if (i >= n_train || j >= n_train) {
        v = (i == j) ? 1.0 : 0.0;
    } else if (i == j && !composite) {
        v = var + noise;
    } else {
        double r2 = 0.0;
        for (int c = 0; c < dim; ++c) {
            const double df = xs[(long)i * dim + c] - xs[(long)j * dim + c];
            r2 = fma(df, df, r2);
        }
        if (kern == SEGP_KERN_RBF || kern == SEGP_KERN_LIN_RBF) {
            v = var * exp(-0.5 * r2);
        } else {
            const double sqrt5 = 2.23606797749978969641;
            const double rr = sqrt(r2);
            v = var * (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * exp(-sqrt5 * rr);
        }
        if (composite) {   // (sum a x_i x_j) k_stat + sum v x_i x_j   (_k_lin_rbf / _k_lin_mat52)
            double lp = 0.0, ll = 0.0;
            for (int c = 0; c < dim; ++c) {
                const double xx = xraw[(long)i * dim + c] * xraw[(long)j * dim + c];
                lp = fma(plin[c], xx, lp);
                ll = fma(lin[c], xx, ll);
            }
            v = fma(lp, v, ll);
            if (i == j) v += noise;
        }
    }
    k[(long)blockIdx.y * n_pad + j] = v;
}

int launch_kmat_rows(double* k, const double* xs_d, int kern, double var, double noise, SetupDims s, const double* xraw,
                     const double* plin_d, const double* lin_d, int row0, int nrows, cudaStream_t st) {
    if (kern_is_composite(kern) && (xraw == nullptr || plin_d == nullptr || lin_d == nullptr)) {
        set_error("composite kernel without linear terms (call segp_set_linear_terms)");
        return SEGP_ERR_INVALID;
    }
    dim3 grid((unsigned)((s.n_pad + 127) / 128), (unsigned)nrows);
    kmat_kernel<<<grid, 128, 0, st>>>(k, xs_d, kern, var, noise, s.n_train, s.n_pad, s.dim, xraw, plin_d, lin_d, row0);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

int launch_kmat(double* k, const double* xs_d, int kern, double var, double noise, SetupDims s, const double* xraw,
                const double* plin_d, const double* lin_d, cudaStream_t st) {
    return launch_kmat_rows(k, xs_d, kern, var, noise, s, xraw, plin_d, lin_d, 0, s.n_pad, st);
}

// xtb[j] = sum_i beta[i] xraw[i][j]: one block per input dimension, fixed-order tree reduction
__global__ void __launch_bounds__(256) xtb_kernel(const double* __restrict__ xraw, const double* __restrict__ beta,
                                                  double* __restrict__ xtb, int n_pad, int dim) {
    __shared__ double s[256];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_pad; i += 256) acc = fma(beta[i], xraw[(long)i * dim + j], acc);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) xtb[j] = s[0];
}

int launch_xtb(const double* xraw, const double* beta, double* xtb, int n_pad, int dim, cudaStream_t st) {
    xtb_kernel<<<dim, 256, 0, st>>>(xraw, beta, xtb, n_pad, dim);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== gemm64 (DMMA)
// C[z] (M_z x N) = alpha * A[z] (M_z x K) * op(B[z]) + beta * C[z];  all row-major, all dims multiples of 64.
//   TRANS_B: B is N x K (C = A B^T), else B is K x N.
//   flags  GEMM_A_LOWER: A[i,k] == 0 for k > i   -> k loop stops at the row tile's end
//          GEMM_B_LOWER: B[k,n] == 0 for k < n   -> k loop starts at the column tile's start (non-transposed B)
//          GEMM_C_LOWER: only tiles on or below the diagonal are computed (syrk)
//   batch  blockIdx.z: every pointer advances by z * zstride elements; M_z = min(M, m_total - z * zrows).
// 64x64 tile, 4 warps (2x2) of 32x32 = 4x4 DMMA m8n8k4 tiles; operands staged in shared memory in fragment order
// [k/4][row][k%4] so every fragment is one conflict-free LDS.64; global loads are register-prefetched one chunk ahead.
constexpr int GT = 64;    // tile edge
constexpr int GK = 16;    // k chunk

struct GemmArgs {
    const double* a;
    const double* b;
    double* c;
    long lda, ldb, ldc;
    int m, n, k;
    double alpha, beta;
    int flags;
    long zstride_a, zstride_b, zstride_c;
    int m_total, zrows;
};

__device__ __forceinline__ void dmma884_s(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool TRANS_B>
__global__ void __launch_bounds__(128) gemm64_kernel(const GemmArgs g) {
    const int z = blockIdx.z;
    const int m_z = min(g.m, g.m_total - z * g.zrows);
    const int row0 = blockIdx.y * GT;
    const int col0 = blockIdx.x * GT;
    if (row0 >= m_z) return;
    if ((g.flags & GEMM_C_LOWER) && col0 > row0) return;
    const double* __restrict__ A = g.a + (long)z * g.zstride_a;
    const double* __restrict__ B = g.b + (long)z * g.zstride_b;
    double* __restrict__ C = g.c + (long)z * g.zstride_c;

    int k_begin = 0, k_end = g.k;
    if (g.flags & GEMM_A_LOWER) k_end = min(k_end, row0 + GT);
    if (g.flags & GEMM_B_LOWER) k_begin = col0;

    __shared__ __align__(16) double s_a[GK * GT];   // [k/4][row][k%4]
    __shared__ __align__(16) double s_b[GK * GT];   // [k/4][col][k%4]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
    const int frag = (lane >> 2) * 4 + (lane & 3);

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // register staging of one chunk: A: 2 x (row, kgroup) items of 4 doubles; B likewise
    double ra[2][4], rb[2][4];
    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int item = tid + it * 128;        // 0..255
            const int r = item >> 2, kg = item & 3;
            const double2* src = reinterpret_cast<const double2*>(A + (long)(row0 + r) * g.lda + k0 + kg * 4);
            const double2 v0 = src[0], v1 = src[1];
            ra[it][0] = v0.x; ra[it][1] = v0.y; ra[it][2] = v1.x; ra[it][3] = v1.y;
        }
        if (TRANS_B) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int item = tid + it * 128;
                const int r = item >> 2, kg = item & 3;
                const double2* src = reinterpret_cast<const double2*>(B + (long)(col0 + r) * g.ldb + k0 + kg * 4);
                const double2 v0 = src[0], v1 = src[1];
                rb[it][0] = v0.x; rb[it][1] = v0.y; rb[it][2] = v1.x; rb[it][3] = v1.y;
            }
        } else {
            // item = (kgroup, column pair): 4 x 32 = 128 items, one per thread; two columns x four k each
            const int kg = tid >> 5, cp = tid & 31;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const double2 v =
                    *reinterpret_cast<const double2*>(B + (long)(k0 + kg * 4 + kk) * g.ldb + col0 + cp * 2);
                rb[0][kk] = v.x;
                rb[1][kk] = v.y;
            }
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int item = tid + it * 128;
            const int r = item >> 2, kg = item & 3;
            double2* dst = reinterpret_cast<double2*>(s_a + (kg * GT + r) * 4);
            dst[0] = make_double2(ra[it][0], ra[it][1]);
            dst[1] = make_double2(ra[it][2], ra[it][3]);
        }
        if (TRANS_B) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int item = tid + it * 128;
                const int r = item >> 2, kg = item & 3;
                double2* dst = reinterpret_cast<double2*>(s_b + (kg * GT + r) * 4);
                dst[0] = make_double2(rb[it][0], rb[it][1]);
                dst[1] = make_double2(rb[it][2], rb[it][3]);
            }
        } else {
            const int kg = tid >> 5, cp = tid & 31;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                double2* dst = reinterpret_cast<double2*>(s_b + (kg * GT + cp * 2 + cc) * 4);
                dst[0] = make_double2(rb[cc][0], rb[cc][1]);
                dst[1] = make_double2(rb[cc][2], rb[cc][3]);
            }
        }
    };

    if (k_begin < k_end) load_chunk(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += GK) {
        store_chunk();
        __syncthreads();
        if (k0 + GK < k_end) load_chunk(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK / 4; ++kk) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = s_a[(kk * GT + wr + i * 8) * 4 + frag];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = s_b[(kk * GT + wc + j * 8) * 4 + frag];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884_s(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    // epilogue: thread holds rows wr + i*8 + lane/4, columns wc + j*8 + (lane%4)*2 + {0,1}
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + wr + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cidx = col0 + wc + j * 8 + (lane & 3) * 2;
            double2* dst = reinterpret_cast<double2*>(C + (long)r * g.ldc + cidx);
            double2 out = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
            if (g.beta != 0.0) {
                const double2 old = *dst;
                out.x = fma(g.beta, old.x, out.x);
                out.y = fma(g.beta, old.y, out.y);
            }
            *dst = out;
        }
    }
}

static int launch_gemm64(const GemmArgs& g, bool trans_b, int batch, cudaStream_t st) {
    if (g.m <= 0 || g.n <= 0 || batch <= 0) return SEGP_OK;
    dim3 grid((unsigned)(g.n / GT), (unsigned)(g.m / GT), (unsigned)batch);
    if (trans_b)
        gemm64_kernel<true><<<grid, 128, 0, st>>>(g);
    else
        gemm64_kernel<false><<<grid, 128, 0, st>>>(g);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== potrf
// Diagonal block: Cholesky of the 64 x 64 block + its explicit inverse.  This kernel is one CTA on the critical path
// of the whole factorisation (n_pad / 64 of them in a row), so it is written for latency:
//   factor   right-looking with the matrix in REGISTERS: thread (row, part) owns the entries (row, 4 i + part), i < 16,
//            of its row; per column j two barriers -- the owner of (j, j) publishes the pivot; every thread takes its
//            reciprocal square root (MUFU + two Newton steps), the owners of column j publish the scaled column; then
//            16 independent FMAs per thread against broadcast shared-memory reads.  The column loop is fully unrolled
//            so that the register indices are static.
//   inverse  the two 32 x 32 diagonal blocks side by side (row by row, 4 lanes per entry + shuffle reduction), then
//            X21 = -X22 (L21 X11) as two 32^3 products over all 256 threads.
// (round 1: 77 us per block -- three barriers per column around shared-memory updates, one-thread-per-column
// substitution; the shared-memory version of this layout with one barrier: 61 us.)
__global__ void __launch_bounds__(256) potf2_inv_kernel(double* __restrict__ a, int n_pad, int kb,
                                                        double* __restrict__ diag_inv, int* __restrict__ fail) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double(*s)[NBLK + 1] = reinterpret_cast<double(*)[NBLK + 1]>(dyn_smem);   // T of the inverse
    double(*l)[NBLK + 1] = s + NBLK;                                          // L
    double(*x)[NBLK + 1] = l + NBLK;                                          // L^-1
    __shared__ double s_dinv[NBLK], s_lcol[NBLK], s_piv;
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    const int row = tid & 63, part = tid >> 6;
    double* blk = a + ((long)kb * NBLK) * n_pad + (long)kb * NBLK;
    if (tid == 0) s_fail = 0;
    // coalesced load through shared memory, then each thread takes its 16 entries
    for (int idx = tid; idx < NBLK * NBLK; idx += 256) {
        const int r = idx / NBLK, c = idx % NBLK;
        s[r][c] = (c <= r) ? blk[(long)r * n_pad + c] : 0.0;
        l[r][c] = 0.0;
        x[r][c] = 0.0;
    }
    __syncthreads();
    double av[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) av[i] = s[row][4 * i + part];
    if (tid == 0) s_piv = s[0][0];
#pragma unroll
    for (int j = 0; j < NBLK; ++j) {
        const int pj = j & 3, ij = j >> 2;
        __syncthreads();                              // pivot (j, j) published
        const double d = s_piv;
        const bool bad = !(d > 0.0);
        if (bad && tid == 0) s_fail = kb * NBLK + j + 1;
        const double rinv = bad ? 1.0 : rsqrt_fast(d);
        if (part == pj && row >= j) {
            const double lij = (row == j) ? (bad ? 1.0 : d * rinv) : av[ij] * rinv;
            s_lcol[row] = lij;
            l[row][j] = lij;
            if (row == j) s_dinv[j] = rinv;
        }
        __syncthreads();                              // scaled column j published
        if (row > j) {
            const double lrow = s_lcol[row];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = 4 * i + part;
                if (k > j && k <= row) av[i] = fma(-lrow, s_lcol[k], av[i]);
            }
            if (j + 1 < NBLK && row == j + 1 && part == ((j + 1) & 3)) s_piv = av[(j + 1) >> 2];
        }
    }
    __syncthreads();
    // inverse of the two diagonal 32 x 32 blocks: group g = tid / 128, lane quad = 4 partial sums of one column
    {
        const int g = tid >> 7, t = tid & 127;
        const int c = t >> 2, p4 = t & 3;            // column within the block, partial-sum index
        const int o = g * 32;
        for (int i = 0; i < 32; ++i) {
            double acc = 0.0;
            for (int k = c + p4; k < i; k += 4) acc = fma(l[o + i][o + k], x[o + k][o + c], acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (p4 == 0 && c <= i) x[o + i][o + c] = (c == i) ? s_dinv[o + i] : -acc * s_dinv[o + i];
            __syncthreads();
        }
    }
    // T = L21 X11 (X11 lower: k >= c) into s[0..31][0..31]; then X21 = -X22 T (X22 lower: k <= r)
    for (int idx = tid; idx < 32 * 32; idx += 256) {
        const int r = idx >> 5, c = idx & 31;
        double acc = 0.0;
        for (int k = c; k < 32; ++k) acc = fma(l[32 + r][k], x[k][c], acc);
        s[r][c] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < 32 * 32; idx += 256) {
        const int r = idx >> 5, c = idx & 31;
        double acc = 0.0;
        for (int k = 0; k <= r; ++k) acc = fma(x[32 + r][32 + k], s[k][c], acc);
        x[32 + r][c] = -acc;
    }
    __syncthreads();
    for (int idx = tid; idx < NBLK * NBLK; idx += 256) {
        const int r = idx / NBLK, c = idx % NBLK;
        if (c <= r) blk[(long)r * n_pad + c] = l[r][c];
        diag_inv[(long)kb * NBLK * NBLK + idx] = x[r][c];
    }
    if (tid == 0 && s_fail != 0) atomicCAS(fail, 0, s_fail);
}

// Panel: A[i,k] <- A[i,k] * Linv_kk^T for the row blocks i > k (64 rows per CTA).  Thread (c, rg) computes column c of
// the rows rg * 16 .. + 16, eight rows at a time: eight independent accumulators against one broadcast row of Linv.
__global__ void __launch_bounds__(256) panel_trsm_kernel(double* __restrict__ a, int n_pad, int kb,
                                                         const double* __restrict__ diag_inv) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double(*s_a)[NBLK + 1] = reinterpret_cast<double(*)[NBLK + 1]>(dyn_smem);
    double(*s_l)[NBLK + 1] = s_a + NBLK;
    const int tid = threadIdx.x;
    const long row0 = ((long)kb + 1 + blockIdx.x) * NBLK;
    double* blk = a + row0 * n_pad + (long)kb * NBLK;
    const double* li = diag_inv + (long)kb * NBLK * NBLK;
    for (int idx = tid; idx < NBLK * NBLK; idx += 256) {
        const int r = idx / NBLK, c = idx % NBLK;
        s_a[r][c] = blk[(long)r * n_pad + c];
        s_l[r][c] = li[idx];
    }
    __syncthreads();
    // out[r][c] = sum_{m <= c} A[r][m] * Linv[c][m]
    const int c = tid & 63, rg = tid >> 6;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r0 = rg * 16 + half * 8;
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
        for (int m = 0; m <= c; ++m) {
            const double lv = s_l[c][m];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(s_a[r0 + i][m], lv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) blk[(long)(r0 + i) * n_pad + c] = acc[i];
    }
}

// Two-level right-looking blocking when the tensor-core GEMM is available (fd != NULL): 64-wide blocks are factorised
// as before (potf2_inv, panel_trsm) but their float64 DMMA updates stay inside the current 256-wide panel; the rest of
// the trailing matrix is updated once per panel, A22 -= P P^T with K = 256, on tcgen05 from digit planes of P
// (fact_i8.cu): ~1 - 1.5 * 256 / n_pad of the factorisation's flops run on the tensor cores.
int potrf_lower(double* a, int n_pad, double* diag_inv, int* d_fail, cudaStream_t st, long* launches,
                const FdScratch* fd) {
    const int nb = n_pad / NBLK;
    constexpr int kBlockSmem = 2 * NBLK * (NBLK + 1) * (int)sizeof(double);
    constexpr int kPotf2Smem = 3 * NBLK * (NBLK + 1) * (int)sizeof(double);
    constexpr int PANEL = 256;
    SEGP_CUDA_CHECK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPotf2Smem));
    SEGP_CUDA_CHECK(cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlockSmem));
    SEGP_CUDA_CHECK(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
    for (int kb = 0; kb < nb; ++kb) {
        potf2_inv_kernel<<<1, 256, kPotf2Smem, st>>>(a, n_pad, kb, diag_inv, d_fail);
        ++*launches;
        const int rem = nb - kb - 1;
        if (rem <= 0) break;
        panel_trsm_kernel<<<rem, 256, kBlockSmem, st>>>(a, n_pad, kb, diag_inv);
        ++*launches;
        // trailing update: A22 -= A21 A21^T (lower tiles only); with fd only up to the end of the 256-wide panel
        GemmArgs g{};
        const long off = ((long)kb + 1) * NBLK;
        const long panel_end = std::min<long>(((long)kb * NBLK / PANEL + 1) * PANEL, n_pad);
        g.a = a + off * n_pad + (long)kb * NBLK;
        g.b = g.a;
        g.c = a + off * n_pad + off;
        g.lda = g.ldb = g.ldc = n_pad;
        g.m = rem * NBLK;
        g.n = fd != nullptr ? (int)(panel_end - off) : g.m;
        g.k = NBLK;
        g.alpha = -1.0;
        g.beta = 1.0;
        g.flags = GEMM_C_LOWER;
        g.m_total = g.m;
        g.zrows = 0;
        if (g.n > 0) {
            SEGP_CHECK(launch_gemm64(g, true, 1, st));
            ++*launches;
        }
        if (fd != nullptr && off == panel_end && off < n_pad) {
            // the panel [p0, off) is final below row off: digit planes of P = A[off:, p0:off], then the K = pw update
            const long p0 = panel_end - PANEL < 0 ? 0 : ((long)kb * NBLK / PANEL) * PANEL;
            const int pw = (int)(off - p0);
            const int rows = (int)(n_pad - off);
            SEGP_CHECK(fd_split(a + off * n_pad + p0, n_pad, 0, rows, pw, 0, 0, 0, 0, 0, 1, fd->as, fd->ap, st));
            GemmI8Args t{};
            t.ap = t.bp = fd->ap;
            t.as = t.bs = fd->as;
            t.a_kb = t.b_kb = pw / NBLK;
            t.c = a + off * n_pad + off;
            t.ldc = n_pad;
            t.m = t.n = rows;
            t.k = pw;
            t.alpha = -1.0;
            t.beta = 1.0;
            t.flags = GEMM_C_LOWER;
            t.m_total = rows;
            t.zrows = 0;
            SEGP_CHECK(launch_gemm_i8d(t, 1, st));
            *launches += 4;
        }
    }
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== trtri
__global__ void copy_diag_inv_kernel(double* __restrict__ w, int n_pad, const double* __restrict__ diag_inv) {
    const int kb = blockIdx.x;
    double* blk = w + ((long)kb * NBLK) * n_pad + (long)kb * NBLK;
    for (int idx = threadIdx.x; idx < NBLK * NBLK; idx += blockDim.x)
        blk[(long)(idx / NBLK) * n_pad + (idx % NBLK)] = diag_inv[(long)kb * NBLK * NBLK + idx];
}

int trtri_lower(const double* l, double* w, int n_pad, const double* diag_inv, double* tmp, cudaStream_t st,
                long* launches, const FdScratch* fd) {
    const int nb = n_pad / NBLK;
    copy_diag_inv_kernel<<<nb, 256, 0, st>>>(w, n_pad, diag_inv);
    ++*launches;
    // level with block size s merges diagonal blocks [r0, r0+s) and [r0+s, r0+2s) (clipped at n_pad):
    //   T   = L21 W11      (W11 lower)       -> tmp, at the coordinates of the (2,1) block
    //   W21 = -W22 T       (W22 lower)
    for (long s = NBLK; s < n_pad; s *= 2) {
        const int pairs = (int)((n_pad - s + 2 * s - 1) / (2 * s));   // pairs whose second block is non-empty
        if (fd != nullptr && s >= 256) {
            // the same two products from digit planes on tcgen05 (fact_i8.cu); every operand is split with the exact
            // max-abs of its rows over this level's k range as the scale
            const long zs = 2 * s * ((long)n_pad + 1);
            const int lim = (int)(n_pad - s), zr = (int)(2 * s), si = (int)s;
            const long zplanes = (long)fd_plane_bytes(si, si, 1);
            GemmI8Args t{};
            t.ap = fd->ap;
            t.as = fd->as;
            t.bp = fd->bp;
            t.bs = fd->bs;
            t.a_kb = t.b_kb = si / NBLK;
            t.ldc = n_pad;
            t.m = t.n = t.k = si;
            t.z_ap = t.z_bp = zplanes;
            t.z_as = t.z_bs = s;
            t.z_c = zs;
            t.m_total = lim;
            t.zrows = zr;
            // T = L21 * W11: A = L21 (rows clipped in the last pair), B^T = W11^T (zero for k < n)
            SEGP_CHECK(fd_split(l + s * n_pad, n_pad, zs, si, si, 0, 0, 1, lim, zr, pairs, fd->as, fd->ap, st));
            SEGP_CHECK(fd_split(w, n_pad, zs, si, si, 1, 2, 0, 0, 0, pairs, fd->bs, fd->bp, st));
            t.c = tmp + s * n_pad;
            t.alpha = 1.0;
            t.beta = 0.0;
            t.flags = GEMM_B_LOWER;
            SEGP_CHECK(launch_gemm_i8d(t, pairs, st));
            // W21 = -W22 * T: A = W22 (lower; rows and k clipped), B^T = T^T (k clipped)
            SEGP_CHECK(fd_split(w + s * n_pad + s, n_pad, zs, si, si, 0, 1, 3, lim, zr, pairs, fd->as, fd->ap, st));
            SEGP_CHECK(fd_split(tmp + s * n_pad, n_pad, zs, si, si, 1, 0, 2, lim, zr, pairs, fd->bs, fd->bp, st));
            t.c = w + s * n_pad;
            t.alpha = -1.0;
            t.beta = 0.0;
            t.flags = GEMM_A_LOWER;
            SEGP_CHECK(launch_gemm_i8d(t, pairs, st));
            *launches += 14;
            continue;
        }
        GemmArgs g{};
        g.lda = g.ldb = g.ldc = n_pad;
        g.m = (int)s;
        g.n = (int)s;
        g.k = (int)s;
        g.zstride_a = g.zstride_b = g.zstride_c = 2 * s * ((long)n_pad + 1);
        g.m_total = (int)(n_pad - s);
        g.zrows = (int)(2 * s);
        // T = L21 * W11
        g.a = l + s * n_pad;
        g.b = w;
        g.c = tmp + s * n_pad;
        g.alpha = 1.0;
        g.beta = 0.0;
        g.flags = GEMM_B_LOWER;
        SEGP_CHECK(launch_gemm64(g, false, pairs, st));
        ++*launches;
        // W21 = -W22 * T   (K = rows of the second block; the A_LOWER trim bounds the k loop by the row tile)
        g.a = w + s * n_pad + s;
        g.b = tmp + s * n_pad;
        g.c = w + s * n_pad;
        g.alpha = -1.0;
        g.beta = 0.0;
        g.flags = GEMM_A_LOWER;
        SEGP_CHECK(launch_gemm64(g, false, pairs, st));
        ++*launches;
    }
    return SEGP_OK;
}

// =========================================================================================== beta = W^T (W y)
// u = W y : one warp per row (W lower: columns 0..row)
__global__ void wy_kernel(const double* __restrict__ w, const double* __restrict__ y, double* __restrict__ u,
                          int n_pad) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    double acc = 0.0;
    const double* wr = w + (long)row * n_pad;
    for (int c = lane; c <= row; c += 32) acc = fma(wr[c], y[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) u[row] = acc;
}
// beta = W^T u : one thread per column, rows split over blockIdx.y with a fixed-order second pass
__global__ void wtu_partial_kernel(const double* __restrict__ w, const double* __restrict__ u,
                                   double* __restrict__ part, int n_pad, int rows_per_split) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_pad) return;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(n_pad, r0 + rows_per_split);
    double acc = 0.0;
    for (int r = max(r0, c); r < r1; ++r) acc = fma(w[(long)r * n_pad + c], u[r], acc);
    part[(long)blockIdx.y * n_pad + c] = acc;
}
__global__ void wtu_reduce_kernel(const double* __restrict__ part, double* __restrict__ beta, int n_pad, int nsplit) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_pad) return;
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += part[(long)s * n_pad + c];
    beta[c] = acc;
}

int solve_beta(const double* w, const double* y, double* u_tmp, double* beta, int n_pad, cudaStream_t st) {
    // u_tmp: (1 + nsplit) * n_pad doubles
    const int nsplit = 32;
    const int rows_per_split = (n_pad + nsplit - 1) / nsplit;
    wy_kernel<<<(n_pad + 7) / 8, 256, 0, st>>>(w, y, u_tmp, n_pad);
    dim3 grid((unsigned)((n_pad + 127) / 128), nsplit);
    wtu_partial_kernel<<<grid, 128, 0, st>>>(w, u_tmp, u_tmp + n_pad, n_pad, rows_per_split);
    wtu_reduce_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(u_tmp + n_pad, beta, n_pad, nsplit);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== pack_w
// wt[tile(bi,bj)][k/4][row][k%4], tile index bi (bi+1)/2 + bj for bj <= bi; k indexes the columns of W.
__global__ void pack_w_kernel(const double* __restrict__ w, double* __restrict__ wt, int n_pad) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const long tile = (long)bi * (bi + 1) / 2 + bj;
    double* dst = wt + tile * (TILE * TILE);
    const double* src = w + ((long)bi * TILE) * n_pad + (long)bj * TILE;
    for (int item = threadIdx.x; item < TILE * (TILE / 4); item += blockDim.x) {
        const int kg = item & 31;      // fastest over k-groups: coalesced reads of a W row
        const int r = item >> 5;
        const double2* s2 = reinterpret_cast<const double2*>(src + (long)r * n_pad + kg * 4);
        double2* d2 = reinterpret_cast<double2*>(dst + ((long)kg * TILE + r) * 4);
        d2[0] = s2[0];
        d2[1] = s2[1];
    }
}

int pack_w(const double* w, double* wt, int n_pad, cudaStream_t st) {
    const int nblk = n_pad / TILE;
    dim3 grid((unsigned)nblk, (unsigned)nblk);
    pack_w_kernel<<<grid, 256, 0, st>>>(w, wt, n_pad);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== logdet
__global__ void logdet_kernel(const double* __restrict__ l, int n_train, int n_pad, double* __restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_train; i += 256) acc += log(l[(long)i * n_pad + i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = 2.0 * s[0];
}

int logdet_from_chol(const double* l, int n_train, int n_pad, double* d_out, cudaStream_t st) {
    logdet_kernel<<<1, 256, 0, st>>>(l, n_train, n_pad, d_out);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// the same from W = L^-1 (W_ii = 1 / L_ii): -2 sum log W_ii
__global__ void logdet_winv_kernel(const double* __restrict__ w, int n_train, int n_pad, double* __restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_train; i += 256) acc += log(w[(long)i * n_pad + i]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = -2.0 * s[0];
}

int logdet_from_winv(const double* w, int n_train, int n_pad, double* d_out, cudaStream_t st) {
    logdet_winv_kernel<<<1, 256, 0, st>>>(w, n_train, n_pad, d_out);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== append_rows
// Rows [r0, r0 + s) of K + noise I are new (appended training points, re-evaluated old rows of the same 64-row blocks,
// identity rows of the padding); rows [0, r0) and their part W11 of W = L^-1 are unchanged.  With
//   L21 = K21 W11^T,   S = K22 - L21 L21^T = L22 L22^T,   W22 = L22^-1,   W21 = -W22 (L21 W11)
// the rows [r0, r0 + s) of W follow in O(s N^2) instead of O(N^3) (the block form trtri_lower uses level by level).
// r0 and s are multiples of 64; rows below r0 + s must be padding (identity).  krows: s x n_pad, rows of K + noise I.
int append_rows(double* w, int n_pad, int r0, int s, const double* krows, double* l21, double* tbuf, double* sbuf,
                double* w22, double* tmp, double* diag_inv, int* d_fail, cudaStream_t st, long* launches) {
    SEGP_CUDA_CHECK(cudaMemcpy2DAsync(sbuf, (size_t)s * sizeof(double), krows + r0, (size_t)n_pad * sizeof(double),
                                      (size_t)s * sizeof(double), (size_t)s, cudaMemcpyDeviceToDevice, st));
    GemmArgs g{};
    g.m_total = s;
    g.zrows = 0;
    if (r0 > 0) {
        // L21 = K21 W11^T
        g.a = krows;
        g.b = w;
        g.c = l21;
        g.lda = g.ldb = g.ldc = n_pad;
        g.m = s;
        g.n = r0;
        g.k = r0;
        g.alpha = 1.0;
        g.beta = 0.0;
        g.flags = 0;
        SEGP_CHECK(launch_gemm64(g, true, 1, st));
        // S = K22 - L21 L21^T
        g.a = l21;
        g.b = l21;
        g.c = sbuf;
        g.ldc = s;
        g.m = s;
        g.n = s;
        g.k = r0;
        g.alpha = -1.0;
        g.beta = 1.0;
        SEGP_CHECK(launch_gemm64(g, true, 1, st));
        *launches += 2;
    }
    SEGP_CHECK(potrf_lower(sbuf, s, diag_inv, d_fail, st, launches));
    SEGP_CUDA_CHECK(cudaMemsetAsync(w22, 0, (size_t)s * s * sizeof(double), st));
    SEGP_CHECK(trtri_lower(sbuf, w22, s, diag_inv, tmp, st, launches));
    if (r0 > 0) {
        // T = L21 W11 ; W21 = -W22 T
        g.a = l21;
        g.b = w;
        g.c = tbuf;
        g.lda = g.ldb = g.ldc = n_pad;
        g.m = s;
        g.n = r0;
        g.k = r0;
        g.alpha = 1.0;
        g.beta = 0.0;
        g.flags = GEMM_B_LOWER;
        SEGP_CHECK(launch_gemm64(g, false, 1, st));
        g.a = w22;
        g.lda = s;
        g.b = tbuf;
        g.ldb = n_pad;
        g.c = w + (long)r0 * n_pad;
        g.ldc = n_pad;
        g.m = s;
        g.n = r0;
        g.k = s;
        g.alpha = -1.0;
        g.beta = 0.0;
        g.flags = GEMM_A_LOWER;
        SEGP_CHECK(launch_gemm64(g, false, 1, st));
        *launches += 2;
    }
    SEGP_CUDA_CHECK(cudaMemcpy2DAsync(w + (long)r0 * n_pad + r0, (size_t)n_pad * sizeof(double), w22,
                                      (size_t)s * sizeof(double), (size_t)s * sizeof(double), (size_t)s,
                                      cudaMemcpyDeviceToDevice, st));
    return SEGP_OK;
}

}  // namespace segp
