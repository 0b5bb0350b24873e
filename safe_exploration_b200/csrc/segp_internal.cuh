// Internal declarations shared by the libsegp translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "segp.h"

namespace segp {

constexpr int TILE = 128;   // tile edge of the variance contraction; N is padded to a multiple of it
constexpr int KC = 16;      // k-depth of one pipeline stage of tri_sumsq
constexpr int NBLK = 64;    // block size of the setup factorisation kernels
constexpr int MAX_D = SEGP_MAX_NS + SEGP_MAX_NU;

void set_error(const char* fmt, ...);

#define SEGP_CUDA_CHECK(expr)                                                                       \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::segp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,    \
                              __LINE__);                                                            \
            return SEGP_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

#define SEGP_CHECK(expr)             \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != SEGP_OK) return rc__; \
    } while (0)

// true the first time it is called on the current device with this flag array: function attributes
// (cudaFuncSetAttribute) are per device, a process may drive several
inline bool first_call_on_device(bool (&seen)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (seen[dev]) return false;
    seen[dev] = true;
    return true;
}

// Shared reachability parameters, resident in device memory (uploaded once per call).
struct StepParams {
    double a[SEGP_MAX_NS * SEGP_MAX_NS];     // row-major n_s x n_s
    double b[SEGP_MAX_NS * SEGP_MAX_NU];     // row-major n_s x n_u
    double t[SEGP_MAX_NS * SEGP_MAX_NS];     // row-major n_in x n_s (only if has_t)
    double l_mu[SEGP_MAX_NS];
    double l_sigma[SEGP_MAX_NS];
    double c_safety;
    int has_t;
    int prop_mode;   // SEGP_PROP_*
};

#ifdef __CUDACC__
// Branch-free reciprocal / reciprocal square root (MUFU seed, two Newton steps: ~1 ulp) for the latency-bound chains:
// the Jacobi rotations of the ellipsoid step and the pivots of the 64 x 64 diagonal-block Cholesky.
// The IEEE division and square root of the compiler are 4x longer dependent chains with a slow-path branch each,
// and the rotation parameters sit on the critical path of a latency-bound kernel; a rotation only has to be
// orthogonal to rounding error, not correctly rounded.  Arguments here are positive and far from the subnormals.
__device__ __forceinline__ double rcp_fast(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}
#endif

// ---------------------------------------------------------------- kernel-matrix block + mean/Jacobian partials
struct KstarArgs {
    const double* xs;       // [n_s][Np][D]  training inputs scaled by 1/lengthscale_d
    const double* invls;    // [n_s][D]
    const double* var;      // [n_s]
    const double* beta;     // [n_s][Np]     zero padded
    int kern[SEGP_MAX_NS];
    int n_train, n_pad, dim, n_in, n_u, n_s_state;
    const double* z;        // direct inputs [B x D] (predict) or NULL
    const double* p;        // state centres (used when z == NULL), n_s_state per trajectory
    long p_stride;
    const double* kff;      // feed-forward controls, n_u per trajectory
    long kff_stride;
    const StepParams* sp;   // for t_z_gp (may be NULL when z != NULL)
    long n_batch, b_cap;
    int groups_per_split;   // 4-row groups handled by one blockIdx.z
    double* ks;             // [n_s][Np/4][b_cap][4]
    double* mu_part;        // [nsplit][n_s][b_cap]
    double* jac_part;       // [nsplit][n_s][D][b_cap]   (scaled coordinates, sign not yet applied)
    // composite (linear x stationary + linear) kernels only; all NULL otherwise
    const double* xraw;     // [Np][D] unscaled training inputs, zero padded
    const double* plin;     // [n_s][D] product-linear weights a_j
    const double* lin;      // [n_s][D] linear variances v_j
    const double* xtb;      // [n_s][D] X^T beta_d (the z-independent part of the Jacobian of the linear kernel)
    double* jac2_part;      // [nsplit][n_s][D][b_cap]   additive Jacobian terms, final form
    double* kss;            // [n_s][b_cap]  prior variance k_d(z_b, z_b)
};
__host__ __device__ inline bool kern_is_composite(int k) { return k == SEGP_KERN_LIN_RBF || k == SEGP_KERN_LIN_MAT52; }
int launch_kstar(const KstarArgs& a, int n_s, int nsplit, cudaStream_t st);

// ---------------------------------------------------------------- variance contraction  |W K*|^2 column sums
struct TriArgs {
    const double* wt;   // [n_s][ntri][32][128][4]   packed tiles of W = L^-1 (lower block triangle)
    const double* ks;   // [n_s][Np/4][b_cap][4]
    double* qpart;      // [n_s][nblk][b_cap]
    int nblk;           // Np / 128
    int npanels;        // ceil(B / 128)
    int group;          // panels per L2 group
    long b_cap;
    long ntri;          // nblk (nblk+1) / 2
};
int launch_tri_sumsq(const TriArgs& a, int n_s, cudaStream_t st);
int tri_sumsq_init();   // sets the dynamic shared memory attribute once per device

// ---------------------------------------------------------------- the same contraction on tcgen05 (int8 slices)
// Ozaki-style error-free splitting: every row of W and every column of K* is scaled into [-1,1] and written as
// I8_S balanced base-256 digits (int8); digit planes are multiplied pairwise on the int8 tensor
// pipe (tcgen05.mma kind::i8, exact int32 accumulation in TMEM), digit pairs (a,c) with a+c < I8_S only, one
// TMEM accumulator per diagonal a+c; the epilogue recombines the diagonals exactly in int64 (Horner, base 256),
// converts once to float64, squares and column-sums.  See DESIGN.md section 4.
constexpr int I8_S = 5;        // digit planes per operand  -> 15 int8 products, ~2^-39 relative resolution
constexpr int I8_SS = 4;       // digit planes of the diagonal-split set of W -> 10 products (+ the diagonal's extra digit)
constexpr int I8_N = 96;       // trajectories per panel (UMMA N); I8_S * I8_N = 480 <= 512 TMEM columns
constexpr int I8_KB = 64;      // bytes (= training points) per k-block row: one SWIZZLE_64B atom
constexpr int I8_A_TILE = TILE * I8_KB;    // 8192 B: 128 rows of W x 64 k, one digit plane, swizzled smem image
constexpr int I8_B_TILE = I8_N * I8_KB;    // 6144 B: 96 trajectories x 64 k
constexpr long I8_MAX_NPAD = 16384;        // int32 accumulators: 5 * 127^2 * n_pad < 2^31 and the int64 Horner bound
constexpr double I8_BASE0 = 127.0;         // first digit scale
constexpr double I8_BASE = 256.0;          // following digits (balanced: -128..127)

struct KstarI8Args {
    KstarArgs k;            // model, inputs and mean/Jacobian partial outputs (k.ks unused)
    int8_t* ki8;            // [n_s][npanel_cap][n_pad/64][I8_S][I8_B_TILE]
    long npanel_cap;
    // composite kernels: values are signed and unbounded, so every trajectory gets its own scale s_b (an analytic bound
    // on |k(z_b, x_i)| over the training inputs); the digits are those of k / s_b and the contraction multiplies its
    // column sums by s_b^2.  colfac2 [n_s][b_cap] = s_b^2 (1 for the non-composite outputs of such a model), NULL
    // for models without composite kernels; xmax [dim] = max_i |x_ij| of the raw training inputs.
    double* colfac2;
    const double* xmax;
    int panel0;             // first panel of this launch (blockIdx.x counts from it): sub-chunk pipelining
};
int launch_kstar_i8(const KstarI8Args& a, int n_s, int nsplit, cudaStream_t st);

struct TriI8Args {
    const int8_t* wi8;      // classic set [n_s][nblk (nblk+1) k-blocks][I8_S][I8_A_TILE]   (block row bi starts at bi (bi+1))
    const double* rowfac;   // [n_s][n_pad]   rowmax_i * var_d / (127^2 256^(S-1))
    const int8_t* ki8;      // as above
    double* qpart;          // [n_s][nblk][b_cap]   column sums of v^2 per block row
    float* epart;           // [n_s][nblk][b_cap]   column sums of w_i v_i^2 (error-model variance, see pack_w_i8); may be NULL
    const float* werr;      // [n_s][n_pad] variance weights w_i of the digit set in use (NULL with epart)
    const double* colfac2;  // composite kernels: per-trajectory factors s_b^2 on the column sums [n_s][b_cap], else NULL
    // diagonal-split set (digits == 4): wi8 = [n_s][nblk (nblk+1)][I8_SS][I8_A_TILE], rowfac = its row factors,
    const int8_t* wm1;      // [n_s][nblk][2][I8_A_TILE] leading digit of the diagonal entries (diagonal k-blocks only)
    int digits;             // 5 = classic set, 15 products; 4 = diagonal-split set, 10 products
    const int32_t* pflag;   // optional [npanel_cap]: only panels with a non-zero flag are computed (precision fallback)
    int nblk, npanels;      // npanels = END of the panel range of this launch, panel0 its begin (tri_i8m only; else 0)
    int panel0;
    int pgroup;             // tri_i8m / tri_i8mp: panels per L2 group (even; 0 = 24)
    int cluster;            // tri_i8m: CTAs per cluster sharing one W stage by multicast (2 or 4; 0 = 2)
    long npanel_cap, b_cap;
    int32_t* dbg;           // optional raw accumulators [I8_S][128][I8_N] of one tile (tri_i8 self-test)
    int fix_bi;             // >= 0: single-tile self-test mode (block row)
    long long* prof;        // profiling only (persistent kernel): [cluster][8] clock64 counters of the MMA thread
};
// reference kernel (one CTA per tile, classic set only): self-test and cross-check of the production kernels
int launch_tri_i8(const TriI8Args& a, int n_s, cudaStream_t st);
// single-CTA MMAs, two CTAs per cluster share one block row of W through multicast bulk copies
int launch_tri_i8m(const TriI8Args& a, int n_s, cudaStream_t st);
// persistent tri_i8m: one resident cluster per TPC walks a static list of folded (equal-length) tiles
int launch_tri_i8mp(const TriI8Args& a, int n_s, cudaStream_t st);
int tri_i8_init();
// precision guard of the 10-product digit set (tri_i8.cu)
struct GuardArgs {
    const double* qpart;    // [n_s][nblk][b_cap]
    const float* epart;     // [n_s][nblk][b_cap]
    const double* gp_var;   // [n_s] prior variances k**
    const double* kss;      // composite kernels: per-trajectory prior variances [n_s][b_cap] (else NULL: gp_var)
    int nblk, n_s, panel0;
    long b_cap, n_batch;    // n_batch = END of the trajectory range
    double gs;              // (2 kappa / rtol)^2
    int32_t* pflag;         // [npanel_cap] out
    unsigned int* counter;  // optional: number of flagged panels (accumulates)
};
int launch_i8_guard(const GuardArgs& a, cudaStream_t st);
int launch_scale_f32(float* x, long n, float f, cudaStream_t st);
// W (n_pad x n_pad fp64, lower) -> both digit-plane sets, row factors and error-model weights of output dimension d
struct PackI8Out {
    int8_t* wi8;        // [nblk (nblk+1)][I8_S][I8_A_TILE]
    double* rowfac;     // [n_pad]
    int8_t* wi8s;       // [nblk (nblk+1)][I8_SS][I8_A_TILE]
    int8_t* wm1;        // [nblk][2][I8_A_TILE]
    double* rowfac_s;   // [n_pad]
    float* werr5;       // [n_pad]
    float* werr4;       // [n_pad]
};
int pack_w_i8(const double* w, const PackI8Out& o, double var, int n_pad, int n_train, cudaStream_t st);
int i8_peak(int umma_n, int iters, int pattern, double* tops);

// ---------------------------------------------------------------- posterior finalise / ellipsoid step
struct StepArgs {
    // GP outputs as partials (fused path) ...
    const double* mu_part;
    const double* jac_part;
    const double* qpart;
    const double* gp_var;   // [n_s] signal variances
    const double* invls;    // [n_s][D]
    const double* jac2_part;   // composite kernels: additive Jacobian partials (else NULL)
    const double* kss;         // composite kernels: per-trajectory prior variance [n_s][b_cap] (else NULL: gp_var)
    const float* epart;        // int8 contraction: error-model variance partials [n_s][nblk][b_cap] (else NULL)
    double guard_gs;           // (2 kappa / rtol)^2: SEGP_STATUS_LOW_PRECISION when guard_gs * e2 > sigma^4
    int nsplit, nblk;
    long b_cap;
    // ... or given directly (foreign state-space model): [B x n_s], [B x n_s], [B x n_s x D]
    const double* mu_d;
    const double* var_d;
    const double* jac_d;
    // state
    const double* p;
    long p_stride;
    const double* q;        // NULL -> point branch
    long q_stride;
    const double* kff;
    long kff_stride;
    const double* kfb;
    long kfb_stride;
    const StepParams* sp;
    double* p_out;
    long p_out_stride;
    double* q_out;
    long q_out_stride;
    double* var_out;        // may be NULL
    long var_out_stride;
    int32_t* status;        // may be NULL
    long n_batch;           // END of the trajectory range of this launch, b0 its begin (all indices are absolute)
    long b0;
    int n_s, n_in, n_u;
};
int launch_ellipsoid_step(const StepArgs& a, cudaStream_t st);

struct FinalizeArgs {
    const double* mu_part;
    const double* jac_part;
    const double* qpart;
    const double* gp_var;
    const double* invls;
    const double* jac2_part;   // composite kernels only (else NULL)
    const double* kss;
    const float* epart;        // as StepArgs
    double guard_gs;
    int32_t* status;           // [B] or NULL: SEGP_STATUS_BAD_VARIANCE / SEGP_STATUS_LOW_PRECISION per input
    int nsplit, nblk, n_s, dim;
    long b_cap, n_batch;
    double* mu;    // [B x n_s]
    double* var;   // [B x n_s]
    double* jac;   // [B x n_s x D] or NULL
};
int launch_finalize_predict(const FinalizeArgs& a, cudaStream_t st);

int launch_remainder(long n_batch, int n_s, int n_u, const double* q, const double* kfb, long kfb_stride,
                     const StepParams* sp, double* u_mu, double* u_sigma, cudaStream_t st);
int launch_sum_two(long n_batch, int n, const double* p1, const double* q1, const double* p2, const double* q2,
                   double* p, double* q, cudaStream_t st);
int launch_from_rectangle(long n_batch, int n, const double* ub, double* q, int32_t* status, cudaStream_t st);
int launch_safety_distance(long n_items, int n_s, int m, const double* p, const double* q, const double* hmat,
                           const double* hvec, double c, double* dist, cudaStream_t st);

// ---------------------------------------------------------------- candidate scoring (constraints, cost, arg-best)
struct ScoreParams {
    double u_min[SEGP_MAX_NU], u_max[SEGP_MAX_NU];
    double h_mat_obs[SEGP_MAX_CONSTR * SEGP_MAX_NS], h_obs[SEGP_MAX_CONSTR];
    double h_mat_safe[SEGP_MAX_CONSTR * SEGP_MAX_NS], h_safe[SEGP_MAX_CONSTR];
    double wx[SEGP_MAX_NS * SEGP_MAX_NS], wu[SEGP_MAX_NU * SEGP_MAX_NU], x_ref[SEGP_MAX_NS];
    double c_safety, eps_constraints, eps_noise;
    double q0[SEGP_MAX_NS * SEGP_MAX_NS], kfb0[SEGP_MAX_NU * SEGP_MAX_NS];   // init_uncertainty: shared Q_0, K_fb_0
    int has_ctrl, m_obs, m_safe, cost_type, layout, has_q0;
};
struct ScoreArgs {
    const double* p_all;      // [B][H][n_s]
    const double* q_all;      // [B][H][n_s][n_s]
    const double* var_all;    // [B][H][n_s]  (exploration cost only)
    const double* kff;        // [B][H][n_u]
    const double* kfb;        // [(B)][H-1][n_u][n_s]
    long kfb_stride;
    const int32_t* status;    // [B] or NULL
    const ScoreParams* sp;
    double* cost;             // [B]
    int32_t* feasible;        // [B]
    double* violation;        // [B]
    double* g;                // [B][n_g] or NULL
    long n_batch;
    int horizon, n_s, n_u, n_g;
};
struct BestCandidate {
    long index;
    int feasible;
    int pad_;
    double cost, violation;
};
int launch_score(const ScoreArgs& a, cudaStream_t st);
int launch_argbest(long n, const double* cost, const int32_t* feasible, const double* violation, BestCandidate* out,
                   cudaStream_t st);

// ---------------------------------------------------------------- factorisation GEMMs
// flags shared by gemm64 (float64 DMMA, setup.cu) and gemm_i8d (tcgen05 digit planes, fact_i8.cu)
constexpr int GEMM_A_LOWER = 1;   // A[i,k] == 0 for k > i   -> the k loop stops at the row tile's end
constexpr int GEMM_B_LOWER = 2;   // B[k,n] == 0 for k < n   -> the k loop starts at the column tile's start
constexpr int GEMM_C_LOWER = 4;   // only tiles on or below the diagonal are computed (syrk)

// C[z] (M_z x N) = alpha A[z] B[z]^T + beta C[z] from digit planes (fd_split): planes [row tile of 128][k-block of
// 64][8][8 KB] + one scale per operand row; M_z = min(m, m_total - z * zrows) as GemmArgs.
struct GemmI8Args {
    const int8_t* ap;
    const double* as;
    int a_kb;            // k-blocks stored per row tile of A (the stride between row tiles)
    const int8_t* bp;
    const double* bs;
    int b_kb;
    double* c;
    long ldc;
    int m, n, k;         // multiples of 128, 128, 64
    double alpha, beta;
    int flags;
    long z_ap, z_as, z_bp, z_bs, z_c;   // batch strides: bytes (planes), elements (scales, C)
    int m_total, zrows;
};
// scratch of one factorisation stream: two operand plane sets + their row scales
struct FdScratch {
    int8_t* ap;
    int8_t* bp;
    double* as;
    double* bs;
};
size_t fd_scratch_plane_bytes(int n_pad);          // bytes per operand plane set for a model of n_pad points
size_t fd_plane_bytes(int rows, int cols, int batch);
// operand O[r][k] = transposed ? X[k ld + r] : X[r ld + k]; tri 1: zero for k > r, 2: zero for k < r; clip bit 0 / 1:
// the ragged last batch entry bounds the rows / the k range by min(nominal, lim_total - z zrows)
int fd_split(const double* x, long ld, long zstride, int rows, int cols, int transposed, int tri, int clip, int lim_total,
             int zrows, int batch, double* scale, int8_t* planes, cudaStream_t st);
int launch_gemm_i8d(const GemmI8Args& g, int batch, cudaStream_t st);
int gemm_i8d_selftest(int m, int n, int k, const double* h_a, const double* h_b, double* h_c, double alpha, double beta,
                       int trans_b, int flags);

// ---------------------------------------------------------------- setup (factorisation), all float64 on device
struct SetupDims {
    int n_train, n_pad, dim;
};
// K[d] (n_pad x n_pad row-major) = k_d(X,X) + noise I; identity on the padded diagonal.
// composite kernels: xraw [n_pad][dim] unscaled inputs, plin_d / lin_d [dim] device vectors (else NULL)
int launch_kmat(double* k, const double* xs_d, int kern, double var, double noise, SetupDims s, const double* xraw,
                const double* plin_d, const double* lin_d, cudaStream_t st);
// rows [row0, row0 + nrows) of the same matrix into a nrows x n_pad buffer
int launch_kmat_rows(double* k, const double* xs_d, int kern, double var, double noise, SetupDims s, const double* xraw,
                     const double* plin_d, const double* lin_d, int row0, int nrows, cudaStream_t st);
// incremental update of the dense W = L^-1 when rows [r0, r0 + s) of K + noise I are new (setup.cu)
int append_rows(double* w, int n_pad, int r0, int s, const double* krows, double* l21, double* tbuf, double* sbuf,
                double* w22, double* tmp, double* diag_inv, int* d_fail, cudaStream_t st, long* launches);
int logdet_from_winv(const double* w, int n_train, int n_pad, double* d_out, cudaStream_t st);
// xtb[j] = sum_i beta[i] xraw[i][j]  (fixed-order block reduction)
int launch_xtb(const double* xraw, const double* beta, double* xtb, int n_pad, int dim, cudaStream_t st);
// in-place blocked Cholesky (lower) of a (n_pad x n_pad); diag_inv gets the inverses of the 64x64 diagonal blocks;
// *d_fail (device int) is set to 1+pivot index on a non-positive pivot.
// fd != NULL: the K = 256 trailing updates (two-level blocking) run on the tcgen05 digit-plane GEMM (fact_i8.cu).
int potrf_lower(double* a, int n_pad, double* diag_inv, int* d_fail, cudaStream_t st, long* launches,
                const FdScratch* fd = nullptr);
// w (zero-initialised n_pad x n_pad) = inverse of the lower factor l; tmp is n_pad x n_pad scratch.
// fd != NULL: the levels with blocks of >= 256 rows run on the tcgen05 digit-plane GEMM.
int trtri_lower(const double* l, double* w, int n_pad, const double* diag_inv, double* tmp, cudaStream_t st,
                long* launches, const FdScratch* fd = nullptr);
// beta = W^T (W y)
int solve_beta(const double* w, const double* y, double* u_tmp, double* beta, int n_pad, cudaStream_t st);
// pack W into [ntri][32][128][4] tiles
int pack_w(const double* w, double* wt, int n_pad, cudaStream_t st);
// logdet = 2 sum log L_ii over the first n_train rows (single block reduction into *d_out)
int logdet_from_chol(const double* l, int n_train, int n_pad, double* d_out, cudaStream_t st);

}  // namespace segp
