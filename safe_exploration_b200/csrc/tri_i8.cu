// The variance contraction |L_d^-1 K*_d[:,b]|^2 on the 5th-generation tensor cores (tcgen05, int8 -> int32 in TMEM).
//
// Why int8 and not bf16/tf32: sigma^2 = k** - |L^-1 k*|^2 cancels ~4 digits on the benchmark models, so the
// contraction needs ~2^-36 relative accuracy -- out of reach of any fp32-accumulating pipe.
// Integer accumulation is exact, so an error-free splitting (Ozaki scheme) recovers float64-grade results from
// int8 products: every row of W = L^-1 and every column of K* is scaled into [-1, 1] and written as I8_S balanced
// base-256 digits; the digit planes are multiplied pairwise (a + c < I8_S: 15 products) by tcgen05.mma kind::i8,
// one TMEM accumulator per diagonal a + c, and the epilogue recombines the diagonals exactly in int64 before the
// single conversion to float64.  Replaces, like tri_sumsq, `sum2(mtimes(k*, K^-1) * k*)` of
// ssm_gpy/gp_models_utils_casadi.py:190-193 / GPy predict_noiseless.
//
//   pack_w_i8      W_d (fp64)            -> digit planes as SWIZZLE_64B shared-memory images + per-row factors
//   kstar_i8       K*_d[i,b] (fp64 exp)  -> digit planes (same image format), fused with the mean / Jacobian sums
//   tri_i8         per (d, block row, 96-trajectory panel): bulk-copy ring -> tcgen05.mma -> TMEM -> column sums
//   i8_peak        register/shared-only issue loop that measures the int8 tensor-pipe rate (roofline denominator)
#include <math.h>

#include <algorithm>

#include "segp_internal.cuh"
#include "tc_i8.cuh"

namespace segp {

// =========================================================================================== digits
// r in [-1, 1]  ->  V = rn(r 127 2^32), an integer of at most 39 bits + sign, written in BALANCED base 256:
//   V = d0 2^32 + d1 2^24 + d2 2^16 + d3 2^8 + d4,   d0 in [-127, 127], d1..d4 in [-128, 127]   (all int8)
// i.e. r ~= d0/127 + d1/(127 256) + ... with a remainder below 0.5 / (127 256^4) = 9.2e-13.  A power-of-two base makes
// the digit split shifts and masks (no divisions), and the epilogue's Horner steps exact multiplications.
__device__ __forceinline__ void split_digits(double r, int (&dg)[I8_S]) {
    static_assert(I8_S == 5, "digit layout below is for 5 planes");
    long long v = __double2ll_rn(r * (I8_BASE0 * 4294967296.0));
#pragma unroll
    for (int a = I8_S - 1; a >= 1; --a) {
        const int d = (int)((v + 128) & 255) - 128;
        dg[a] = d;
        v = (v - d) >> 8;
    }
    dg[0] = (int)v;
}

// The same digits for u in [0, 1] (kernel values) in four integer instructions: V = rn(u 127 2^32) lands in the low
// mantissa bits of u 127 2^32 + 1.5 2^52; adding 0x80 to each of the four low bytes turns them into excess-128 digits
// (the carries ripple into the next byte, which is exactly the balanced recoding), and the XOR with 0x80 makes them
// two's complement.  Returns d1..d4 packed as int8 bytes (d1 in byte 3 ... d4 in byte 0) and d0 (0..127) in d0.
__device__ __forceinline__ uint32_t split_digits_unit(double u, int& d0) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    const double t = fma(u, I8_BASE0 * 4294967296.0, MAGIC);
    const uint32_t lo = (uint32_t)__double2loint(t);
    const uint32_t hi = (uint32_t)__double2hiint(t) & 0x7FFFFu;   // V >> 32 (the 2^51 of MAGIC sits at bit 19)
    const uint32_t lo2 = lo + 0x80808080u;
    d0 = (int)(hi + (lo2 < lo ? 1u : 0u));
    return lo2 ^ 0x80808080u;
}

// The same for r in [-1, 1] (composite kernel values scaled by the per-trajectory bound): the mantissa of
// r 127 2^32 + 1.5 2^52 holds 2^51 + V in two's complement, so bits 32..50 are V >> 32 modulo 2^19 -- sign-extend them.
__device__ __forceinline__ uint32_t split_digits_signed(double r, int& d0) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    const double t = fma(r, I8_BASE0 * 4294967296.0, MAGIC);
    const uint32_t lo = (uint32_t)__double2loint(t);
    const int hi = ((int)((uint32_t)__double2hiint(t) << 13)) >> 13;   // sign-extended 19 bits
    const uint32_t lo2 = lo + 0x80808080u;
    d0 = hi + (lo2 < lo ? 1 : 0);
    return lo2 ^ 0x80808080u;
}

// =========================================================================================== pack_w_i8
// Two digit-plane sets of W = L^-1, both stored as SWIZZLE_64B shared-memory images:
//   classic  rows scaled by their max-abs, 5 digits                                   (15 products, tri digits = 5)
//   split    rows scaled by the max-abs of their OFF-diagonal entries, 4 digits; the diagonal entry 1/L_ii, which
//            dominates every row of L^-1 (median 15 x the largest off-diagonal entry at C4) and would waste the
//            leading plane, keeps an extra leading digit d_-1 (weight 256) in a plane of its own that is non-zero on
//            the diagonal only and is multiplied in the two diagonal k-blocks of a block row only (10 products)
// plus, per row, the factor that turns the recombined integer into v_i and the variance weight of the statistical
// error model (see i8_err_weight): the epilogue column-sums w_i v_i^2 next to v_i^2, which is the a-posteriori error
// estimate of |v|^2 behind the precision guard.
struct RowStat {
    double maxall, maxoff, sqall, sqoff, diag;
};
__global__ void w_rowstat_kernel(const double* __restrict__ w, RowStat* __restrict__ rs, int n_pad) {
    __shared__ double s0[256], s1[256], s2[256], s3[256];
    const int i = blockIdx.x;
    double m = 0.0, mo = 0.0, q = 0.0, qo = 0.0;
    for (int j = threadIdx.x; j <= i; j += 256) {
        const double v = w[(long)i * n_pad + j];
        m = fmax(m, fabs(v));
        q = fma(v, v, q);
        if (j < i) {
            mo = fmax(mo, fabs(v));
            qo = fma(v, v, qo);
        }
    }
    s0[threadIdx.x] = m;
    s1[threadIdx.x] = mo;
    s2[threadIdx.x] = q;
    s3[threadIdx.x] = qo;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] = fmax(s0[threadIdx.x], s0[threadIdx.x + o]);
            s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + o]);
            s2[threadIdx.x] += s2[threadIdx.x + o];
            s3[threadIdx.x] += s3[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) rs[i] = RowStat{s0[0], s1[0], s2[0], s3[0], w[(long)i * n_pad + i]};
}

// row scale of the split planes: off-diagonal max, but never so small that the diagonal's leading digit overflows
__device__ __forceinline__ double split_scale(const RowStat& r) { return fmax(r.maxoff, fabs(r.diag) * (1.0 / I8_BASE)); }

// Variance of the error of v_i = sum_j W_ij k*_j computed with S digits per operand and the digit pairs a + c < S,
// digits taken as independent and uniform:  per product term, in units u = 1 / (127 256^(S-1)) of the last digit,
//   truncation of k* against w~      w~_ij^2 / 12
//   truncation of w~ against k~      k~_j^2 / 12          (k~ = k*/s_f^2 in [0,1]: mean square bounded by 1/2)
//   dropped pairs a + c = S (S - 1 of them)   (256 / 127)^2 / 144 = 1 / 35.4 each
// times (row scale  s_f^2  u)^2.  i8_scheme.py (scripts/experiments) checks it against the emulated scheme.
__device__ __forceinline__ double i8_err_weight(double scale, double var, double sumsq_scaled, int n_terms, int digits) {
    double u = 1.0 / I8_BASE0;
    for (int a = 1; a < digits; ++a) u /= I8_BASE;
    const double f = scale * var * u;
    // a dropped pair d_a d_c / (127 256) in units of u: both digits uniform over 256 values -> variance (256/127)^2 / 144
    const double pair = (I8_BASE / I8_BASE0) * (I8_BASE / I8_BASE0) / 144.0;
    return f * f * (sumsq_scaled * (1.0 / 12.0) + (double)n_terms * (0.5 / 12.0 + (double)(digits - 1) * pair));
}

__global__ void __launch_bounds__(TILE) pack_w_i8_kernel(const double* __restrict__ w, const RowStat* __restrict__ rs,
                                                         const PackI8Out o, double var, int n_pad, int n_train) {
    const int kb = blockIdx.x, bi = blockIdx.y;
    if (kb >= 2 * (bi + 1)) return;
    const int r = threadIdx.x;
    const long row = (long)bi * TILE + r;
    const RowStat st = rs[row];
    const double rm = st.maxall;
    const double inv = rm > 0.0 ? 1.0 / rm : 0.0;
    const double sc = split_scale(st);
    const double inv_s = sc > 0.0 ? 1.0 / sc : 0.0;
    if (kb == 0) {
        double f = var / (I8_BASE0 * I8_BASE0);
#pragma unroll
        for (int a = 1; a < I8_S; ++a) f /= I8_BASE;
        o.rowfac[row] = rm * f;
        o.rowfac_s[row] = sc * f * I8_BASE;   // accumulator slots 0..4 hold the diagonals g = -1..3: one power less
        const bool real = row < n_train;
        o.werr5[row] = real ? (float)i8_err_weight(rm, var, st.sqall * inv * inv, (int)row + 1, 5) : 0.f;
        o.werr4[row] = real ? (float)i8_err_weight(sc, var, st.sqoff * inv_s * inv_s, (int)row + 1, 4) : 0.f;
    }
    const double* src = w + row * n_pad + (long)kb * I8_KB;
    int8_t* dst = o.wi8 + ((long)bi * (bi + 1) + kb) * (I8_S * I8_A_TILE);
    int8_t* dst_s = o.wi8s + ((long)bi * (bi + 1) + kb) * (I8_SS * I8_A_TILE);
    const bool diag_kb = kb >= 2 * bi;   // the two k-blocks that hold the diagonal block
    int8_t* dst_m1 = o.wm1 + ((long)bi * 2 + (kb - 2 * bi)) * I8_A_TILE;
#pragma unroll 1
    for (int c = 0; c < I8_KB / 16; ++c) {
        uint32_t pk[I8_S][4], ps[I8_SS + 1][4];
#pragma unroll
        for (int a = 0; a < I8_S; ++a) pk[a][0] = pk[a][1] = pk[a][2] = pk[a][3] = 0u;
#pragma unroll
        for (int a = 0; a <= I8_SS; ++a) ps[a][0] = ps[a][1] = ps[a][2] = ps[a][3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const long col = (long)kb * I8_KB + c * 16 + j;
            const double v = (col <= row) ? src[c * 16 + j] : 0.0;   // strictly lower + diagonal only
            int dg[I8_S];
            split_digits(v * inv, dg);
#pragma unroll
            for (int a = 0; a < I8_S; ++a) pk[a][j >> 2] |= (uint32_t)(dg[a] & 0xff) << ((j & 3) * 8);
            // split set: off-diagonal entries v / sc in [-1, 1] -> digits 0..3; the diagonal v / (256 sc) in [-1, 1]
            // -> digits -1..3 (one level up); slot 0 of ps is the d_-1 plane
            int ds[I8_S];
            if (col == row) {
                split_digits(v * inv_s * (1.0 / I8_BASE), ds);
#pragma unroll
                for (int a = 0; a <= I8_SS; ++a) ps[a][j >> 2] |= (uint32_t)(ds[a] & 0xff) << ((j & 3) * 8);
            } else {
                split_digits(v * inv_s, ds);
#pragma unroll
                for (int a = 0; a < I8_SS; ++a) ps[a + 1][j >> 2] |= (uint32_t)(ds[a] & 0xff) << ((j & 3) * 8);
            }
        }
        const int off = sw64_offset(r, c * 16);
#pragma unroll
        for (int a = 0; a < I8_S; ++a)
            *reinterpret_cast<uint4*>(dst + (long)a * I8_A_TILE + off) = make_uint4(pk[a][0], pk[a][1], pk[a][2], pk[a][3]);
#pragma unroll
        for (int a = 0; a < I8_SS; ++a)
            *reinterpret_cast<uint4*>(dst_s + (long)a * I8_A_TILE + off) =
                make_uint4(ps[a + 1][0], ps[a + 1][1], ps[a + 1][2], ps[a + 1][3]);
        if (diag_kb) *reinterpret_cast<uint4*>(dst_m1 + off) = make_uint4(ps[0][0], ps[0][1], ps[0][2], ps[0][3]);
    }
}

int pack_w_i8(const double* w, const PackI8Out& o, double var, int n_pad, int n_train, cudaStream_t st) {
    const int nblk = n_pad / TILE;
    RowStat* rs = nullptr;
    SEGP_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&rs), (size_t)n_pad * sizeof(RowStat), st));
    w_rowstat_kernel<<<n_pad, 256, 0, st>>>(w, rs, n_pad);
    dim3 grid((unsigned)(2 * nblk), (unsigned)nblk);
    pack_w_i8_kernel<<<grid, TILE, 0, st>>>(w, rs, o, var, n_pad, n_train);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(rs, st);
    SEGP_CUDA_CHECK(e);
    return SEGP_OK;
}

// =========================================================================================== kstar_i8
// One block = one 96-trajectory panel (thread = trajectory = one row of the B tile images), blockIdx.y = output
// dimension, blockIdx.z = split of the training points.  Same arithmetic as kstar_mean_jac (predict.cu); the kernel
// value leaves as I8_S digit bytes instead of one double.
// exp(x) for x <= 0 without branches: x = (64 m + j) ln2/64 + r, |r| <= ln2/128, exp(x) = 2^m 2^(j/64) e^r with a
// 64-entry table in shared memory and a degree-5 polynomial (remainder 3.5e-17 relative).  ~1.5 ulp; 10 FP64
// instructions instead of ~17 plus a slow-path branch for the library exp, and it schedules across unrolled points.
__device__ __forceinline__ double exp_neg_tab(double x, const double* __restrict__ s_tab) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer
    const double t = fma(x, 92.33248261689366, MAGIC);
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -0.01083042469326756, x);    // ln2/64, high part (32 significant bits: nd * hi is exact)
    r = fma(nd, -2.9815858269852933e-12, r);        // low part
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = s_tab[n & 63] * p;
    const double scaled = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
    return x < -700.0 ? 0.0 : scaled;
}

// 128 staged training points against one trajectory: kernel values -> digit bytes, mean / Jacobian sums.
template <int DM, int KERN, int UNR, int ROWS>
__device__ __forceinline__ void kstar_i8_rows(const double* __restrict__ s_x, const double* __restrict__ s_beta,
                                              const double* __restrict__ s_tab,
                                              const double (&zs)[DM], int dim, double var, int row0, int n_train,
                                              bool active, int8_t* __restrict__ kb_base, int rowp, long plane_stride,
                                              double& mu, double (&jac)[DM]) {
    const double sqrt5 = 2.23606797749978969641;
    // UNR independent points per iteration (8 or 16): ILP against code size / registers
    static_assert(UNR == 8 || UNR == 16, "unroll");
#pragma unroll 1
    for (int r = 0; r < ROWS; r += UNR) {
        uint32_t pk[I8_S][UNR / 4];
        uint32_t wlo[UNR];   // digits 1..4 of every point, packed; digit 0 in whi
        int whi[UNR];
#pragma unroll
        for (int qd = 0; qd < UNR; ++qd) {
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
            double unit, g;   // k / var  and  (dk/dr2-type factor) / var
            if (KERN == SEGP_KERN_RBF) {
                unit = exp_neg_tab(-0.5 * r2, s_tab);
                g = unit;
            } else {
                const double rr = sqrt(r2);
                const double e = exp_neg_tab(-sqrt5 * rr, s_tab);
                unit = (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * e;
                g = (5.0 / 3.0) * (1.0 + sqrt5 * rr) * e;
            }
            if (row0 + r + qd >= n_train || !active) unit = 0.0;   // padded rows / columns: zero digits
            const double bt = s_beta[r + qd];   // beta_i sigma_f^2 (multiplied once when staged)
            mu = fma(bt, unit, mu);
            const double w = bt * g;
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (j < dim) jac[j] = fma(w, diff[j], jac[j]);
            wlo[qd] = split_digits_unit(unit, whi[qd]);
        }
        // 4 x 4 byte transposes: plane s gets digit s of four consecutive points per word (three PRMTs each)
#pragma unroll
        for (int q4 = 0; q4 < UNR / 4; ++q4) {
            const uint32_t w0 = wlo[4 * q4], w1 = wlo[4 * q4 + 1], w2 = wlo[4 * q4 + 2], w3 = wlo[4 * q4 + 3];
            pk[0][q4] = __byte_perm(__byte_perm((uint32_t)whi[4 * q4], (uint32_t)whi[4 * q4 + 1], 0x0040),
                                    __byte_perm((uint32_t)whi[4 * q4 + 2], (uint32_t)whi[4 * q4 + 3], 0x0040), 0x5410);
            pk[1][q4] = __byte_perm(__byte_perm(w0, w1, 0x0073), __byte_perm(w2, w3, 0x0073), 0x5410);
            pk[2][q4] = __byte_perm(__byte_perm(w0, w1, 0x0062), __byte_perm(w2, w3, 0x0062), 0x5410);
            pk[3][q4] = __byte_perm(__byte_perm(w0, w1, 0x0051), __byte_perm(w2, w3, 0x0051), 0x5410);
            pk[4][q4] = __byte_perm(__byte_perm(w0, w1, 0x0040), __byte_perm(w2, w3, 0x0040), 0x5410);
        }
        const int kglob = row0 + r;
        int8_t* dst = kb_base + (long)(kglob >> 6) * (I8_S * I8_B_TILE) + sw64_offset(rowp, kglob & 63);
#pragma unroll
        for (int s = 0; s < I8_S; ++s) {
            if (UNR == 16)
                *reinterpret_cast<uint4*>(dst + (long)s * plane_stride) =
                    make_uint4(pk[s][0], pk[s][1], pk[s][UNR / 4 - 2], pk[s][UNR / 4 - 1]);
            else
                *reinterpret_cast<uint2*>(dst + (long)s * plane_stride) = make_uint2(pk[s][0], pk[s][1]);
        }
    }
}

// GP input of trajectory b: given directly (predict), or [T p; k_ff] from the rollout state (raw, not yet scaled)
template <int DM>
__device__ __forceinline__ void kstar_i8_load_inputs(const KstarArgs& a, long b, bool active, int dim, double (&zs)[DM]) {
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
}

// The composite kernels (linear x stationary + linear, predict.cu:kstar_composite) on the digit-plane path: the same
// float64 kernel values, mean / Jacobian sums and prior variance, with the value leaving as digit bytes of k / s_b.
// s_b = s_f^2 sum_j |a_j z_j| xmax_j + sum_j |v_j z_j| xmax_j bounds |k(z_b, x_i)| for every training input.
// Same structure as the stationary kernels' kstar_i8_rows: 128 training points staged through shared memory (scaled
// and raw coordinates, beta), the table exponential, 8 independent points per iteration, three PRMTs per plane word.
// training points staged at a time: scaled + raw coordinates must fit the 48 KB of static shared memory
__host__ __device__ constexpr int comp_stage_rows(int dm) { return dm > 12 ? 64 : TILE; }

template <int DM, int KERN, int ROWS>
__device__ __forceinline__ void kstar_i8_rows_composite(const double* __restrict__ s_x, const double* __restrict__ s_xr,
                                                        const double* __restrict__ s_beta, const double* __restrict__ s_tab,
                                                        const double (&zs)[DM], const double (&za)[DM],
                                                        const double (&zv)[DM], int dim, double var, double inv_s, int row0,
                                                        int n_train, bool active, int8_t* __restrict__ kb_base, int rowp,
                                                        double& mu, double (&jac)[DM], double (&jac2)[DM]) {
    constexpr int UNR = 8;
    const double sqrt5 = 2.23606797749978969641;
#pragma unroll 1
    for (int r = 0; r < ROWS; r += UNR) {
        uint32_t pk[I8_S][UNR / 4];
        uint32_t wlo[UNR];
        int whi[UNR];
#pragma unroll
        for (int qd = 0; qd < UNR; ++qd) {
            const double* xsr = s_x + (r + qd) * dim;
            const double* xr = s_xr + (r + qd) * dim;
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
            if (KERN == SEGP_KERN_LIN_RBF) {
                stat = var * exp_neg_tab(-0.5 * r2, s_tab);
                gg = stat;
            } else {
                const double rr = sqrt(r2);
                const double e = var * exp_neg_tab(-sqrt5 * rr, s_tab);
                stat = (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * e;
                gg = (5.0 / 3.0) * (1.0 + sqrt5 * rr) * e;
            }
            double kval = fma(lp, stat, ll);
            if (row0 + r + qd >= n_train || !active) kval = 0.0;
            const double bt = s_beta[r + qd];   // zero on padded rows
            mu = fma(bt, kval, mu);
            const double w = bt * lp * gg, w2 = bt * stat;
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (j < dim) {
                    jac[j] = fma(w, diff[j], jac[j]);
                    jac2[j] = fma(w2, xr[j], jac2[j]);
                }
            // |kval| <= s_b up to rounding: clamp so the leading digit stays within +-127
            wlo[qd] = split_digits_signed(fmax(-1.0, fmin(1.0, kval * inv_s)), whi[qd]);
        }
#pragma unroll
        for (int q4 = 0; q4 < UNR / 4; ++q4) {
            const uint32_t w0 = wlo[4 * q4], w1 = wlo[4 * q4 + 1], w2 = wlo[4 * q4 + 2], w3 = wlo[4 * q4 + 3];
            pk[0][q4] = __byte_perm(__byte_perm((uint32_t)whi[4 * q4], (uint32_t)whi[4 * q4 + 1], 0x0040),
                                    __byte_perm((uint32_t)whi[4 * q4 + 2], (uint32_t)whi[4 * q4 + 3], 0x0040), 0x5410);
            pk[1][q4] = __byte_perm(__byte_perm(w0, w1, 0x0073), __byte_perm(w2, w3, 0x0073), 0x5410);
            pk[2][q4] = __byte_perm(__byte_perm(w0, w1, 0x0062), __byte_perm(w2, w3, 0x0062), 0x5410);
            pk[3][q4] = __byte_perm(__byte_perm(w0, w1, 0x0051), __byte_perm(w2, w3, 0x0051), 0x5410);
            pk[4][q4] = __byte_perm(__byte_perm(w0, w1, 0x0040), __byte_perm(w2, w3, 0x0040), 0x5410);
        }
        const int kglob = row0 + r;
        int8_t* dst = kb_base + (long)(kglob >> 6) * (I8_S * I8_B_TILE) + sw64_offset(rowp, kglob & 63);
#pragma unroll
        for (int pl = 0; pl < I8_S; ++pl)
            *reinterpret_cast<uint2*>(dst + (long)pl * I8_B_TILE) = make_uint2(pk[pl][0], pk[pl][1]);
    }
}

template <int DM>
__device__ __forceinline__ void kstar_i8_composite(const KstarI8Args& aa, const double (&z)[DM], int dim, int d, int split,
                                                   int n_s, long b, bool active, int8_t* __restrict__ kb_base, int rowp,
                                                   double* __restrict__ s_x, double* __restrict__ s_xr,
                                                   double* __restrict__ s_beta, const double* __restrict__ s_tab) {
    const KstarArgs& a = aa.k;
    const int kern = a.kern[d];
    const double var = a.var[d];
    const double* __restrict__ pl = a.plin + d * dim;
    const double* __restrict__ lv = a.lin + d * dim;
    double zs[DM], za[DM], zv[DM], jac[DM], jac2[DM];
    double kss = 0.0, kss_lin = 0.0, bound = 0.0;
#pragma unroll
    for (int j = 0; j < DM; ++j) {
        zs[j] = za[j] = zv[j] = jac[j] = jac2[j] = 0.0;
        if (j < dim) {
            zs[j] = z[j] * a.invls[d * dim + j];
            za[j] = z[j] * pl[j];
            zv[j] = z[j] * lv[j];
            kss = fma(za[j], z[j], kss);
            kss_lin = fma(zv[j], z[j], kss_lin);
            bound += (fabs(za[j]) * var + fabs(zv[j])) * aa.xmax[j];
        }
    }
    bound *= 1.0 + 1e-12;
    if (!(bound > 0.0) || !active) bound = 1.0;
    const double inv_s = 1.0 / bound;
    double mu = 0.0;
    constexpr int STAGE = comp_stage_rows(DM);
    const int row_begin = split * a.groups_per_split * 4;
    const int row_end = min(row_begin + a.groups_per_split * 4, a.n_pad);
    for (int row0 = row_begin; row0 < row_end; row0 += STAGE) {
        __syncthreads();
        const double* src = a.xs + ((long)d * a.n_pad + row0) * dim;
        const double* srcr = a.xraw + (long)row0 * dim;
        for (int idx = threadIdx.x; idx < STAGE * dim; idx += I8_N) {
            s_x[idx] = src[idx];
            s_xr[idx] = srcr[idx];
        }
        for (int idx = threadIdx.x; idx < STAGE; idx += I8_N) s_beta[idx] = a.beta[(long)d * a.n_pad + row0 + idx];
        __syncthreads();
        if (kern == SEGP_KERN_LIN_RBF)
            kstar_i8_rows_composite<DM, SEGP_KERN_LIN_RBF, STAGE>(s_x, s_xr, s_beta, s_tab, zs, za, zv, dim, var, inv_s, row0,
                                                                  a.n_train, active, kb_base, rowp, mu, jac, jac2);
        else
            kstar_i8_rows_composite<DM, SEGP_KERN_LIN_MAT52, STAGE>(s_x, s_xr, s_beta, s_tab, zs, za, zv, dim, var, inv_s, row0,
                                                                    a.n_train, active, kb_base, rowp, mu, jac, jac2);
    }
    if (!active) return;
    a.mu_part[((long)split * n_s + d) * a.b_cap + b] = mu;
#pragma unroll
    for (int j = 0; j < DM; ++j)
        if (j < dim) {
            a.jac_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = jac[j];
            a.jac2_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] =
                pl[j] * jac2[j] + (split == 0 ? lv[j] * a.xtb[d * dim + j] : 0.0);
        }
    if (split == 0) {
        a.kss[(long)d * a.b_cap + b] = fma(var, kss, kss_lin);
        aa.colfac2[(long)d * a.b_cap + b] = bound * bound;
    }
}

// One work item = (96-trajectory panel, output dimension, split of the training points).  STAGE training points are
// staged through shared memory at a time.
template <int D_T, int STAGE>
__device__ __forceinline__ void kstar_i8_item(const KstarI8Args& aa, int panel, int d, int split, int n_s,
                                              double* __restrict__ s_x, double* __restrict__ s_beta,
                                              const double* __restrict__ s_tab) {
    const KstarArgs& a = aa.k;
    constexpr int DM = D_T > 0 ? D_T : MAX_D;
    const int dim = D_T > 0 ? D_T : a.dim;
    const int trow = threadIdx.x;
    const long b = (long)panel * I8_N + trow;
    const bool active = b < a.n_batch;

    double zs[DM];
    kstar_i8_load_inputs<DM>(a, b, active, dim, zs);
    const int kern = a.kern[d];
    const int nkb = a.n_pad / I8_KB;
    if (kern_is_composite(kern)) return;   // block-uniform (d): composite outputs are built by kstar_i8_composite_kernel
    if (active) {
#pragma unroll
        for (int j = 0; j < DM; ++j)
            if (j < dim) zs[j] *= a.invls[d * dim + j];
    }

    double mu = 0.0;
    double jac[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) jac[j] = 0.0;

    const double var = a.var[d];
    const int row_begin = split * a.groups_per_split * 4;
    const int row_end = min(row_begin + a.groups_per_split * 4, a.n_pad);
    int8_t* panel_base = aa.ki8 + (((long)d * aa.npanel_cap + panel) * nkb) * (long)(I8_S * I8_B_TILE);

    // destination of this trajectory's bytes inside a k-block image
    const int rowp = trow;
    const long half_off = 0;
    const long plane_stride = I8_B_TILE;

    for (int row0 = row_begin; row0 < row_end; row0 += STAGE) {
        __syncthreads();
        const double* src = a.xs + ((long)d * a.n_pad + row0) * dim;
        for (int idx = threadIdx.x; idx < STAGE * dim; idx += I8_N) s_x[idx] = src[idx];
        for (int idx = threadIdx.x; idx < STAGE; idx += I8_N) s_beta[idx] = a.beta[(long)d * a.n_pad + row0 + idx] * var;
        __syncthreads();
        int8_t* kb_base = panel_base + half_off;
        // One loop body per kernel type: with both types inlined the 16-point body was 52 KB of SASS and stalled on
        // instruction fetch (ncu: no_instruction 1.1 per issue).  16 independent points per iteration are needed
        // to cover the FP64 latency (a 4-point body ran 40 % slower).
        if (kern == SEGP_KERN_RBF)
            kstar_i8_rows<DM, SEGP_KERN_RBF, 16, STAGE>(s_x, s_beta, s_tab, zs, dim, var, row0,
                                                                             a.n_train, active, kb_base, rowp,
                                                                             plane_stride, mu, jac);
        else
            kstar_i8_rows<DM, SEGP_KERN_MAT52, 8, STAGE>(s_x, s_beta, s_tab, zs, dim, var, row0,
                                                                              a.n_train, active, kb_base, rowp,
                                                                              plane_stride, mu, jac);
    }
    if (active) {
        a.mu_part[((long)split * n_s + d) * a.b_cap + b] = mu;
#pragma unroll
        for (int j = 0; j < DM; ++j)
            if (j < dim) {
                a.jac_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = jac[j];
                if (a.jac2_part != nullptr) a.jac2_part[(((long)split * n_s + d) * dim + j) * a.b_cap + b] = 0.0;
            }
        if (split == 0 && aa.colfac2 != nullptr) {   // a stationary output of a model that also has composite ones
            a.kss[(long)d * a.b_cap + b] = var;
            aa.colfac2[(long)d * a.b_cap + b] = 1.0;
        }
    }
}

// composite outputs of the model (the blocks of the other outputs exit at once)
template <int D_T>
__global__ void __launch_bounds__(I8_N) kstar_i8_composite_kernel(const KstarI8Args aa) {
    constexpr int DM = D_T > 0 ? D_T : MAX_D;
    __shared__ double s_x[comp_stage_rows(DM) * DM];
    __shared__ double s_xr[comp_stage_rows(DM) * DM];
    __shared__ double s_beta[TILE];
    __shared__ double s_tab[64];
    const KstarArgs& a = aa.k;
    const int d = (int)blockIdx.y;
    if (!kern_is_composite(a.kern[d])) return;
    if (threadIdx.x < 64) s_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 64.0));   // visible after the first
                                                                                           // __syncthreads of the stage loop
    const int dim = D_T > 0 ? D_T : a.dim;
    const int panel = aa.panel0 + (int)blockIdx.x;
    const long b = (long)panel * I8_N + threadIdx.x;
    const bool active = b < a.n_batch;
    double zs[DM];
    kstar_i8_load_inputs<DM>(a, b, active, dim, zs);
    int8_t* base = aa.ki8 + (((long)d * aa.npanel_cap + panel) * (a.n_pad / I8_KB)) * (long)(I8_S * I8_B_TILE);
    kstar_i8_composite<DM>(aa, zs, dim, d, (int)blockIdx.z, (int)gridDim.y, b, active, base, (int)threadIdx.x, s_x, s_xr,
                           s_beta, s_tab);
}

#ifndef SEGP_KS_MINB
#define SEGP_KS_MINB 1   // tuning experiments: minimum resident blocks per SM (caps the registers)
#endif
template <int D_T>
__global__ void __launch_bounds__(I8_N, SEGP_KS_MINB) kstar_i8_kernel(const KstarI8Args aa) {
    constexpr int DM = D_T > 0 ? D_T : MAX_D;
    __shared__ double s_x[TILE * DM];
    __shared__ double s_beta[TILE];
    __shared__ double s_tab[64];
    if (threadIdx.x < 64) s_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 64.0));   // visible after the first
                                                                                           // __syncthreads of the item
    kstar_i8_item<D_T, TILE>(aa, aa.panel0 + (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)gridDim.y, s_x,
                             s_beta, s_tab);
}

int launch_kstar_i8(const KstarI8Args& a, int n_s, int nsplit, cudaStream_t st) {
    // panels [panel0, ceil(n_batch / 96)): n_batch is the END of the trajectory range
    const int npanels = (int)((a.k.n_batch + I8_N - 1) / I8_N) - a.panel0;
    dim3 grid((unsigned)npanels, (unsigned)n_s, (unsigned)nsplit);
    dim3 block(I8_N);
    bool any_comp = false, all_comp = true;
    for (int d = 0; d < n_s; ++d) {
        any_comp = any_comp || kern_is_composite(a.k.kern[d]);
        all_comp = all_comp && kern_is_composite(a.k.kern[d]);
    }
    if (any_comp) {
        switch (a.k.dim) {
            case 2: kstar_i8_composite_kernel<2><<<grid, block, 0, st>>>(a); break;
            case 3: kstar_i8_composite_kernel<3><<<grid, block, 0, st>>>(a); break;
            case 4: kstar_i8_composite_kernel<4><<<grid, block, 0, st>>>(a); break;
            case 5: kstar_i8_composite_kernel<5><<<grid, block, 0, st>>>(a); break;
            default: kstar_i8_composite_kernel<0><<<grid, block, 0, st>>>(a);
        }
        SEGP_CUDA_CHECK(cudaGetLastError());
        if (all_comp) return SEGP_OK;
    }
    switch (a.k.dim) {
#define SEGP_KS8_CASE(D) \
    case D:              \
        kstar_i8_kernel<D><<<grid, block, 0, st>>>(a); \
        break;
        SEGP_KS8_CASE(2)
        SEGP_KS8_CASE(3)
        SEGP_KS8_CASE(4)
        SEGP_KS8_CASE(5)
        SEGP_KS8_CASE(6)
        SEGP_KS8_CASE(7)
        SEGP_KS8_CASE(8)
        SEGP_KS8_CASE(13)
#undef SEGP_KS8_CASE
        default:
            kstar_i8_kernel<0><<<grid, block, 0, st>>>(a);
    }
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== epilogue helper
// One warp, its 32 accumulator rows, 32 columns [col0, col0 + 32): recombine the I8_S diagonals
//   v = sum_g C_g 256^(S-1-g)   as   ((C0 256 + C1) 256^2 + (C2 256 + C3)) 256 + C4
// -- the two inner brackets exactly in int64 (one IMAD.WIDE each), the outer two steps as float64 FMAs (relative
// error 2^-53, far below the 2^-39 resolution of the digits) -- scale by the row factor, square, and transpose-
// reduce over the 32 rows: after the 5 halving steps lane l returns the sum of column col0 + l.
__device__ __forceinline__ double i8_epilogue_chunk(uint32_t tmem_quadrant_base, int col0, double rf, int lane,
                                                    int32_t* dbg_row, long dbg_plane_stride) {
    static_assert(I8_S == 5, "recombination below is written for 5 diagonals");
    uint32_t v[32];
    long long t01[32], t23[32];
    tmem_ld32(tmem_quadrant_base + (uint32_t)(0 * I8_N + col0), v);
    if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[0 * dbg_plane_stride + col0 + j] = (int)v[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) t01[j] = (long long)(int)v[j] * 256;
    tmem_ld32(tmem_quadrant_base + (uint32_t)(1 * I8_N + col0), v);
    if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[1 * dbg_plane_stride + col0 + j] = (int)v[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) t01[j] += (long long)(int)v[j];
    tmem_ld32(tmem_quadrant_base + (uint32_t)(2 * I8_N + col0), v);
    if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[2 * dbg_plane_stride + col0 + j] = (int)v[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) t23[j] = (long long)(int)v[j] * 256;
    tmem_ld32(tmem_quadrant_base + (uint32_t)(3 * I8_N + col0), v);
    if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[3 * dbg_plane_stride + col0 + j] = (int)v[j];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) t23[j] += (long long)(int)v[j];
    tmem_ld32(tmem_quadrant_base + (uint32_t)(4 * I8_N + col0), v);
    if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[4 * dbg_plane_stride + col0 + j] = (int)v[j];
    }
    double vals[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double hi = fma((double)t01[j], I8_BASE * I8_BASE, (double)t23[j]);
        const double x = fma(hi, I8_BASE, (double)(int)v[j]) * rf;
        vals[j] = x * x;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const double keep = up ? vals[j + off] : vals[j];
            const double give = up ? vals[j] : vals[j + off];
            vals[j] = keep + __shfl_xor_sync(0xffffffffu, give, off);
        }
    }
    return vals[0];
}

// Same recombination, same roundings, for the 12-warp epilogue of tri_i8m (one 32-column chunk per warp, three warps
// per TMEM lane quadrant, so the TMEM round trips of one warp hide behind the arithmetic of the other two):
//   * Horner entirely in float64:  a = C0; a = a 256 + C_g (g = 1..4).  The first three steps are exact (|a| < 2^47),
//     the fourth rounds the exact value t01 256^2 + t23 once and the fifth rounds once more -- the same two
//     roundings as i8_epilogue_chunk, hence bit-identical column sums;
//   * int32 -> float64 without I2F (a quarter-rate conversion, and the int64 flavour is slower still): the bit pattern
//     0x43300000:(x ^ 0x80000000) is 2^52 + 2^31 + x exactly, one full-rate DADD removes the bias;
//   * 64 + 32 live registers instead of 192, which is what lets 448 threads fit the register file.
// `werr` >= 0: also returns in `esum` this lane's column sum of werr_row * v^2 (float32: it is an error estimate).
__device__ __forceinline__ double i8_epilogue_chunk_fast(uint32_t tmem_quadrant_base, int col0, double rf, int lane,
                                                         float werr, float& esum) {
    static_assert(I8_S == 5, "recombination below is written for 5 diagonals");
    uint32_t v[32];
    double acc[32];
    tmem_ld32(tmem_quadrant_base + (uint32_t)(0 * I8_N + col0), v);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = i8_cvt_s32(v[j]);
#pragma unroll
    for (int g = 1; g < I8_S; ++g) {
        tmem_ld32(tmem_quadrant_base + (uint32_t)(g * I8_N + col0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma(acc[j], I8_BASE, i8_cvt_s32(v[j]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double x = acc[j] * rf;
        acc[j] = x * x;
    }
    if (werr >= 0.f) {
        float ef[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ef[j] = (float)acc[j] * werr;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
                const float keep = up ? ef[j + off] : ef[j];
                const float give = up ? ef[j] : ef[j + off];
                ef[j] = keep + __shfl_xor_sync(0xffffffffu, give, off);
            }
        }
        esum = ef[0];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const double keep = up ? acc[j + off] : acc[j];
            const double give = up ? acc[j] : acc[j + off];
            acc[j] = keep + __shfl_xor_sync(0xffffffffu, give, off);
        }
    }
    return acc[0];
}

// =========================================================================================== tri_i8
constexpr int I8_STAGES = 3;
constexpr int I8_STAGE_BYTES = I8_S * (I8_A_TILE + I8_B_TILE);   // 71680
constexpr int I8_THREADS = 192;                                  // producer, MMA issuer, 4 epilogue warps
constexpr int I8_TMEM_COLS = 512;
constexpr size_t I8_SMEM = (size_t)I8_STAGES * I8_STAGE_BYTES + 1024 /* alignment slack */ + 4 * I8_N * 8 + 128;
constexpr int I8_PANEL_GROUP = 24;   // panels whose K* digit planes (24 x 2.5 MB at N = 5000) stay L2-resident

__global__ void __launch_bounds__(I8_THREADS, 1) tri_i8_kernel(const TriI8Args a) {
    // ---- tile decode: (d, panel group) outer, block row descending (heavy first), panel inner
    int d, bi, panel;
    if (a.fix_bi >= 0) {
        d = 0;
        bi = a.fix_bi;
        panel = 0;
    } else {
        const int tiles_per_group = I8_PANEL_GROUP * a.nblk;
        const int npg = (a.npanels + I8_PANEL_GROUP - 1) / I8_PANEL_GROUP;
        const long gid = blockIdx.x / tiles_per_group;
        const int r = blockIdx.x % tiles_per_group;
        d = (int)(gid / npg);
        const int pg = (int)(gid % npg);
        bi = a.nblk - 1 - r / I8_PANEL_GROUP;
        panel = pg * I8_PANEL_GROUP + r % I8_PANEL_GROUP;
        if (panel >= a.npanels) return;
    }

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_addr(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;                    // SWIZZLE images need aligned tile bases
    unsigned char* tail = smem_raw + (stage0 - raw) + (size_t)I8_STAGES * I8_STAGE_BYTES;
    double* s_col = reinterpret_cast<double*>(tail);                   // [4][I8_N]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_col + 4 * I8_N);     // full[3], empty[3], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * I8_STAGES + 1);
    const uint32_t bar0 = smem_addr(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (I8_STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * I8_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "n"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nk = 2 * (bi + 1);   // k-blocks of 64 training points: columns 0 .. 128 (bi + 1) of the block row
    const int nkb_total = a.nblk * 2;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer: two bulk copies per stage
        if (lane == 0) {
            const int8_t* wsrc = a.wi8 + ((long)d * a.nblk * (a.nblk + 1) + (long)bi * (bi + 1)) * (I8_S * I8_A_TILE);
            const int8_t* ksrc = a.ki8 + (((long)d * a.npanel_cap + panel) * nkb_total) * (long)(I8_S * I8_B_TILE);
            for (int it = 0; it < nk; ++it) {
                const int s = it % I8_STAGES;
                if (it >= I8_STAGES) mbar_wait(empty_bar(s), (uint32_t)((it / I8_STAGES - 1) & 1));
                const uint32_t dst = stage0 + (uint32_t)s * I8_STAGE_BYTES;
                mbar_expect_tx(full_bar(s), I8_STAGE_BYTES);
                bulk_g2s(dst, wsrc + (long)it * (I8_S * I8_A_TILE), I8_S * I8_A_TILE, full_bar(s));
                bulk_g2s(dst + I8_S * I8_A_TILE, ksrc + (long)it * (I8_S * I8_B_TILE), I8_S * I8_B_TILE, full_bar(s));
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_i8_idesc(TILE, I8_N);
            for (int it = 0; it < nk; ++it) {
                const int s = it % I8_STAGES;
                mbar_wait(full_bar(s), (uint32_t)((it / I8_STAGES) & 1));
                tc_fence_after();
                const uint32_t sa = stage0 + (uint32_t)s * I8_STAGE_BYTES;
                const uint32_t sb = sa + I8_S * I8_A_TILE;
#pragma unroll
                for (int ks = 0; ks < I8_KB / 32; ++ks) {
#pragma unroll
                    for (int pa = 0; pa < I8_S; ++pa) {
                        const uint64_t adesc = make_sw64_desc(sa + pa * I8_A_TILE + ks * 32);
#pragma unroll
                        for (int pc = 0; pc < I8_S - pa; ++pc) {
                            const uint64_t bdesc = make_sw64_desc(sb + pc * I8_B_TILE + ks * 32);
                            // diagonal pa + pc accumulates in TMEM columns [(pa+pc) N, (pa+pc+1) N)
                            tc_mma_i8(tmem_base + (uint32_t)((pa + pc) * I8_N), adesc, bdesc, idesc,
                                      (uint32_t)((it | ks | pa) != 0));
                        }
                    }
                }
                tc_commit(empty_bar(s));   // frees the stage when these MMAs have read it
            }
            tc_commit(tmem_full_bar);
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 4 warps, one TMEM quadrant each
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const double rf = a.rowfac[((long)d * a.nblk + bi) * TILE + row];
        mbar_wait(tmem_full_bar, 0u);
        tc_fence_after();
        int32_t* dbg_row = a.dbg != nullptr ? a.dbg + (long)row * I8_N : nullptr;
#pragma unroll 1
        for (int chunk = 0; chunk < I8_N / 32; ++chunk)
            s_col[q * I8_N + chunk * 32 + lane] = i8_epilogue_chunk(tmem_base + ((uint32_t)(q * 32) << 16), chunk * 32, rf,
                                                                    lane, dbg_row, (long)TILE * I8_N);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int c = threadIdx.x - 64;
        if (c < I8_N) {
            const double sum = (s_col[c] + s_col[I8_N + c]) + (s_col[2 * I8_N + c] + s_col[3 * I8_N + c]);
            const long bcol = (long)panel * I8_N + c;
            if (bcol < a.b_cap)
                a.qpart[((long)d * a.nblk + bi) * a.b_cap + bcol] =
                    sum * (a.colfac2 != nullptr ? a.colfac2[(long)d * a.b_cap + bcol] : 1.0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(I8_TMEM_COLS)
                     : "memory");
    }
}

int tri_i8_init();

int launch_tri_i8(const TriI8Args& a, int n_s, cudaStream_t st) {
    long nblocks = 1;
    if (a.fix_bi < 0) {
        const int npg = (a.npanels + I8_PANEL_GROUP - 1) / I8_PANEL_GROUP;
        nblocks = (long)n_s * npg * I8_PANEL_GROUP * a.nblk;
    }
    if (nblocks <= 0 || nblocks > 2147483647L) {
        set_error("tri_i8: grid of %ld tiles out of range", nblocks);
        return SEGP_ERR_INVALID;
    }
    tri_i8_kernel<<<(unsigned)nblocks, I8_THREADS, I8_SMEM, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(local_bar),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "I8C_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra I8C_WAIT_DONE;\n"
        "bra I8C_WAIT_LOOP;\n"
        "I8C_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// =========================================================================================== tri_i8m (multicast)
// The single-CTA tile of tri_i8 (M = 128, N = 96: 58.6 clocks per MMA, bound by the 4 KB + 3 KB shared-memory operand
// read at ~122 B/clk, see segp_i8_peak_pattern) with the L2 -> SM fill cut from 70 to 50 KB per k-block: two CTAs of a
// cluster work on the SAME block row of W and adjacent trajectory panels, each fetches half of the W stage and
// multicasts it into both shared memories, and its own K* stage.
// Barriers: full (own expect_tx of the whole stage; bytes arrive from both producers), empty with CL arrivals (all
// MMA threads commit with the cluster multicast mask: a stage is rewritten only when every CTA has consumed it).
//
// SPLIT = false: classic digit set, 5 x 5 planes, pairs a + c < 5 (15 products, 9 MMAs per k-step).
// SPLIT = true : diagonal-split set (pack_w_i8): 4 W planes x 4 K* planes, pairs a + c < 4 (10 products, 6 MMAs per
//                k-step); in the two diagonal k-blocks of the block row additionally the plane of the diagonal's
//                leading digit d_-1 against all 5 K* planes (3 MMAs).  Accumulator slot s holds the diagonal
//                g = s - 1, so the epilogue (Horner over 5 slots) is the same code with another row factor.
// Stage layout (71680 B in both modes): [W planes: 5 x 8 KB | 4 x 8 KB + d_-1 tile 8 KB][K* planes 5 x 6 KB]; the split
// mode copies the d_-1 tile and the fifth K* plane in diagonal k-blocks only (57344 B otherwise).
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                   uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

constexpr int I8M_EPI_WARPS = 4 * (I8_N / 32);            // one warp per (TMEM lane quadrant, 32-column chunk)
constexpr int I8M_THREADS = 64 + 32 * I8M_EPI_WARPS;      // producer, MMA issuer, 12 epilogue warps = 448
constexpr uint32_t I8M_SB_OFF = I8_S * I8_A_TILE;         // K* planes start here in both stage layouts
constexpr uint32_t I8M_M1_OFF = I8_SS * I8_A_TILE;        // split mode: the d_-1 tile of a diagonal k-block

// Bulk copies of k-block `kb` of a block row into one stage.  wsrc: this block row's W planes (set in use), already
// advanced by rank * (this CTA's share) -- every CTA of the cluster fetches 1/CL of the W bytes and multicasts them.
// diag: 0 / 1 = first / second diagonal k-block of the block row (split mode: d_-1 tile + fifth K* plane), else -1.
template <bool SPLIT, int CL>
__device__ __forceinline__ void i8m_produce(uint32_t dst, uint32_t full_bar, const int8_t* wsrc, const int8_t* m1src,
                                            const int8_t* ksrc, int kb, int diag, uint32_t rank) {
    constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);
    if (!SPLIT) {
        constexpr uint32_t A_PART = I8_S * I8_A_TILE / CL;
        mbar_expect_tx(full_bar, I8_STAGE_BYTES);
        bulk_g2s_multicast(dst + rank * A_PART, wsrc + (long)kb * (I8_S * I8_A_TILE), A_PART, full_bar, MASK);
        bulk_g2s(dst + I8M_SB_OFF, ksrc + (long)kb * (I8_S * I8_B_TILE), I8_S * I8_B_TILE, full_bar);
    } else {
        constexpr uint32_t A_PART = I8_SS * I8_A_TILE / CL;
        constexpr uint32_t M1_PART = I8_A_TILE / CL;
        const uint32_t kbytes = (diag >= 0 ? I8_S : I8_SS) * I8_B_TILE;
        mbar_expect_tx(full_bar, I8_SS * I8_A_TILE + kbytes + (diag >= 0 ? I8_A_TILE : 0));
        bulk_g2s_multicast(dst + rank * A_PART, wsrc + (long)kb * (I8_SS * I8_A_TILE), A_PART, full_bar, MASK);
        if (diag >= 0)
            bulk_g2s_multicast(dst + I8M_M1_OFF + rank * M1_PART, m1src + (long)diag * I8_A_TILE + rank * M1_PART, M1_PART,
                               full_bar, MASK);
        bulk_g2s(dst + I8M_SB_OFF, ksrc + (long)kb * (I8_S * I8_B_TILE), kbytes, full_bar);
    }
}

// All MMAs of one k-block (two k-steps of 32).  Two K* planes per instruction where possible: planes pc and pc+1 are
// adjacent in the stage (2 x 96 rows) and their products with one W plane belong to adjacent accumulators, so ONE
// N = 192 MMA does both; an N = 192 instruction runs at the full pipe rate (100 clocks) instead of the
// operand-read-bound 58.6 of N = 96.  first: first k-block of the tile (accumulators are overwritten, not added to).
template <bool SPLIT>
__device__ __forceinline__ void i8m_issue(uint32_t tmem_base, uint32_t sa, bool first, int diag) {
    constexpr uint32_t idesc1 = make_i8_idesc(TILE, I8_N);
    constexpr uint32_t idesc2 = make_i8_idesc(TILE, 2 * I8_N);
    const uint32_t sb = sa + I8M_SB_OFF;
#pragma unroll
    for (int ks = 0; ks < I8_KB / 32; ++ks) {
        if (!SPLIT) {
#pragma unroll
            for (int pa = 0; pa < I8_S; ++pa) {
                const uint64_t adesc = make_sw64_desc(sa + pa * I8_A_TILE + ks * 32);
                const uint32_t acc = (uint32_t)(!first || ks != 0 || pa != 0);
#pragma unroll
                for (int pc = 0; pc < I8_S - pa; pc += 2) {
                    const uint64_t bdesc = make_sw64_desc(sb + pc * I8_B_TILE + ks * 32);
                    const bool two = pc + 1 < I8_S - pa;
                    tc_mma_i8(tmem_base + (uint32_t)((pa + pc) * I8_N), adesc, bdesc, two ? idesc2 : idesc1, acc);
                }
            }
        } else {
            // diagonals g = pa + pc in 0..3 live in slots 1..4; W plane 0 writes all four first
#pragma unroll
            for (int pa = 0; pa < I8_SS; ++pa) {
                const uint64_t adesc = make_sw64_desc(sa + pa * I8_A_TILE + ks * 32);
                const uint32_t acc = (uint32_t)(!first || ks != 0 || pa != 0);
#pragma unroll
                for (int pc = 0; pc < I8_SS - pa; pc += 2) {
                    const uint64_t bdesc = make_sw64_desc(sb + pc * I8_B_TILE + ks * 32);
                    const bool two = pc + 1 < I8_SS - pa;
                    tc_mma_i8(tmem_base + (uint32_t)((pa + pc + 1) * I8_N), adesc, bdesc, two ? idesc2 : idesc1, acc);
                }
            }
            if (diag >= 0) {
                // d_-1 plane against K* planes 0..4 -> diagonals -1..3 = slots 0..4.  Slot 0 is touched here only: its
                // first MMA (first k-step of the first diagonal k-block) overwrites.  Issued after the regular MMAs so
                // that slots 1..4 are initialised even when the diagonal k-block is the tile's first (block row 0).
                const uint64_t adesc = make_sw64_desc(sa + I8M_M1_OFF + ks * 32);
                tc_mma_i8(tmem_base, adesc, make_sw64_desc(sb + ks * 32), idesc1, (uint32_t)(diag != 0 || ks != 0));
                tc_mma_i8(tmem_base + (uint32_t)(1 * I8_N), adesc, make_sw64_desc(sb + 1 * I8_B_TILE + ks * 32), idesc2, 1u);
                tc_mma_i8(tmem_base + (uint32_t)(3 * I8_N), adesc, make_sw64_desc(sb + 3 * I8_B_TILE + ks * 32), idesc2, 1u);
            }
        }
    }
}

// CL = CTAs per cluster (2 or 4): the cluster works on CL adjacent panels of one block row; every CTA fetches 1/CL of
// the W stage and multicasts it to all, so the L2 -> SM fill per CTA and k-block is 40/CL + 30 KB (32/CL + 24 KB).
template <int CL, bool SPLIT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(I8M_THREADS, 1) tri_i8m_kernel(const TriI8Args a) {
    static_assert(CL == 2 || CL == 4, "cluster of 2 or 4 CTAs");
    const uint32_t rank = cluster_ctarank();
    const int pgroup = a.pgroup > 0 ? a.pgroup : I8_PANEL_GROUP;   // multiple of CL
    const int PG2 = pgroup / CL;   // clusters (CL adjacent panels) per L2 group
    int d, bi, panel;
    {
        const int cid = blockIdx.x / CL;
        const int tiles_per_group = PG2 * a.nblk;
        const int npg = (a.npanels - a.panel0 + pgroup - 1) / pgroup;
        const int gid = cid / tiles_per_group;
        const int r = cid % tiles_per_group;
        d = gid / npg;
        const int pg = gid % npg;
        bi = a.nblk - 1 - r / PG2;
        panel = a.panel0 + pg * pgroup + CL * (r % PG2) + (int)rank;
        if (panel - (int)rank >= a.npanels) return;   // the whole cluster is past the last panel
        if (a.pflag != nullptr) {                     // precision fallback: only clusters with a flagged panel run
            int any = 0;
            for (int r2 = 0; r2 < CL; ++r2) {
                const int p2 = panel - (int)rank + r2;
                if (p2 < a.npanels) any |= a.pflag[p2];
            }
            if (any == 0) return;
        }
    }
    const bool valid = panel < a.npanels;             // ragged panel count: the last cluster's spare CTAs only help loading
    const int panel_ld = valid ? panel : a.npanels - 1;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_addr(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    unsigned char* tail = smem_raw + (stage0 - raw) + (size_t)I8_STAGES * I8_STAGE_BYTES;
    double* s_col = reinterpret_cast<double*>(tail);
    float* s_ecol = reinterpret_cast<float*>(s_col + 4 * I8_N);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ecol + 4 * I8_N);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * I8_STAGES + 1);
    const uint32_t bar0 = smem_addr(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (I8_STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * I8_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CL);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "n"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nk = 2 * (bi + 1);
    const int nkb_total = a.nblk * 2;
    constexpr int NA = SPLIT ? I8_SS : I8_S;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);

    if (warp == 0) {
        if (lane == 0) {
            const int8_t* wsrc = a.wi8 + ((long)d * a.nblk * (a.nblk + 1) + (long)bi * (bi + 1)) * (NA * I8_A_TILE) +
                                 (long)rank * (NA * I8_A_TILE / CL);
            const int8_t* m1src = SPLIT ? a.wm1 + ((long)d * a.nblk + bi) * (2 * I8_A_TILE) : nullptr;
            const int8_t* ksrc = a.ki8 + (((long)d * a.npanel_cap + panel_ld) * nkb_total) * (long)(I8_S * I8_B_TILE);
            for (int it = 0; it < nk; ++it) {
                const int s = it % I8_STAGES;
                if (it >= I8_STAGES) mbar_wait_cluster(empty_bar(s), (uint32_t)((it / I8_STAGES - 1) & 1));
                i8m_produce<SPLIT, CL>(stage0 + (uint32_t)s * I8_STAGE_BYTES, full_bar(s), wsrc, m1src, ksrc, it,
                                       it - (nk - 2), rank);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % I8_STAGES;
                mbar_wait_cluster(full_bar(s), (uint32_t)((it / I8_STAGES) & 1));
                tc_fence_after();
                i8m_issue<SPLIT>(tmem_base, stage0 + (uint32_t)s * I8_STAGE_BYTES, it == 0, it - (nk - 2));
                tc_commit_multicast(empty_bar(s), CL_MASK);   // one of the CL arrivals on EVERY CTA's empty barrier
            }
            tc_commit(tmem_full_bar);
        }
    } else {
        // The epilogue is exposed (480 of the 512 TMEM columns hold accumulators: no second buffer for the next tile's
        // MMAs), so it is spread over 12 warps: warp -> (lane quadrant warp % 4 -- the hardware's TMEM access rule --,
        // 32-column chunk (warp - 2) / 4).
        const int q = warp & 3;
        const int chunk = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const long grow = ((long)d * a.nblk + bi) * TILE + row;
        const double rf = a.rowfac[grow];
        const float we = a.epart != nullptr ? a.werr[grow] : -1.f;
        mbar_wait(tmem_full_bar, 0u);
        tc_fence_after();
        float es = 0.f;
        s_col[q * I8_N + chunk * 32 + lane] =
            i8_epilogue_chunk_fast(tmem_base + ((uint32_t)(q * 32) << 16), chunk * 32, rf, lane, we, es);
        s_ecol[q * I8_N + chunk * 32 + lane] = es;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * I8M_EPI_WARPS) : "memory");
        const int c = threadIdx.x - 64;
        if (c < I8_N && valid) {
            const double sum = (s_col[c] + s_col[I8_N + c]) + (s_col[2 * I8_N + c] + s_col[3 * I8_N + c]);
            const long bcol = (long)panel * I8_N + c;
            if (bcol < a.b_cap) {
                const double cf = a.colfac2 != nullptr ? a.colfac2[(long)d * a.b_cap + bcol] : 1.0;
                a.qpart[((long)d * a.nblk + bi) * a.b_cap + bcol] = sum * cf;
                if (a.epart != nullptr)
                    a.epart[((long)d * a.nblk + bi) * a.b_cap + bcol] =
                        ((s_ecol[c] + s_ecol[I8_N + c]) + (s_ecol[2 * I8_N + c] + s_ecol[3 * I8_N + c])) * (float)(cf * cf);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();   // the peer may multicast into this shared memory / signal these barriers until it is done too
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(I8_TMEM_COLS)
                     : "memory");
    }
}

constexpr size_t I8M_SMEM = (size_t)I8_STAGES * I8_STAGE_BYTES + 1024 /* alignment slack */ + 4 * I8_N * 12 + 128;

int launch_tri_i8m(const TriI8Args& a, int n_s, cudaStream_t st) {
    const int cl = a.cluster == 4 ? 4 : 2;
    const int pgroup = a.pgroup > 0 ? a.pgroup : I8_PANEL_GROUP;
    if (pgroup % cl != 0) {
        set_error("tri_i8m: the panel group (%d) must be a multiple of the cluster size (%d)", pgroup, cl);
        return SEGP_ERR_INVALID;
    }
    const int npg = (a.npanels - a.panel0 + pgroup - 1) / pgroup;
    const long nclusters = (long)n_s * npg * (pgroup / cl) * a.nblk;
    if (nclusters <= 0 || cl * nclusters > 2147483647L) {
        set_error("tri_i8m: grid of %ld cluster tiles out of range", nclusters);
        return SEGP_ERR_INVALID;
    }
    const bool split = a.digits == 4;
    if (cl == 4) {
        if (split)
            tri_i8m_kernel<4, true><<<(unsigned)(4 * nclusters), I8M_THREADS, I8M_SMEM, st>>>(a);
        else
            tri_i8m_kernel<4, false><<<(unsigned)(4 * nclusters), I8M_THREADS, I8M_SMEM, st>>>(a);
    } else {
        if (split)
            tri_i8m_kernel<2, true><<<(unsigned)(2 * nclusters), I8M_THREADS, I8M_SMEM, st>>>(a);
        else
            tri_i8m_kernel<2, false><<<(unsigned)(2 * nclusters), I8M_THREADS, I8M_SMEM, st>>>(a);
    }
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== tri_i8mp (persistent tri_i8m)
// Same tiles, same arithmetic, same barriers as tri_i8m, but ONE resident cluster per TPC walks a static list of
// FOLDED tiles: a tile is (d, panel pair, fold f) and covers block rows nblk-1-f and f one after the other, so every
// tile holds exactly 2 (nblk + 1) k-blocks and a static round-robin assignment is balanced.  TMEM allocation, barrier
// initialisation and the cluster handshakes happen once per kernel; the producer runs ahead into the next block row
// while the epilogue of the current one drains TMEM, so the pipeline refill overlaps the (exposed) epilogue.  One
// more barrier: tmem_empty (one arrival per epilogue warp) gates the first MMA of the next block row.  What this buys
// is per-tile overhead: most at small N (C3: nblk = 16), little at C4/C5 where the board's power cap sets the pace.
struct MpSub {
    int d, bi, panel, panel_ld;
    bool valid;
};
// valid tile v of [0, n_s * nfold * npairs) -> (d, fold, panel pair); order: d, panel group of PG2 pairs (its K*
// planes stay L2-resident), fold, pair -- the clusters of one round work on neighbouring block rows of one group.
__device__ __forceinline__ void mp_decode(long v, int nfold, int npairs, int PG2, int& d, int& f, int& pair) {
    const long per_d = (long)nfold * npairs;
    d = (int)(v / per_d);
    const int r = (int)(v % per_d);
    const int full_groups = npairs / PG2;
    const int full = full_groups * nfold * PG2;
    if (r < full) {
        const int pg = r / (nfold * PG2);
        const int q = r % (nfold * PG2);
        f = q / PG2;
        pair = pg * PG2 + q % PG2;
    } else {
        const int rem = npairs - full_groups * PG2;
        const int q = r - full;
        f = q / rem;
        pair = full_groups * PG2 + q % rem;
    }
}

// 12 epilogue warps: one 32-column chunk each (448 threads, 128 registers: fills the register file).
template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8M_THREADS, 1) tri_i8mp_kernel(const TriI8Args a) {
    constexpr int EPI_WARPS = I8M_EPI_WARPS;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1;
    const int nclusters = gridDim.x >> 1;
    const int nfold = (a.nblk + 1) / 2;
    const int npairs = (a.npanels - a.panel0 + 1) / 2;
    const long ntiles = (long)a.fix_bi * nfold * npairs;    // fix_bi carries n_s (launch_tri_i8mp)

    if (a.pflag != nullptr) {
        // Precision fallback: the usual case is that the guard flagged nothing.  Every CTA reads the whole flag array
        // once (<= a few hundred ints, coalesced) and leaves before TMEM allocation, barrier set-up and the cluster
        // handshake; both CTAs of a cluster see the same flags, so they leave together.  Without this every role of
        // every cluster walked its tile list with one dependent L2 read per tile pair (~0.1 ms per step at C4).
        int any = 0;
        for (int i = a.panel0 + (int)threadIdx.x; i < a.npanels; i += (int)blockDim.x) any |= a.pflag[i];
        if (__syncthreads_or(any) == 0) return;
    }

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_addr(smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    unsigned char* tail = smem_raw + (stage0 - raw) + (size_t)I8_STAGES * I8_STAGE_BYTES;
    double* s_col = reinterpret_cast<double*>(tail);                        // [2][4][I8_N]
    float* s_ecol = reinterpret_cast<float*>(s_col + 2 * 4 * I8_N);         // [2][4][I8_N]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ecol + 2 * 4 * I8_N);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * I8_STAGES + 2);
    const uint32_t bar0 = smem_addr(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (I8_STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * I8_STAGES);
    const uint32_t tmem_empty_bar = bar0 + 8u * (2 * I8_STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 2);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "n"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nkb_total = a.nblk * 2;
    constexpr int NA = SPLIT ? I8_SS : I8_S;

    // the block rows of this cluster, in order; every role walks the same sequence
    auto sub_of = [&](long v, int which, MpSub& out) -> bool {
        int f, pair;
        mp_decode(v, nfold, npairs, (a.pgroup > 0 ? a.pgroup : I8_PANEL_GROUP) / 2, out.d, f, pair);
        const int bi_a = a.nblk - 1 - f;
        if (which == 1 && bi_a == f) return false;   // odd block-row count: the middle row is alone in its tile
        out.bi = which == 0 ? bi_a : f;
        out.panel = a.panel0 + 2 * pair + (int)rank;
        out.valid = out.panel < a.npanels;            // odd panel count: the second CTA of the last pair only helps loading
        out.panel_ld = out.valid ? out.panel : out.panel - 1;
        if (a.pflag != nullptr) {                     // precision fallback: only pairs with a flagged panel are computed
            const int p0 = a.panel0 + 2 * pair;       // (the same answer in every role and in both CTAs of the pair)
            if ((a.pflag[p0] | (p0 + 1 < a.npanels ? a.pflag[p0 + 1] : 0)) == 0) return false;
        }
        return true;
    };

    if (warp == 0) {
        if (lane == 0) {
            long it = 0;
            for (long v = cluster; v < ntiles; v += nclusters)
                for (int which = 0; which < 2; ++which) {
                    MpSub t;
                    if (!sub_of(v, which, t)) continue;
                    const int nk = 2 * (t.bi + 1);
                    const int8_t* wsrc = a.wi8 +
                                         ((long)t.d * a.nblk * (a.nblk + 1) + (long)t.bi * (t.bi + 1)) * (NA * I8_A_TILE) +
                                         (long)rank * (NA * I8_A_TILE / 2);
                    const int8_t* m1src = SPLIT ? a.wm1 + ((long)t.d * a.nblk + t.bi) * (2 * I8_A_TILE) : nullptr;
                    const int8_t* ksrc =
                        a.ki8 + (((long)t.d * a.npanel_cap + t.panel_ld) * nkb_total) * (long)(I8_S * I8_B_TILE);
                    for (int kb = 0; kb < nk; ++kb, ++it) {
                        const int s = (int)(it % I8_STAGES);
                        if (it >= I8_STAGES) mbar_wait_cluster(empty_bar(s), (uint32_t)((it / I8_STAGES - 1) & 1));
                        i8m_produce<SPLIT, 2>(stage0 + (uint32_t)s * I8_STAGE_BYTES, full_bar(s), wsrc, m1src, ksrc, kb,
                                              kb - (nk - 2), rank);
                    }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            long it = 0;
            uint32_t nsub = 0;
            for (long v = cluster; v < ntiles; v += nclusters)
                for (int which = 0; which < 2; ++which) {
                    MpSub t;
                    if (!sub_of(v, which, t)) continue;
                    const int nk = 2 * (t.bi + 1);
                    if (nsub > 0) {   // the epilogue of the previous block row has read its accumulators
                        mbar_wait(tmem_empty_bar, (nsub - 1) & 1u);
                        tc_fence_after();
                    }
                    for (int kb = 0; kb < nk; ++kb, ++it) {
                        const int s = (int)(it % I8_STAGES);
                        mbar_wait_cluster(full_bar(s), (uint32_t)((it / I8_STAGES) & 1));
                        tc_fence_after();
                        i8m_issue<SPLIT>(tmem_base, stage0 + (uint32_t)s * I8_STAGE_BYTES, kb == 0, kb - (nk - 2));
                        tc_commit_multicast(empty_bar(s), (uint16_t)3);
                    }
                    tc_commit(tmem_full_bar);
                    ++nsub;
                }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t nsub = 0;
        for (long v = cluster; v < ntiles; v += nclusters)
            for (int which = 0; which < 2; ++which) {
                MpSub t;
                if (!sub_of(v, which, t)) continue;
                const long grow = ((long)t.d * a.nblk + t.bi) * TILE + row;
                const double rf = a.rowfac[grow];
                const float we = a.epart != nullptr ? a.werr[grow] : -1.f;
                mbar_wait(tmem_full_bar, nsub & 1u);
                tc_fence_after();
                double* col = s_col + (nsub & 1u) * (4 * I8_N);   // double-buffered: the next block row's epilogue may
                float* ecol = s_ecol + (nsub & 1u) * (4 * I8_N);  // start while slow threads still read this one
                {
                    const int chunk = (warp - 2) >> 2;
                    float es = 0.f;
                    col[q * I8_N + chunk * 32 + lane] =
                        i8_epilogue_chunk_fast(tmem_base + ((uint32_t)(q * 32) << 16), chunk * 32, rf, lane, we, es);
                    ecol[q * I8_N + chunk * 32 + lane] = es;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar) : "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                const int c = threadIdx.x - 64;
                if (c < I8_N && t.valid) {
                    const double sum = (col[c] + col[I8_N + c]) + (col[2 * I8_N + c] + col[3 * I8_N + c]);
                    const long bcol = (long)t.panel * I8_N + c;
                    if (bcol < a.b_cap) {
                        const double cf = a.colfac2 != nullptr ? a.colfac2[(long)t.d * a.b_cap + bcol] : 1.0;
                        a.qpart[((long)t.d * a.nblk + t.bi) * a.b_cap + bcol] = sum * cf;
                        if (a.epart != nullptr)
                            a.epart[((long)t.d * a.nblk + t.bi) * a.b_cap + bcol] =
                                ((ecol[c] + ecol[I8_N + c]) + (ecol[2 * I8_N + c] + ecol[3 * I8_N + c])) * (float)(cf * cf);
                    }
                }
                ++nsub;
            }
    }
    tc_fence_before();
    cluster_sync_all();   // the peer may multicast into this shared memory / signal these barriers until it is done too
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(I8_TMEM_COLS)
                     : "memory");
    }
}

constexpr size_t I8MP_SMEM = (size_t)I8_STAGES * I8_STAGE_BYTES + 1024 /* alignment slack */ + 2 * 4 * I8_N * 12 + 128;

int launch_tri_i8mp(const TriI8Args& a, int n_s, cudaStream_t st) {
    const int nfold = (a.nblk + 1) / 2;
    const int npairs = (a.npanels - a.panel0 + 1) / 2;
    const long ntiles = (long)n_s * nfold * npairs;
    if (ntiles <= 0) {
        set_error("tri_i8mp: empty tile list");
        return SEGP_ERR_INVALID;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        SEGP_CUDA_CHECK(cudaGetDevice(&dev));
        SEGP_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const long nclusters = std::min<long>(ntiles, std::max(1, n_sm / 2));
    TriI8Args b = a;
    b.fix_bi = n_s;   // fix_bi (self-test tile selector of the other kernels) carries n_s into this one
    const bool split = a.digits == 4;
    const unsigned grid = (unsigned)(2 * nclusters);
    if (split)
        tri_i8mp_kernel<true><<<grid, I8M_THREADS, I8MP_SMEM, st>>>(b);
    else
        tri_i8mp_kernel<false><<<grid, I8M_THREADS, I8MP_SMEM, st>>>(b);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

int tri_i8_init() {
    SEGP_CUDA_CHECK(cudaFuncSetAttribute(tri_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM));
#define SEGP_I8_ATTR(K, BYTES) SEGP_CUDA_CHECK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)))
    SEGP_I8_ATTR((tri_i8m_kernel<2, false>), I8M_SMEM);
    SEGP_I8_ATTR((tri_i8m_kernel<2, true>), I8M_SMEM);
    SEGP_I8_ATTR((tri_i8m_kernel<4, false>), I8M_SMEM);
    SEGP_I8_ATTR((tri_i8m_kernel<4, true>), I8M_SMEM);
    SEGP_I8_ATTR((tri_i8mp_kernel<false>), I8MP_SMEM);
    SEGP_I8_ATTR((tri_i8mp_kernel<true>), I8MP_SMEM);
#undef SEGP_I8_ATTR
    return SEGP_OK;
}

// =========================================================================================== precision guard
// After a contraction on the 10-product digit set: per trajectory and output dimension,
//   sigma^2 = k** - sum_bi qpart,   e2 = sum_bi epart  (error-model variance of the |v|^2 just computed),
// a panel is flagged when  (2 kappa)^2 e2 > (rtol sigma^2)^2  for any of its trajectories (or sigma^2 <= 0): the
// 15-product kernel then recomputes the flagged panels (TriI8Args::pflag).  One block per 96-trajectory panel,
// blockDim.y = 4 splits the block rows; the flag needs no particular summation order.
__global__ void __launch_bounds__(4 * I8_N) i8_guard_kernel(const GuardArgs a) {
    __shared__ double s_q[4][I8_N];
    __shared__ float s_e[4][I8_N];
    const int panel = a.panel0 + (int)blockIdx.x;
    const long b = (long)panel * I8_N + threadIdx.x;
    int flag = 0;
    for (int d = 0; d < a.n_s; ++d) {
        double q = 0.0;
        float e = 0.f;
        if (b < a.n_batch) {
            for (int i = threadIdx.y; i < a.nblk; i += 4) {
                const long idx = ((long)d * a.nblk + i) * a.b_cap + b;
                q += a.qpart[idx];
                e += a.epart[idx];
            }
        }
        s_q[threadIdx.y][threadIdx.x] = q;
        s_e[threadIdx.y][threadIdx.x] = e;
        __syncthreads();
        if (threadIdx.y == 0 && b < a.n_batch) {
            const double qs = (s_q[0][threadIdx.x] + s_q[1][threadIdx.x]) + (s_q[2][threadIdx.x] + s_q[3][threadIdx.x]);
            const double es = (double)((s_e[0][threadIdx.x] + s_e[1][threadIdx.x]) + (s_e[2][threadIdx.x] + s_e[3][threadIdx.x]));
            const double s2 = (a.kss != nullptr ? a.kss[(long)d * a.b_cap + b] : a.gp_var[d]) - qs;
            if (!(s2 > 0.0) || a.gs * es > s2 * s2) flag = 1;
        }
        __syncthreads();
    }
    const int any = __syncthreads_or(flag);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        a.pflag[panel] = any;
        if (any && a.counter != nullptr) atomicAdd(a.counter, 1u);
    }
}

int launch_i8_guard(const GuardArgs& a, cudaStream_t st) {
    const int npanels = (int)((a.n_batch + I8_N - 1) / I8_N) - a.panel0;
    if (npanels <= 0) return SEGP_OK;
    i8_guard_kernel<<<(unsigned)npanels, dim3(I8_N, 4), 0, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

__global__ void scale_f32_kernel(float* __restrict__ x, long n, float f) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= f;
}
int launch_scale_f32(float* x, long n, float f, cudaStream_t st) {
    scale_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, f);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// =========================================================================================== i8_peak
// One CTA per SM issues `iters` back-to-back tcgen05.mma kind::i8 (M = 128, N = umma_n, K = 32) on fixed shared
// memory tiles; no loads, no epilogue: the sustained int8 tensor-pipe rate a kernel of this shape can reach.
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int umma_n, int iters, int pattern) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_addr(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* tiles = smem_raw + (base - raw);
    for (int i = threadIdx.x; i < I8_STAGE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x01010101u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(smem_addr(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_slot)),
                     "n"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy writes above must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_i8_idesc(TILE, umma_n);
        const uint64_t adesc0 = make_sw64_desc(base);
        const uint64_t bdesc0 = make_sw64_desc(base + I8_A_TILE);
        if (pattern == 0) {
            for (int it = 0; it < iters; ++it) {
                const uint32_t koff = (uint32_t)(it & 1) * 2u;                  // alternate the two k-steps of the tile
                const uint32_t col = (umma_n <= 256 && (it & 2)) ? 256u : 0u;   // and two accumulators
                tc_mma_i8(tmem_base + col, adesc0 + koff, bdesc0 + koff, idesc, (uint32_t)(it > 3));
            }
        } else if (pattern == 1) {   // one accumulator, back to back
            for (int it = 0; it < iters; ++it)
                tc_mma_i8(tmem_base, adesc0 + (uint32_t)(it & 1) * 2u, bdesc0 + (uint32_t)(it & 1) * 2u, idesc,
                          (uint32_t)(it > 0));
        } else {
            // the 15 digit-plane products of tri_i8 on its real stage layout (N must be I8_N):
            // pattern 2 = plane-major order (pa outer, pc inner: the accumulator changes every instruction),
            // pattern 3 = diagonal-major order (all products of one accumulator back to back)
            for (int it = 0; it < iters; it += 30) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    if (pattern == 2) {
#pragma unroll
                        for (int pa = 0; pa < I8_S; ++pa)
#pragma unroll
                            for (int pc = 0; pc < I8_S - pa; ++pc)
                                tc_mma_i8(tmem_base + (uint32_t)((pa + pc) * I8_N),
                                          make_sw64_desc(base + pa * I8_A_TILE + ks * 32),
                                          make_sw64_desc(base + I8_S * I8_A_TILE + pc * I8_B_TILE + ks * 32), idesc,
                                          (uint32_t)((it | ks | pa) != 0));
                    } else {
#pragma unroll
                        for (int g = 0; g < I8_S; ++g)
#pragma unroll
                            for (int pa = 0; pa <= g; ++pa)
                                tc_mma_i8(tmem_base + (uint32_t)(g * I8_N), make_sw64_desc(base + pa * I8_A_TILE + ks * 32),
                                          make_sw64_desc(base + I8_S * I8_A_TILE + (g - pa) * I8_B_TILE + ks * 32), idesc,
                                          (uint32_t)((it | ks | pa) != 0));
                    }
                }
            }
        }
        tc_commit(smem_addr(&bar));
        mbar_wait(smem_addr(&bar), 0u);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(I8_TMEM_COLS)
                     : "memory");
    }
}

int i8_peak(int umma_n, int iters, int pattern, double* tops) {
    if (umma_n < 16 || umma_n > 256 || umma_n % 16 != 0 || iters < 1 || tops == nullptr || pattern < 0 || pattern > 3 ||
        (pattern >= 2 && umma_n != I8_N)) {
        set_error("i8_peak: umma_n must be a multiple of 16 in [16, 256] (96 for patterns 2, 3), pattern in [0, 3]");
        return SEGP_ERR_INVALID;
    }
    if (pattern >= 2) iters = (iters + 29) / 30 * 30;
    int dev = 0, sms = 0;
    SEGP_CUDA_CHECK(cudaGetDevice(&dev));
    SEGP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = (size_t)I8_STAGE_BYTES + 1024;
    SEGP_CUDA_CHECK(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    SEGP_CUDA_CHECK(cudaEventCreate(&e0));
    SEGP_CUDA_CHECK(cudaEventCreate(&e1));
    i8_peak_kernel<<<sms, 128, smem>>>(umma_n, iters, pattern);   // warm-up
    SEGP_CUDA_CHECK(cudaEventRecord(e0));
    i8_peak_kernel<<<sms, 128, smem>>>(umma_n, iters, pattern);
    SEGP_CUDA_CHECK(cudaEventRecord(e1));
    SEGP_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    SEGP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SEGP_CUDA_CHECK(cudaGetLastError());
    *tops = 2.0 * TILE * umma_n * 32.0 * iters * sms / (ms * 1e-3) / 1e12;
    return SEGP_OK;
}

}  // namespace segp
