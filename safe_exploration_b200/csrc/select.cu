// Greedy maximum-predicted-variance data selection on the device.
//
// Replaces the selection loop of SimpleGPModel.choose_datapoints_maxvar (reference
// ssm_gpy/gaussian_process.py:280-345): at every step the pool point with the largest predictive variance, summed
// over the output dimensions, of the GPs conditioned on the points chosen so far.  The reference re-runs a full GPy
// predict over the pool per step (O(t^3 + n t^2) each); with fixed hyper-parameters the same arg-max sequence comes
// from one partial Cholesky factor per output dimension that grows by a column per step:
//
//   c_t(i)   = k(x_i, x_j) - sum_{s<t} l_s(i) l_s(j)          posterior covariance with the new point j = j_t
//   l_t(i)   = c_t(i) / sqrt(v(j) + noise)
//   v(i)    -= l_t(i)^2                                         posterior variance of every pool point
//
// O(n t) per step, HBM-bound on the l_s columns (stored [d][s][i], i fastest: coalesced).  Two launches per step:
// select_argmax (one block, fixed-order reduction, ties to the lowest index) and select_update (grid over i x d).
#include <math.h>

#include <vector>

#include "segp_internal.cuh"

namespace segp {

struct SelectArgs {
    const double* xs;      // [n_s][n][dim] inputs scaled by 1 / lengthscale_d
    const double* xraw;    // [n][dim]
    const double* plin;    // [n_s][dim] or NULL
    const double* lin;     // [n_s][dim] or NULL
    int kern[SEGP_MAX_NS];
    double var[SEGP_MAX_NS], noise[SEGP_MAX_NS];
    int n, dim, n_s;
    double* v;             // [n_s][n] current predictive variances
    double* cols;          // [n_s][m][n]
    int32_t* chosen;       // [n] 0/1
    int32_t* idx;          // [m] selected indices
    double* score;         // [m] summed variance at selection
    double* pivot;         // [n_s] v_d(j) of the current step
};

__device__ __forceinline__ double select_kernel_value(const SelectArgs& a, int d, int i, int j) {
    const double* xi = a.xs + ((long)d * a.n + i) * a.dim;
    const double* xj = a.xs + ((long)d * a.n + j) * a.dim;
    const int kern = a.kern[d];
    const bool composite = kern_is_composite(kern);
    double val;
    if (i == j) {
        val = a.var[d];
    } else {
        double r2 = 0.0;
        for (int c = 0; c < a.dim; ++c) {
            const double df = xi[c] - xj[c];
            r2 = fma(df, df, r2);
        }
        if (kern == SEGP_KERN_RBF || kern == SEGP_KERN_LIN_RBF) {
            val = a.var[d] * exp(-0.5 * r2);
        } else {
            const double sqrt5 = 2.23606797749978969641;
            const double rr = sqrt(r2);
            val = a.var[d] * (1.0 + sqrt5 * rr + (5.0 / 3.0) * r2) * exp(-sqrt5 * rr);
        }
    }
    if (composite) {
        double lp = 0.0, ll = 0.0;
        for (int c = 0; c < a.dim; ++c) {
            const double xx = a.xraw[(long)i * a.dim + c] * a.xraw[(long)j * a.dim + c];
            lp = fma(a.plin[d * a.dim + c], xx, lp);
            ll = fma(a.lin[d * a.dim + c], xx, ll);
        }
        val = fma(lp, val, ll);
    }
    return val;
}

__global__ void select_init_kernel(const SelectArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    for (int d = 0; d < a.n_s; ++d) a.v[(long)d * a.n + i] = select_kernel_value(a, d, i, i);
    a.chosen[i] = 0;
}

// arg-max over the pool of sum_d v_d(i), chosen points excluded; ties to the lowest index; NaN never wins
__global__ void __launch_bounds__(1024) select_argmax_kernel(const SelectArgs a, int t) {
    __shared__ double s_val[1024];
    __shared__ int s_idx[1024];
    double best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < a.n; i += 1024) {
        if (a.chosen[i]) continue;
        double s = 0.0;
        for (int d = 0; d < a.n_s; ++d) s += a.v[(long)d * a.n + i];
        if (!(s == s)) s = -INFINITY;
        if (bi == 0x7fffffff || s > best) {   // i ascends within a thread: ties keep the lower index
            best = s;
            bi = i;
        }
    }
    s_val[threadIdx.x] = best;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = s_val[threadIdx.x + o];
            const int oi = s_idx[threadIdx.x + o];
            if (oi != 0x7fffffff &&
                (s_idx[threadIdx.x] == 0x7fffffff || ov > s_val[threadIdx.x] ||
                 (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x]))) {
                s_val[threadIdx.x] = ov;
                s_idx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int j = s_idx[0];
        a.idx[t] = j;
        a.score[t] = s_val[0];
        a.chosen[j] = 1;
        for (int d = 0; d < a.n_s; ++d) a.pivot[d] = a.v[(long)d * a.n + j];
    }
}

__global__ void __launch_bounds__(256) select_update_kernel(const SelectArgs a, int t, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (i >= a.n) return;
    const int j = a.idx[t];
    double c = select_kernel_value(a, d, i, j);
    const double* col = a.cols + (long)d * m * a.n;
    for (int s = 0; s < t; ++s) c = fma(-col[(long)s * a.n + i], col[(long)s * a.n + j], c);
    const double l = c / sqrt(a.pivot[d] + a.noise[d]);
    a.cols[((long)d * m + t) * a.n + i] = l;
    a.v[(long)d * a.n + i] = fma(-l, l, a.v[(long)d * a.n + i]);
}

}  // namespace segp

using namespace segp;

extern "C" int segp_select_maxvar(int device, int n, int n_s_out, int dim, const int* kern_type, const double* h_x,
                                  const double* h_lengthscale, const double* h_variance, const double* h_noise,
                                  const double* h_prod_linear, const double* h_linear, int m, int32_t* h_index,
                                  double* h_score, void* stream) {
    if (n < 1 || m < 1 || m > n || n_s_out < 1 || n_s_out > SEGP_MAX_NS || dim < 1 || dim > MAX_D ||
        kern_type == nullptr || h_x == nullptr || h_lengthscale == nullptr || h_variance == nullptr ||
        h_noise == nullptr || h_index == nullptr) {
        set_error("segp_select_maxvar: bad argument (1 <= m <= n, dimensions within limits, non-null buffers)");
        return SEGP_ERR_INVALID;
    }
    bool composite = false;
    for (int d = 0; d < n_s_out; ++d) {
        if (kern_type[d] < SEGP_KERN_RBF || kern_type[d] > SEGP_KERN_LIN_MAT52) {
            set_error("segp_select_maxvar: unsupported kernel type %d", kern_type[d]);
            return SEGP_ERR_UNSUPPORTED;
        }
        composite = composite || kern_is_composite(kern_type[d]);
        if (!(h_variance[d] > 0.0) || !(h_noise[d] > 0.0)) {
            set_error("segp_select_maxvar: variance and noise must be > 0 (output %d)", d);
            return SEGP_ERR_INVALID;
        }
    }
    if (composite && (h_prod_linear == nullptr || h_linear == nullptr)) {
        set_error("segp_select_maxvar: composite kernel without linear terms");
        return SEGP_ERR_INVALID;
    }
    int ndev = 0;
    SEGP_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("segp_select_maxvar: no CUDA device %d; this library has no CPU path", device);
        return SEGP_ERR_CUDA;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    SEGP_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<double> xs((size_t)n_s_out * n * dim);
    for (int d = 0; d < n_s_out; ++d)
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < dim; ++c)
                xs[((size_t)d * n + i) * dim + c] = h_x[(size_t)i * dim + c] / h_lengthscale[d * dim + c];   // x / inf = 0
    SelectArgs a{};
    a.n = n;
    a.dim = dim;
    a.n_s = n_s_out;
    for (int d = 0; d < n_s_out; ++d) {
        a.kern[d] = kern_type[d];
        a.var[d] = h_variance[d];
        a.noise[d] = h_noise[d];
    }
    double *d_xs = nullptr, *d_xraw = nullptr, *d_pl = nullptr, *d_lin = nullptr, *d_v = nullptr, *d_cols = nullptr,
           *d_score = nullptr, *d_pivot = nullptr;
    int32_t *d_chosen = nullptr, *d_idx = nullptr;
    int rc = SEGP_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc != SEGP_OK) return;
        if (cudaMalloc(p, bytes) != cudaSuccess) {
            set_error("segp_select_maxvar: cudaMalloc of %zu bytes failed", bytes);
            rc = SEGP_ERR_CUDA;
        }
    };
    alloc((void**)&d_xs, xs.size() * 8);
    alloc((void**)&d_xraw, (size_t)n * dim * 8);
    alloc((void**)&d_pl, (size_t)n_s_out * dim * 8);
    alloc((void**)&d_lin, (size_t)n_s_out * dim * 8);
    alloc((void**)&d_v, (size_t)n_s_out * n * 8);
    alloc((void**)&d_cols, (size_t)n_s_out * m * n * 8);
    alloc((void**)&d_score, (size_t)m * 8);
    alloc((void**)&d_pivot, (size_t)n_s_out * 8);
    alloc((void**)&d_chosen, (size_t)n * 4);
    alloc((void**)&d_idx, (size_t)m * 4);
    if (rc == SEGP_OK) {
        cudaError_t e = cudaMemcpyAsync(d_xs, xs.data(), xs.size() * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_xraw, h_x, (size_t)n * dim * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && composite) {
            e = cudaMemcpyAsync(d_pl, h_prod_linear, (size_t)n_s_out * dim * 8, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_lin, h_linear, (size_t)n_s_out * dim * 8, cudaMemcpyHostToDevice, st);
        }
        a.xs = d_xs;
        a.xraw = d_xraw;
        a.plin = composite ? d_pl : nullptr;
        a.lin = composite ? d_lin : nullptr;
        a.v = d_v;
        a.cols = d_cols;
        a.chosen = d_chosen;
        a.idx = d_idx;
        a.score = d_score;
        a.pivot = d_pivot;
        if (e == cudaSuccess) {
            select_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(a);
            dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_s_out);
            for (int t = 0; t < m; ++t) {
                select_argmax_kernel<<<1, 1024, 0, st>>>(a, t);
                select_update_kernel<<<grid, 256, 0, st>>>(a, t, m);
            }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_index, d_idx, (size_t)m * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && h_score != nullptr)
            e = cudaMemcpyAsync(h_score, d_score, (size_t)m * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("segp_select_maxvar: %s", cudaGetErrorString(e));
            rc = SEGP_ERR_CUDA;
        }
    }
    cudaFree(d_xs);
    cudaFree(d_xraw);
    cudaFree(d_pl);
    cudaFree(d_lin);
    cudaFree(d_v);
    cudaFree(d_cols);
    cudaFree(d_score);
    cudaFree(d_pivot);
    cudaFree(d_chosen);
    cudaFree(d_idx);
    cudaSetDevice(prev);
    return rc;
}
