// Scoring of B rolled-out candidates: the constraint and cost assembly of the reference's SafeMPC, per candidate.
//   control constraints   SimpleSafeMPC._generate_control_constraint (safempc_simple.py:488-532):
//                         step 0: u_min <= u_0 <= u_max (the start is a point, no feedback term; with an initial
//                         ellipsoid Q_0, K_fb_0 -- init_uncertainty -- the same support term as the later steps);
//                         step i+1: lin_ellipsoid_safety_distance(k_ff[i+1], K_fb[i] Q[i] K_fb[i]^T, [I;-I], [u_max;-u_min])
//   state constraints     generate_safety_constraints (:317-392): obstacle polytope on ellipsoids 0..H-2,
//                         terminal safe-set polytope on ellipsoid H-1, both through
//                         lin_ellipsoid_safety_distance (gp_reachability.py:215-250) with c = 1
//   feasibility           eval_safety_constraints (:911-942): every g < eps_constraints
//   cost                  generate_cost_function (:286-315) default branch without a performance trajectory:
//                         - sum_t sqrt(sum_d (var_d(t) + eps_noise)); or a quadratic tracking cost
// One thread per candidate (the data are ~2.6 kB per candidate at H = 20, n_s = 4: an HBM-bound pass that costs
// microseconds next to the rollout), then a single-block arg-best reduction.
#include <math.h>

#include "segp_internal.cuh"

namespace segp {

__global__ void score_kernel(const ScoreArgs a) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_batch) return;
    const ScoreParams& sp = *a.sp;
    const int n_s = a.n_s, n_u = a.n_u, hor = a.horizon;
    const double* p_all = a.p_all + b * (long)hor * n_s;
    const double* q_all = a.q_all + b * (long)hor * n_s * n_s;
    const double* kff = a.kff + b * (long)hor * n_u;
    double* g = a.g != nullptr ? a.g + b * (long)a.n_g : nullptr;
    int gi = 0;
    double viol = -INFINITY;
    bool finite = true;
    auto emit = [&](double v) {
        if (g != nullptr) g[gi] = v;
        ++gi;
        viol = fmax(viol, v);
        if (!isfinite(v)) finite = false;
    };
    // CautiousMPC (cautious_mpc.py:337-442): beta_safety also scales the support term of the control constraints, the
    // obstacle polytope applies to all H states and there is no terminal set
    const bool cautious = sp.layout == SEGP_SCORE_CAUTIOUS;
    const double c_ctrl = cautious ? sp.c_safety : 1.0;
    // ---- control constraints
    if (sp.has_ctrl) {
        // step 0: the start is a point (no feedback term) unless an initial ellipsoid Q_0 with its gain K_fb_0 is given
        // (init_uncertainty, safempc_simple.py:350: _generate_control_constraint(u_0, q_0, k_fb_0))
        double sd0[SEGP_MAX_NU];
        for (int j = 0; j < n_u; ++j) {
            double acc = 0.0;
            if (sp.has_q0)
                for (int r = 0; r < n_s; ++r) {
                    double t = 0.0;
                    for (int c = 0; c < n_s; ++c) t = fma(sp.q0[r * n_s + c], sp.kfb0[j * n_s + c], t);
                    acc = fma(sp.kfb0[j * n_s + r], t, acc);
                }
            sd0[j] = c_ctrl * sqrt(acc);
        }
        for (int j = 0; j < n_u; ++j) emit(kff[j] + sd0[j] - sp.u_max[j]);
        for (int j = 0; j < n_u; ++j) emit(sp.u_min[j] - kff[j] + sd0[j]);
        for (int i = 0; i + 1 < hor; ++i) {
            const double* q = q_all + (long)i * n_s * n_s;
            const double* kfb = a.kfb + b * a.kfb_stride + (long)i * n_u * n_s;
            const double* u = kff + (long)(i + 1) * n_u;
            double sd[SEGP_MAX_NU];
            for (int j = 0; j < n_u; ++j) {
                double acc = 0.0;   // (K Q K^T)_jj
                for (int r = 0; r < n_s; ++r) {
                    double t = 0.0;
                    for (int c = 0; c < n_s; ++c) t = fma(q[r * n_s + c], kfb[j * n_s + c], t);
                    acc = fma(kfb[j * n_s + r], t, acc);
                }
                sd[j] = sqrt(acc);
            }
            for (int j = 0; j < n_u; ++j) emit(u[j] + c_ctrl * sd[j] - sp.u_max[j]);
            for (int j = 0; j < n_u; ++j) emit(-u[j] + c_ctrl * sd[j] + sp.u_min[j]);
        }
    }
    // ---- polytope constraints on the state ellipsoids
    auto polytope = [&](const double* p, const double* q, const double* hmat, const double* hvec, int m) {
        for (int k = 0; k < m; ++k) {
            const double* h = hmat + k * n_s;
            double hp = 0.0, hqh = 0.0;
            for (int r = 0; r < n_s; ++r) {
                hp = fma(h[r], p[r], hp);
                double t = 0.0;
                for (int c = 0; c < n_s; ++c) t = fma(q[r * n_s + c], h[c], t);
                hqh = fma(h[r], t, hqh);
            }
            emit(hp + sp.c_safety * sqrt(hqh) - hvec[k]);
        }
    };
    for (int i = 0; i + (cautious ? 0 : 1) < hor; ++i)
        polytope(p_all + (long)i * n_s, q_all + (long)i * n_s * n_s, sp.h_mat_obs, sp.h_obs, sp.m_obs);
    if (!cautious)
        polytope(p_all + (long)(hor - 1) * n_s, q_all + (long)(hor - 1) * n_s * n_s, sp.h_mat_safe, sp.h_safe,
                 sp.m_safe);
    // ---- cost
    double cost = 0.0;
    if (sp.cost_type == SEGP_COST_EXPLORATION) {
        const double* var = a.var_all + b * (long)hor * n_s;
        for (int t = 0; t < hor; ++t) {
            double s = 0.0;
            for (int d = 0; d < n_s; ++d) s += var[t * n_s + d] + sp.eps_noise;
            cost -= sqrt(s);
        }
    } else {
        for (int t = 0; t < hor; ++t) {
            const double* p = p_all + (long)t * n_s;
            const double* u = kff + (long)t * n_u;
            for (int r = 0; r < n_s; ++r) {
                double acc = 0.0;
                for (int c = 0; c < n_s; ++c) acc = fma(sp.wx[r * n_s + c], p[c] - sp.x_ref[c], acc);
                cost = fma(p[r] - sp.x_ref[r], acc, cost);
            }
            for (int r = 0; r < n_u; ++r) {
                double acc = 0.0;
                for (int c = 0; c < n_u; ++c) acc = fma(sp.wu[r * n_u + c], u[c], acc);
                cost = fma(u[r], acc, cost);
            }
        }
    }
    if (!isfinite(cost)) finite = false;
    const bool ok_status = a.status == nullptr || a.status[b] == 0;
    a.cost[b] = cost;
    a.violation[b] = viol;
    a.feasible[b] = (finite && ok_status && viol < sp.eps_constraints) ? 1 : 0;
}

int launch_score(const ScoreArgs& a, cudaStream_t st) {
    const int threads = 128;
    score_kernel<<<(unsigned)((a.n_batch + threads - 1) / threads), threads, 0, st>>>(a);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

// Best candidate: lowest cost among the feasible ones; if none is feasible, the one with the smallest
// violation (reported with feasible = 0).  Ties resolve to the lowest index (deterministic).
__global__ void argbest_kernel(long n, const double* __restrict__ cost, const int32_t* __restrict__ feasible,
                               const double* __restrict__ violation, BestCandidate* out) {
    __shared__ double s_key[256];
    __shared__ long s_idx[256];
    __shared__ int s_feas[256];
    double best = INFINITY;
    long idx = -1;
    int feas = 0;
    for (long i = threadIdx.x; i < n; i += 256) {
        const int f = feasible[i];
        const double key = f ? cost[i] : violation[i];
        if (isnan(key)) continue;
        if (f > feas || (f == feas && (key < best || idx < 0))) {
            best = key;
            idx = i;
            feas = f;
        }
    }
    s_key[threadIdx.x] = best;
    s_idx[threadIdx.x] = idx;
    s_feas[threadIdx.x] = feas;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int j = threadIdx.x + o;
            const bool take = s_idx[j] >= 0 &&
                              (s_idx[threadIdx.x] < 0 || s_feas[j] > s_feas[threadIdx.x] ||
                               (s_feas[j] == s_feas[threadIdx.x] &&
                                (s_key[j] < s_key[threadIdx.x] ||
                                 (s_key[j] == s_key[threadIdx.x] && s_idx[j] < s_idx[threadIdx.x]))));
            if (take) {
                s_key[threadIdx.x] = s_key[j];
                s_idx[threadIdx.x] = s_idx[j];
                s_feas[threadIdx.x] = s_feas[j];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out->index = s_idx[0];
        out->feasible = s_feas[0];
        out->cost = s_idx[0] >= 0 ? cost[s_idx[0]] : INFINITY;
        out->violation = s_idx[0] >= 0 ? violation[s_idx[0]] : INFINITY;
    }
}

int launch_argbest(long n, const double* cost, const int32_t* feasible, const double* violation, BestCandidate* out,
                   cudaStream_t st) {
    argbest_kernel<<<1, 256, 0, st>>>(n, cost, feasible, violation, out);
    SEGP_CUDA_CHECK(cudaGetLastError());
    return SEGP_OK;
}

}  // namespace segp
