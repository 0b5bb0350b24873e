/* segp.h -- C ABI of libsegp.so: batched GP one-step posterior + ellipsoid reachability on B200.
 *
 * The reference (befelix/safe-exploration) is pure Python and has NO FFI for this path; the entry
 * points below are what a ctypes binding of that path binds (INTEGRATION.md shows the stub).  Each
 * one cites the reference interface it replaces; paths are relative to
 * /root/reference/safe_exploration/.
 *
 * Conventions
 *   - extern "C", opaque handle, int status returns (SEGP_OK == 0); no C++ / torch types.
 *   - all arithmetic and all buffers are IEEE float64 ("double"), row-major, contiguous.
 *   - pointers named d_* are DEVICE pointers on the handle's device; h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless stated otherwise; one handle is not re-entrant.
 *   - the handle owns the model buffers and its workspace; callers own every input/output buffer.
 *   - per-trajectory failures (non-finite or non-positive variance, a zero box bound) are RETURNED in
 *     the d_status bitmask, never raised: one diverging candidate must not kill a 65k batch.
 */
#ifndef SEGP_H_
#define SEGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGP_ABI_VERSION 2

/* status codes (Python wrapper maps them: INVALID->ValueError, CUDA->RuntimeError,
 * NOT_POSDEF->numpy.linalg.LinAlgError, NOT_TRAINED->RuntimeError, UNSUPPORTED->NotImplementedError) */
#define SEGP_OK 0
#define SEGP_ERR_INVALID 1
#define SEGP_ERR_CUDA 2
#define SEGP_ERR_NOT_POSDEF 3
#define SEGP_ERR_NOT_TRAINED 4
#define SEGP_ERR_UNSUPPORTED 5

/* kernel types, per output dimension (ssm_gpy/gp_models_utils_casadi.py:17-70, 218-231) */
#define SEGP_KERN_RBF 0
#define SEGP_KERN_MAT52 1
/* composite kernels of the journal configs, k = k_lin(product term) * k_stationary + k_lin(all inputs)
 * (_k_lin_rbf / _k_lin_mat52 / _k_lin, gp_models_utils_casadi.py:73-157; GPy objects gaussian_process.py:469-474),
 * evaluated as  k(x,y) = (sum_j a_j x_j y_j) s_f^2 phi(|(x - y) / l|) + sum_j v_j x_j y_j  with the vectors a, v of
 * segp_set_linear_terms.  Their values are signed and unbounded: on the int8 digit-plane path every trajectory is
 * scaled by s_b = s_f^2 sum_j |a_j z_j| max_i |x_ij| + sum_j |v_j z_j| max_i |x_ij| >= |k(z_b, x_i)| before the digit
 * split and the column sums of the contraction are multiplied by s_b^2. */
#define SEGP_KERN_LIN_RBF 2
#define SEGP_KERN_LIN_MAT52 3

/* per-trajectory status bits written to d_status */
#define SEGP_STATUS_NONFINITE 1    /* p or Q became inf/NaN                                  */
#define SEGP_STATUS_BAD_VARIANCE 2 /* predictive variance <= 0 or NaN (sqrt undefined)        */
#define SEGP_STATUS_ZERO_BOUND 4   /* a box bound was <= 0: the reference asserts here
                                      (utils_ellipsoid.py:226-228)                            */
#define SEGP_STATUS_LOW_PRECISION 8 /* int8 tensor-core path only: the a-posteriori error estimate of the variance
                                      contraction exceeds guard_rtol x sigma^2 even on the 15-product digit set, i.e.
                                      sigma^2 may miss the tolerance; re-run with tri_mode 0 (float64)             */

/* limits of this build */
#define SEGP_MAX_NS 16
#define SEGP_MAX_NU 8
#define SEGP_MAX_CONSTR 64 /* rows of one constraint polytope (segp_score_rollouts) */

typedef struct segp_model segp_model;

int segp_abi_version(void);

/* Last error message of the calling thread ("" if none). */
const char* segp_last_error(void);

/* Create an (untrained) model handle on CUDA device `device`.
 * Replaces SimpleGPModel.__init__ (ssm_gpy/gaussian_process.py:32-70): n_s_out independent GPs over
 * inputs [state(n_s_in), action(n_u)]; kern_type[n_s_out] in SEGP_KERN_*. */
int segp_create(segp_model** out, int device, int n_s_out, int n_s_in, int n_u, const int* kern_type);
int segp_destroy(segp_model* m);

/* Upload training data and hyper-parameters (HOST pointers; copies are taken).
 * Replaces the data/hyper-parameter half of SimpleGPModel.train (ssm_gpy/gaussian_process.py:189-278):
 * h_x [n_train x (n_s_in+n_u)], h_y [n_train x n_s_out], h_lengthscale [n_s_out x (n_s_in+n_u)] (ARD),
 * h_variance [n_s_out] (sigma_f^2), h_noise [n_s_out] = TOTAL diagonal term added to K
 * (GPy: Gaussian_noise.variance + noise_diag 1e-5 (gaussian_process.py:252-253) + 1e-8 jitter). */
int segp_set_model(segp_model* m, int n_train, const double* h_x, const double* h_y,
                   const double* h_lengthscale, const double* h_variance, const double* h_noise);

/* Linear terms of the composite kernels (HOST pointers, copies are taken); call after segp_set_model and before
 * segp_factorize whenever an output uses SEGP_KERN_LIN_*.  h_prod_linear [n_s_out x (n_s_in+n_u)]: weights a_j of the
 * linear factor of the product term; h_linear [n_s_out x (n_s_in+n_u)]: variances v_j of the additive linear kernel
 * (hyper-parameters "prod.linear.variances" / "linear.variances", gaussian_process.py:523-538).  Rows of
 * non-composite outputs are ignored.  For a composite output, h_lengthscale of segp_set_model may hold +inf for
 * input dimensions that do not enter the stationary factor: the CasADi form of the reference uses input column 1
 * only (gp_models_utils_casadi.py:82-95), the GPy kernel object all columns. */
int segp_set_linear_terms(segp_model* m, const double* h_prod_linear, const double* h_linear);

/* Build K_d = k_d(X,X)+noise_d I, Cholesky-factorise it, beta_d = K_d^-1 y_d, W_d = L_d^-1 packed for
 * the variance contraction -- all on the device in float64.  Synchronous (returns after completion).
 * Replaces the posterior half of SimpleGPModel.train / update_model
 * (ssm_gpy/gaussian_process.py:238-263, 396-413: GPRegression posterior, woodbury_inv/vector, pdinv).
 * SEGP_ERR_NOT_POSDEF if a pivot is not positive (LAPACK LinAlgError in the reference). */
int segp_factorize(segp_model* m, void* stream);

/* Append n_new training points (HOST pointers h_x [n_new x (n_s_in+n_u)], h_y [n_new x n_s_out]) to a factorised
 * model and update the posterior state.  Replaces SimpleGPModel.update_model(..., replace_old=False)
 * (ssm_gpy/gaussian_process.py:347-419: vstack + a fresh GPRegression, O(N^3)).  While the 128-padded size does not
 * grow, only the rows of W = L^-1 from the first touched 64-row block on are recomputed --
 * L21 = K21 W11^T, L22 L22^T = K22 - L21 L21^T, W22 = L22^-1, W21 = -W22 L21 W11, O(n_new N^2) -- then beta, the
 * log-determinant and the packed operands are refreshed.  This needs W kept dense on the device (n_s_out x
 * N_pad^2 doubles; option "keep_w", switched on by the first call, which therefore factorises from scratch, as
 * does every call that grows the padded size; read-only option "append_incremental" tells which path ran).
 * Synchronous.  SEGP_ERR_NOT_POSDEF as segp_factorize. */
int segp_append(segp_model* m, int n_new, const double* h_x, const double* h_y, void* stream);

/* Multi-GPU setup: the factorised state is ONE device allocation (buffer 0: beta, log-determinants, the int8 digit
 * planes of W = L^-1 with their row factors and error-model weights, and the decisions of the factorize-time probe).
 * Rank `root` factorises, every other rank calls segp_alloc_factor_buffers instead of segp_factorize, the host
 * broadcasts buffer 0 (one ncclBroadcast via torch.distributed), then the non-root ranks call segp_mark_factorized.
 * No collective is needed afterwards.  Only a model that runs the float64 contraction (composite kernels, more
 * than 16384 padded training points, tri_mode 0 / keep_fp64, or a probe that found the int8 path too coarse; option
 * "fp64_operand_needed") has a second buffer, the float64 DMMA operand: segp_num_factor_buffers == 2 on the
 * factorising rank; the others call segp_alloc_fp64_operand and receive buffer 1 before segp_mark_factorized. */
int segp_alloc_factor_buffers(segp_model* m);
int segp_alloc_fp64_operand(segp_model* m);
int segp_num_factor_buffers(segp_model* m);
int segp_factor_buffer(segp_model* m, int index, void** d_ptr, size_t* bytes);
int segp_mark_factorized(segp_model* m);

/* beta_d = (K_d + noise_d I)^-1 y_d, h_out [n_s_out x n_train] (HOST): posterior.woodbury_vector, what
 * SimpleGPModel keeps as self.beta (ssm_gpy/gaussian_process.py:261). */
int segp_beta(segp_model* m, double* h_out);

/* log det(K_d) per output dimension from the Cholesky factor (h_out[n_s_out], HOST).
 * Building block of SimpleGPModel.information_gain (ssm_gpy/gaussian_process.py:621-634). */
int segp_logdet(segp_model* m, double* h_out);

/* Greedy maximum-predicted-variance selection of m of the n rows of h_x (HOST pointers; synchronous): at every step
 * the row with the largest predictive variance, summed over the output dimensions, of the GPs conditioned on the rows
 * chosen so far (fixed hyper-parameters; ties to the lowest index; starts from the empty set).
 * Replaces the selection loop of SimpleGPModel.choose_datapoints_maxvar (ssm_gpy/gaussian_process.py:320-343:
 * predict over the pool, argmax of the summed variance, set_XY) without its k-means / random initial set and
 * without hyper-parameter re-optimisation.  h_noise [n_s_out] is the diagonal term of the conditioning (> 0);
 * h_prod_linear / h_linear as in segp_set_linear_terms (NULL unless a kernel is composite).
 * h_index [m] int32 out: selected rows in selection order; h_score [m] out or NULL: the summed variance of each
 * row at the moment it was selected. */
int segp_select_maxvar(int device, int n, int n_s_out, int dim, const int* kern_type, const double* h_x,
                       const double* h_lengthscale, const double* h_variance, const double* h_noise,
                       const double* h_prod_linear, const double* h_linear, int m, int32_t* h_index, double* h_score,
                       void* stream);

/* Batched predictive posterior at d_z [n_batch x (n_s_in+n_u)]:
 *   d_mu [n_batch x n_s_out], d_var [n_batch x n_s_out], d_jac [n_batch x n_s_out x (n_s_in+n_u)] or NULL.
 * Replaces SimpleGPModel.predict / predictive_gradients (ssm_gpy/gaussian_process.py:546-596), the
 * single-point SimpleGPModel.__call__ -> gp_pred_function (gp_models_utils_casadi.py:177-197, 234-288)
 * and StateSpaceModel.predict (state_space_models.py:74-104). */
int segp_predict(segp_model* m, long n_batch, const double* d_z, double* d_mu, double* d_var, double* d_jac,
                 void* stream);
/* The same with a per-input status word, d_status [n_batch] int32 out or NULL: SEGP_STATUS_BAD_VARIANCE and, on the
 * int8 tensor-core path, SEGP_STATUS_LOW_PRECISION (the reference has no counterpart: it is float64 throughout,
 * ssm_gpy/gaussian_process.py:546-568). */
int segp_predict_ex(segp_model* m, long n_batch, const double* d_z, double* d_mu, double* d_var, double* d_jac,
                    int32_t* d_status, void* stream);

/* what the n_s x n_s matrix carried along a rollout means */
#define SEGP_PROP_ELLIPSOID 0       /* ellipsoid shape matrix Q, gp_reachability.py:19-156 (default)                */
#define SEGP_PROP_TAYLOR 1          /* Gaussian covariance, first-order Taylor: one_step_taylor,
                                       uncertainty_propagation_casadi.py:11-87 / multi_step_taylor_symbolic :90-148   */
#define SEGP_PROP_MEAN_EQUIVALENT 2 /* Gaussian covariance, mean-equivalent: one_step_mean_equivalent :210-283 /
                                       mean_equivalent_multistep :151-207                                              */

/* Shared (not per-trajectory) parameters of the reachability recursion; HOST pointers. */
typedef struct segp_reach_params {
    const double* h_l_mu;    /* [n_s]  Lipschitz constants of the mean gradient (gp_reachability.py:38) */
    const double* h_l_sigma; /* [n_s]  Lipschitz constants of the std deviation (gp_reachability.py:40) */
    double c_safety;         /* gp_reachability.py:46                                                    */
    const double* h_a;       /* [n_s x n_s] linear prior, NULL = identity (gp_reachability.py:61-63)     */
    const double* h_b;       /* [n_s x n_u] linear prior, NULL = zero                                    */
    const double* h_t_z_gp;  /* [n_s_in x n_s] GP input transform or NULL (gp_reachability_casadi.py:60) */
    int propagation;         /* SEGP_PROP_*; with TAYLOR / MEAN_EQUIVALENT d_q0 / d_q_all hold covariances, l_mu,
                                l_sigma and c_safety are ignored (pass zeros)                                    */
} segp_reach_params;

/* n_batch independent H-step ellipsoid reachability recursions.
 * Replaces multistep_reachability (gp_reachability.py:159-212) and, with horizon == 1,
 * onestep_reachability (gp_reachability.py:19-156) -- including compute_remainder_overapproximations
 * (utils.py:108-144), ellipsoid_from_rectangle (utils_ellipsoid.py:197-233) and sum_two_ellipsoids
 * (utils_ellipsoid.py:63-94) -- for every trajectory at once.
 *   d_p0      [n_s] (p0_stride 0) or [n_batch x n_s] (p0_stride n_s): initial centres
 *   d_q0      NULL (start from a point, q_shape=None branch) or [n_s x n_s] (q0_stride 0) /
 *             [n_batch x n_s x n_s] (q0_stride n_s*n_s): initial shape matrices
 *   d_k_ff    [n_batch x horizon x n_u]
 *   d_k_fb    feedback gains for steps 1..horizon-1: [(horizon-1) x n_u x n_s] shared (kfb_stride 0) or
 *             [n_batch x (horizon-1) x n_u x n_s] (kfb_stride (horizon-1)*n_u*n_s); may be NULL if horizon==1
 *   d_k_fb_init  gain of step 0 (only read when d_q0 != NULL): [n_u x n_s] (kfb_init_stride 0) or per trajectory
 *   d_p_all   [n_batch x horizon x n_s]        out
 *   d_q_all   [n_batch x horizon x n_s x n_s]  out
 *   d_var_all [n_batch x horizon x n_s] out or NULL (predictive variances at the centres)
 *   d_status  [n_batch] int32 out or NULL (OR of SEGP_STATUS_* over the steps) */
int segp_multistep(segp_model* m, long n_batch, int horizon, const double* d_p0, long p0_stride,
                   const double* d_q0, long q0_stride, const double* d_k_ff, const double* d_k_fb,
                   long kfb_stride, const double* d_k_fb_init, long kfb_init_stride,
                   const segp_reach_params* params, double* d_p_all, double* d_q_all, double* d_var_all,
                   int32_t* d_status, void* stream);

/* Same call with HOST buffers: copies inputs to the device, runs, copies results back, synchronises.
 * This is the entry point a reference-side binding uses when it holds NumPy arrays. */
int segp_multistep_host(segp_model* m, long n_batch, int horizon, const double* h_p0, long p0_stride,
                        const double* h_q0, long q0_stride, const double* h_k_ff, const double* h_k_fb,
                        long kfb_stride, const double* h_k_fb_init, long kfb_init_stride,
                        const segp_reach_params* params, double* h_p_all, double* h_q_all, double* h_var_all,
                        int32_t* h_status);

/* One ellipsoid step for a foreign state-space model: the caller supplies the GP outputs.
 * Replaces the body of onestep_reachability after the ssm(...) call (gp_reachability.py:75-88, 102-156).
 *   d_mu, d_var [n_batch x n_s]; d_jac [n_batch x n_s x (n_s_in+n_u)] (ignored when d_q == NULL)
 *   d_p [n_batch x n_s]; d_q NULL or [n_batch x n_s x n_s]; d_k_ff [n_batch x n_u];
 *   d_k_fb [n_u x n_s] (kfb_stride 0) or [n_batch x n_u x n_s] (kfb_stride n_u*n_s), required when d_q != NULL */
int segp_ellipsoid_step(int device, long n_batch, int n_s, int n_s_in, int n_u, const double* d_mu,
                        const double* d_var, const double* d_jac, const double* d_p, const double* d_q,
                        const double* d_k_ff, const double* d_k_fb, long kfb_stride,
                        const segp_reach_params* params, double* d_p_out, double* d_q_out, int32_t* d_status,
                        void* stream);

/* Batched leaves of the ellipsoid calculus (each replaces the cited reference function, per item). */
/* utils.py:108-144 -- d_q [n_batch x n_s x n_s], d_k_fb [n_u x n_s] (stride 0) or per item; out [n_batch x n_s] */
int segp_remainder_overapproximations(int device, long n_batch, int n_s, int n_u, const double* d_q,
                                      const double* d_k_fb, long kfb_stride, const double* h_l_mu,
                                      const double* h_l_sigma, double* d_u_mu, double* d_u_sigma, void* stream);
/* utils_ellipsoid.py:63-94 (c=None) -- p [n_batch x n], q [n_batch x n x n] */
int segp_sum_two_ellipsoids(int device, long n_batch, int n, const double* d_p1, const double* d_q1,
                            const double* d_p2, const double* d_q2, double* d_p, double* d_q, void* stream);
/* utils_ellipsoid.py:197-233 -- d_ub [n_batch x n] -> d_q [n_batch x n x n]; status bit ZERO_BOUND if any ub <= 0 */
int segp_ellipsoid_from_rectangle(int device, long n_batch, int n, const double* d_ub, double* d_q,
                                  int32_t* d_status, void* stream);
/* gp_reachability.py:215-250 -- p [n_items x n_s], q [n_items x n_s x n_s], h_mat [m x n_s], h_vec [m] (HOST);
 * out d_dist [n_items x m] */
int segp_safety_distance(int device, long n_items, int n_s, int m, const double* d_p, const double* d_q,
                         const double* h_h_mat, const double* h_h_vec, double c_safety, double* d_dist,
                         void* stream);

/* ---- scoring of rolled-out candidates: what SimpleSafeMPC assembles symbolically around the reachability call ---- */
#define SEGP_COST_EXPLORATION 0 /* - sum_t sqrt(sum_d (var_d(t) + eps_noise))  (safempc_simple.py:303-305)       */
#define SEGP_COST_QUADRATIC 1   /* sum_t (p_t - x_ref)^T Wx (p_t - x_ref) + u_t^T Wu u_t                          */

typedef struct segp_score_params {
    const double* h_u_min;     /* [n_u] control bounds or NULL (ctrl_bounds[:,0], safempc_simple.py:488-532)       */
    const double* h_u_max;     /* [n_u]                                                                            */
    int m_obs;                 /* rows of the obstacle polytope h_mat_obs x <= h_obs (0 = none), steps 0..H-2      */
    const double* h_mat_obs;   /* [m_obs x n_s]  (safempc_simple.py:369-378)                                       */
    const double* h_obs;       /* [m_obs]                                                                          */
    int m_safe;                /* rows of the terminal safe-set polytope, step H-1 (safempc_simple.py:380-389)     */
    const double* h_mat_safe;  /* [m_safe x n_s]                                                                   */
    const double* h_safe;      /* [m_safe]                                                                         */
    double c_safety;           /* factor on the ellipsoid support term; the reference passes 1 here                */
    double eps_constraints;    /* feasible iff every g < eps_constraints (1e-5, safempc_simple.py:913)             */
    int cost_type;             /* SEGP_COST_*                                                                      */
    double eps_noise;          /* exploration cost                                                                 */
    const double* h_wx;        /* [n_s x n_s], quadratic cost                                                      */
    const double* h_wu;        /* [n_u x n_u]                                                                      */
    const double* h_x_ref;     /* [n_s] or NULL (origin)                                                           */
    int layout;                /* SEGP_SCORE_SAFEMPC (0, everything above) or SEGP_SCORE_CAUTIOUS                  */
    const double* h_q0;        /* [n_s x n_s] initial shape matrix shared by all candidates, or NULL: with it (the
                                  solver's init_uncertainty mode, safempc_simple.py:181-199, 350) the bound on u_0
                                  carries the support term of K_fb_0 Q_0 K_fb_0^T like every later control          */
    const double* h_k_fb_0;    /* [n_u x n_s] feedback gain of step 0; required with h_q0                           */
} segp_score_params;

/* constraint layouts */
#define SEGP_SCORE_SAFEMPC 0  /* SimpleSafeMPC.generate_safety_constraints, safempc_simple.py:317-392                */
#define SEGP_SCORE_CAUTIOUS 1 /* CautiousMPC.generate_safety_constraints, cautious_mpc.py:337-395: the obstacle polytope
                                 on ALL H states, no terminal set (m_safe ignored), and c_safety (beta_safety) on the
                                 support term of the control constraints too (_generate_control_constraint :397-442)  */

/* Number of constraint values per candidate: (has_ctrl ? 2 n_u H : 0) + (H-1) m_obs + m_safe
 * (cautious layout: (has_ctrl ? 2 n_u H : 0) + H m_obs). */
int segp_score_num_constraints(int horizon, int n_u, const segp_score_params* params);

/* Constraint values, feasibility and cost of n_batch rolled-out candidates (outputs of segp_multistep).
 * Replaces generate_safety_constraints (safempc_simple.py:317-392), _generate_control_constraint (:488-532),
 * eval_safety_constraints (:911-942) and the default generate_cost_function (:286-315), per candidate.
 *   d_p_all [B x H x n_s], d_q_all [B x H x n_s x n_s], d_var_all [B x H x n_s] (exploration cost; else may be NULL)
 *   d_k_ff [B x H x n_u] (row 0 is u_0), d_k_fb [(H-1) x n_u x n_s] (kfb_stride 0) or per candidate
 *   d_status [B] int32 from the rollout or NULL: a candidate with a non-zero status is infeasible
 *   out: d_cost [B], d_feasible [B] int32, d_violation [B] (max constraint value), d_g [B x n_g] or NULL */
int segp_score_rollouts(int device, long n_batch, int horizon, int n_s, int n_u, const double* d_p_all,
                        const double* d_q_all, const double* d_var_all, const double* d_k_ff, const double* d_k_fb,
                        long kfb_stride, const int32_t* d_status, const segp_score_params* params, double* d_cost,
                        int32_t* d_feasible, double* d_violation, double* d_g, void* stream);

/* Lowest-cost feasible candidate (or, if none is feasible, the least-violating one, *h_feasible = 0); synchronous.
 * Replaces the IPOPT solve + feasibility test of SimpleSafeMPC.solve/_get_solution (:672-742, 874-904) on the
 * sampling path.  *h_index = -1 if n_batch == 0 or every candidate is NaN. */
int segp_argbest(int device, long n_batch, const double* d_cost, const int32_t* d_feasible, const double* d_violation,
                 long* h_index, double* h_cost, double* h_violation, int* h_feasible, void* stream);

/* Diagnostic: sustained FP64 tensor-pipe (DMMA m8n8k4) rate of `device` in TFLOP/s, measured by a register-only
 * kernel (148 x 4 CTAs x 8 warps, `iters` x 64 independent DMMAs per warp).  bench.py reports tri_sumsq against it,
 * next to the bf16 figure MEASURED_PEAKS.json carries -- tcgen05 has no f64 kind, so this is the pipe the float64
 * contraction actually runs on. */
int segp_dmma_peak(int device, int iters, double* tflops);

/* Diagnostic: sustained int8 tensor-pipe rate (tcgen05.mma kind::i8, M = 128, N = umma_n, K = 32 issued back to
 * back by one thread per SM from fixed shared-memory tiles) in TOP/s: the denominator bench.py reports the int8
 * contraction kernel (tri_i8) against, next to twice the measured bf16 figure of MEASURED_PEAKS.json. */
int segp_i8_peak(int device, int umma_n, int iters, double* tops);
/* Same with a choice of issue pattern: 0 = two accumulators alternating (segp_i8_peak), 1 = one accumulator back to
 * back, 2 = the 15 digit-plane products of tri_i8 in plane-major order (accumulator changes every instruction),
 * 3 = the same products in diagonal-major order (umma_n must be 96 for 2 and 3). */
int segp_i8_peak_pattern(int device, int umma_n, int pattern, int iters, double* tops);

/* Diagnostic: run ONE block row of a tcgen05 contraction kernel on caller-supplied digit planes, so descriptor /
 * swizzle / TMEM-layout errors show up as exact integer mismatches.
 *   variant 1: reference kernel (one CTA per tile); 4: production kernel tri_i8m, classic digit set (15 products);
 *   6: tri_i8m, diagonal-split set (10 products + the diagonal's leading digit): plane 0 of h_a is that digit's plane
 *   and may be non-zero on the diagonal (row r, column K - 128 + r) only
 *   h_a [5][128][K] int8, h_b [5][96][K] int8 with K = 128 * k_blocks (HOST)
 *   h_acc [5][128][96] int32 (variant 1 only): h_acc[g] = sum over planes a + c == g of A_a B_c^T
 *   h_colsum [96]: sum over rows of (sum_g acc[g] * 256^(4-g))^2 as float64, where for variant 6
 *                  acc[g] = sum over planes a + c == g, a >= 1 or (a == 0), c <= 4, a + c <= 4 */
int segp_i8_selftest(int device, int variant, int k_blocks, const int8_t* h_a, const int8_t* h_b, int32_t* h_acc,
                     double* h_colsum);

/* Diagnostic: the float64-grade GEMM of the factorisation on tcgen05 (8 int8 digit planes per operand, 36 products;
 * csrc/fact_i8.cu) on host matrices: C = alpha A op(B) + beta C, A m x k, B n x k (trans_b != 0) or k x n, C m x n, all
 * row-major; m, n multiples of 128, k a multiple of 64; flags: 1 = A lower-triangular, 2 = B (k x n) lower-triangular,
 * 4 = only tiles on or below the diagonal.  Used by the tests to pin the kernel against NumPy (the reference does
 * these products inside LAPACK: GPy's Cholesky / woodbury_inv, ssm_gpy/gaussian_process.py:238-263). */
int segp_i8_gemm_selftest(int device, int m, int n, int k, const double* h_a, const double* h_b, double* h_c, double alpha,
                          double beta, int trans_b, int flags);

/* Tuning knobs / introspection ("chunk": trajectories per workspace chunk, "panel_group", "ksplit",
 * "tri_mode": which kernel runs the variance contraction: -1 = automatic (default: int8 digit planes on tcgen05 when
 *   the padded training size is <= 16384, the kernels are not composite and the factorize-time probe did not find the
 *   int8 path too coarse; else 0), 0 = fp64 DMMA (mma.sync m8n8k4.f64), 1 = int8 reference kernel (one CTA per tile,
 *   15 products; test cross-check), 4 = tri_i8m: single-CTA MMAs spanning two K* digit planes (N = 192) with the W
 *   stage multicast over a CTA pair, 5 = the same as a persistent kernel over folded (equal-length) tiles (automatic
 *   mode launches it instead of 4 for models of <= 32 block rows with enough tiles; read-only "tri_persistent" = 1 if
 *   the last launch did); read-only "tri_mode_effective";
 * "i8_digits": digit set of the first contraction pass: 0 = automatic (the probe's choice, read-only
 *   "i8_digits_effective"), 5 = classic set, 15 int8 products, 4 = diagonal-split set, 10 products;
 * "guard": 1 (default) = a-posteriori precision guard: on the 10-product set, panels whose error estimate exceeds
 *   guard_rtol x sigma^2 are recomputed on the 15-product set (read-only "fallback_panels" counts them; when more than
 *   a third of a rollout call's panel contractions were recomputed, automatic mode starts on the 15-product set from the
 *   next call on: read-only "demoted"); whatever still exceeds it gets SEGP_STATUS_LOW_PRECISION;  "probe": 1 (default) = calibrate the error model at segp_factorize;
 * "keep_fp64": keep the float64 operand resident after segp_factorize (else it is dropped unless needed, and a later
 *   tri_mode 0 factorises again); read-only "fp64_operand_resident", "fp64_operand_needed", "factor_bytes";
 * "graph": 1 (default) = replay the launches of a segp_multistep call as a CUDA graph from the second call with the
 *   same arguments on (read-only "graphs_cached");
 * "substreams": small models (<= 8 block rows of 128 training points) run a chunk as independent sub-batches on internal
 *   streams, because there every kernel is latency-bound and leaves most SMs idle (-1 automatic: 2; 0 off; n <= 4);
 * "fact_i8": the dense GEMMs of segp_factorize (trailing updates of potrf with K = 256, the block products of trtri from
 *   256-row blocks up) on tcgen05 from 8-digit int8 planes (csrc/fact_i8.cu): -1 = automatic (default: models of >= 1024
 *   padded points), 0 = float64 DMMA only, 1 = on from 512 points; read-only "fact_i8_effective";
 * "scratch_cache": keep the scratch of segp_factorize (K, L^-1, a temporary and the digit planes: ~0.9 GB per concurrent
 *   output dimension at N = 5000) between calls, so that a loop that refits a same-size model every step does not pay
 *   for allocating and freeing it: -1 = automatic (default: keep when it is <= 4 GB), 0 = never (frees it now), 1 =
 *   always; read-only "scratch_cached_bytes";
 * "time_tri": 1 = bracket every contraction launch with a CUDA-event pair on its stream (resets the counters);
 * read-only: "launches" = kernels launched by this handle so far, "n_train_padded", "workspace_bytes",
 * "tri_launches" / "tri_ns" = number and total device nanoseconds of the timed contraction launches). */
int segp_set_option(segp_model* m, const char* name, long value);
int segp_get_option(segp_model* m, const char* name, long* value);

/* Real-valued parameters: "guard_rtol" (1e-4: the tolerance the precision guard protects; BASELINE.json's rtol),
 * "guard_kappa" (5: standard deviations of the error model that must fit inside it -- an unflagged variance misses the
 * tolerance with probability < 6e-7 per evaluation, and far less away from the threshold); read-only statistics of the
 * factorize-time probe (1024 uniform inputs over the training box against the float64 contraction, worst output
 * dimension): "probe_ran", "probe_frac4" / "probe_frac5" (fraction the guard flags on the 10- / 15-product set),
 * "probe_err4/5" (max abs error of |L^-1 k*|^2), "probe_rel4/5" (max error / sigma^2), "probe_ratio4/5" (max error in
 * predicted standard deviations), "probe_rho4/5" (calibration factors applied), "probe_min_var_ratio", "probe_margin4" (largest kappa x estimated
 * error of the 10-product set relative to guard_rtol x sigma^2 over the probes: at most 0.25 means the first pass runs
 * without guard / recomputation launches, read-only option "unguarded"). */
int segp_set_param(segp_model* m, const char* name, double value);
int segp_get_param(segp_model* m, const char* name, double* value);

#ifdef __cplusplus
}
#endif
#endif /* SEGP_H_ */
