"""Float64 NumPy restatement of the constraint / cost assembly the reference's SafeMPC builds around the
reachability call, for ONE candidate at a time (loops; small cases only).

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

Follows (all citations relative to /root/reference/safe_exploration):

* ``SimpleSafeMPC.generate_safety_constraints``  safempc_simple.py:317-392  (order: control constraints,
  obstacle constraints on ellipsoids 0..H-2, terminal constraint on ellipsoid H-1)
* ``SimpleSafeMPC._generate_control_constraint`` safempc_simple.py:488-532  (u_0: plain bounds; later steps:
  ``lin_ellipsoid_safety_distance(k_ff, K_fb Q K_fb^T, [I;-I], [u_max;-u_min])``)
* ``SimpleSafeMPC.eval_safety_constraints``      safempc_simple.py:911-942  (feasible iff g < eps, eps = 1e-5)
* ``SimpleSafeMPC.generate_cost_function``       safempc_simple.py:286-315  (default branch, no performance
  trajectory: ``- sum_i sqrt(sum2(sigma_safe[i, :] + eps_noise))`` where ``sigma_safe`` holds the predictive
  VARIANCES returned by gp_reachability_casadi.onestep_reachability:79, 134)

``dist_fn`` is the safety-distance function to use: the restatement in reach_oracle by default, or the reference's
own ``gp_reachability.lin_ellipsoid_safety_distance`` (tests/test_oracle.py runs both and compares).

The two-sided bound on u_0 (``lbg = u_min, ubg = u_max`` in the reference) is expressed as the two one-sided
values ``u_0 - u_max`` and ``u_min - u_0`` so that every entry of g reads "feasible iff <= 0".

PARITY STATUS: the safety distance is pinned (reference function, live + golden); the assembly order and the cost
are a restatement of the cited lines (CasADi is absent, SimpleSafeMPC cannot be instantiated here).
"""
import numpy as np

from . import reach_oracle


def constraints_one(p_all, q_all, k_ff, k_fb, ctrl_bounds, h_mat_obs, h_obs, h_mat_safe, h_safe, dist_fn=None):
    """p_all (H,n_s), q_all (H,n_s,n_s), k_ff (H,n_u) with row 0 = u_0, k_fb (H-1,n_u,n_s) -> g (n_g,)"""
    dist = dist_fn or reach_oracle.lin_ellipsoid_safety_distance
    hor, n_s = p_all.shape
    n_u = k_ff.shape[1]
    g = []
    if ctrl_bounds is not None:
        u_min, u_max = ctrl_bounds[:, 0], ctrl_bounds[:, 1]
        g.append(k_ff[0] - u_max)
        g.append(u_min - k_ff[0])
        h_vec = np.vstack((u_max[:, None], -u_min[:, None]))
        h_mat = np.vstack((np.eye(n_u), -np.eye(n_u)))
        for i in range(hor - 1):
            q_u = k_fb[i] @ q_all[i] @ k_fb[i].T
            g.append(np.asarray(dist(k_ff[i + 1][:, None], q_u, h_mat, h_vec)).reshape(-1))
    if h_mat_obs is not None:
        for i in range(hor - 1):
            g.append(np.asarray(dist(p_all[i][:, None], q_all[i], h_mat_obs, h_obs)).reshape(-1))
    g.append(np.asarray(dist(p_all[-1][:, None], q_all[-1], h_mat_safe, h_safe)).reshape(-1))
    return np.concatenate(g)


def constraints_cautious_one(p_all, q_all, k_ff, k_fb, ctrl_bounds, h_mat_obs, h_obs, beta_safety, dist_fn=None):
    """CautiousMPC.generate_safety_constraints (cautious_mpc.py:337-395) with _generate_control_constraint (:397-442):
    u_0 bounds; per step i = 0..H-2 the control distances of (k_ff[i+1], K_fb[i] Sigma[i] K_fb[i]^T) with
    c_safety = beta_safety; then the obstacle distances of ALL H states with c_safety = beta_safety.  No terminal set.
    p_all (H,n_s) means, q_all (H,n_s,n_s) covariances, k_ff (H,n_u) with row 0 = u_0, k_fb (H-1,n_u,n_s)."""
    dist = dist_fn or reach_oracle.lin_ellipsoid_safety_distance
    hor = p_all.shape[0]
    n_u = k_ff.shape[1]
    g = []
    if ctrl_bounds is not None:
        u_min, u_max = ctrl_bounds[:, 0], ctrl_bounds[:, 1]
        g.append(k_ff[0] - u_max)
        g.append(u_min - k_ff[0])
        h_vec = np.vstack((u_max[:, None], -u_min[:, None]))
        h_mat = np.vstack((np.eye(n_u), -np.eye(n_u)))
        for i in range(hor - 1):
            q_u = k_fb[i] @ q_all[i] @ k_fb[i].T
            g.append(np.asarray(dist(k_ff[i + 1][:, None], q_u, h_mat, h_vec, beta_safety)).reshape(-1))
    if h_mat_obs is not None:
        for i in range(hor):
            g.append(np.asarray(dist(p_all[i][:, None], q_all[i], h_mat_obs, h_obs, beta_safety)).reshape(-1))
    return np.concatenate(g) if g else np.zeros(0)


def exploration_cost_one(var_all, eps_noise=0.0):
    return -float(np.sum(np.sqrt(np.sum(var_all + eps_noise, axis=1))))


def quadratic_cost_one(p_all, k_ff, wx, wu, x_ref=None):
    x_ref = np.zeros(p_all.shape[1]) if x_ref is None else x_ref
    c = 0.0
    for t in range(p_all.shape[0]):
        dx = p_all[t] - x_ref
        c += float(dx @ wx @ dx + k_ff[t] @ wu @ k_ff[t])
    return c


def score_batch(p_all, q_all, var_all, k_ff, k_fb, ctrl_bounds, h_mat_obs, h_obs, h_mat_safe, h_safe,
                cost="exploration", wx=None, wu=None, x_ref=None, eps_constraints=1e-5, eps_noise=0.0, dist_fn=None):
    """Loop over candidates.  Returns (cost (B,), feasible (B,) bool, violation (B,), g (B,n_g))."""
    bsz = p_all.shape[0]
    kfb_per = k_fb.ndim == 4
    g_all, costs = [], []
    for b in range(bsz):
        g_all.append(constraints_one(p_all[b], q_all[b], k_ff[b], k_fb[b] if kfb_per else k_fb, ctrl_bounds,
                                     h_mat_obs, h_obs, h_mat_safe, h_safe, dist_fn))
        if cost == "exploration":
            costs.append(exploration_cost_one(var_all[b], eps_noise))
        else:
            costs.append(quadratic_cost_one(p_all[b], k_ff[b], wx, wu, x_ref))
    g_all = np.array(g_all)
    viol = g_all.max(axis=1)
    return np.array(costs), viol < eps_constraints, viol, g_all
