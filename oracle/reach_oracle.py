"""Float64 NumPy restatement of the ellipsoid half of the hot path.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

Citations relative to /root/reference/safe_exploration:

* sum_two_ellipsoids                      utils_ellipsoid.py:63-94
* ellipsoid_from_rectangle                utils_ellipsoid.py:197-233
* compute_remainder_overapproximations    utils.py:108-144
* onestep_reachability                    gp_reachability.py:19-156
* multistep_reachability                  gp_reachability.py:159-212
* lin_ellipsoid_safety_distance           gp_reachability.py:215-250
* optional GP-input transform t_z_gp      gp_reachability_casadi.py:60-61,71,85,94-98

PARITY STATUS: pinned.  tests/test_oracle.py checks every function here against the
reference's own unmodified functions (oracle/ref_loader.py) on seeded inputs, against the
reference's known-answer tests (test/test_utils_ellipsoid.py:13-94) and against the golden
vectors in tests/golden/ that the reference code produced (oracle/make_golden.py).

One deliberate difference: the reference's ``scipy.linalg.eig`` leaks complex128 with a zero
imaginary part (utils.py:133-141); the oracle returns the real part (same values).
"""
import numpy as np


# ----------------------------------------------------------------------------- single
def sum_two_ellipsoids(p_1, q_1, p_2, q_2, c=None):
    """utils_ellipsoid.py:63-94 (trace-minimal outer ellipsoid of the Minkowski sum)."""
    if c is None:
        c = np.sqrt(np.trace(q_1) / np.trace(q_2))
    return p_1 + p_2, (1.0 + 1.0 / c) * q_1 + (1.0 + c) * q_2


def ellipsoid_from_rectangle(u_b):
    """utils_ellipsoid.py:197-233."""
    u_b = np.asarray(u_b)
    assert u_b.ndim == 1, "lb and ub need to be 1-dimensional (1darrays)!"
    assert np.all(u_b > 0), "all elements of u_b need to be greater than zero!"
    return np.diag(len(u_b) * u_b ** 2)


def compute_remainder_overapproximations(q, k_fb, l_mu, l_sigma):
    """utils.py:108-144."""
    n_u, n_s = np.shape(k_fb)
    s = np.hstack((np.eye(n_s), k_fb.T))
    b = s @ s.T
    evals = np.linalg.eigvals(q @ b)
    r_sqr = np.max(evals.real)
    return l_mu * r_sqr, l_sigma * np.sqrt(r_sqr)


def onestep_reachability(p_center, ssm, k_ff, l_mu, l_sigma, q_shape=None, k_fb=None,
                         c_safety=1., verbose=0, a=None, b=None, t_z_gp=None):
    """gp_reachability.py:19-156 (+ optional t_z_gp of gp_reachability_casadi.py)."""
    n_s = np.shape(p_center)[0]
    n_u = np.shape(k_ff)[0]
    if a is None:
        a = np.eye(n_s)
        b = np.zeros((n_s, n_u))
    if t_z_gp is None:
        t_z_gp = np.eye(n_s)
    x_bar = t_z_gp @ p_center
    if q_shape is None:
        mu_0, sigm_0, _ = ssm(x_bar.T, k_ff.T)
        mu_0 = np.array(mu_0)
        sigm_0 = np.array(sigm_0)
        rkhs_bounds = c_safety * np.sqrt(sigm_0.T).reshape((n_s,))
        q_1 = ellipsoid_from_rectangle(rkhs_bounds)
        p_1 = a @ p_center + b @ k_ff + mu_0
        return p_1, q_1
    mu_0, sigm_0, jac_mu = ssm(x_bar.T, k_ff.T)
    mu_0 = np.array(mu_0)
    sigm_0 = np.array(sigm_0)
    jac_mu = np.array(jac_mu)
    n_in = t_z_gp.shape[0]
    a_mu = jac_mu[:, :n_in] @ t_z_gp
    b_mu = jac_mu[:, n_in:]
    h = a + a_mu + (b_mu + b) @ k_fb
    p_0 = mu_0 + a @ p_center + b @ k_ff
    q_0 = h @ q_shape @ h.T
    ub_mean, ub_sigma = compute_remainder_overapproximations(q_shape, k_fb, l_mu, l_sigma)
    b_sigma_eps = c_safety * (np.sqrt(sigm_0.T) + ub_sigma)
    q_lagrange_sigm = ellipsoid_from_rectangle(b_sigma_eps.squeeze())
    q_lagrange_mu = ellipsoid_from_rectangle(ub_mean)
    zero = np.zeros((n_s, 1))
    p_sum, q_sum = sum_two_ellipsoids(zero, q_lagrange_sigm, zero, q_lagrange_mu)
    return sum_two_ellipsoids(p_sum, q_sum, p_0, q_0)


def multistep_reachability(p_0, gp, k_fb, k_ff, l_mu, l_sigm, q_0=None, c_safety=1.,
                           verbose=0, a=None, b=None, k_fb_init=None, t_z_gp=None):
    """gp_reachability.py:159-212."""
    n_, n_u, n_s = np.shape(k_fb)
    n = n_ + 1
    p_all = np.empty((n, n_s))
    q_all = np.empty((n, n_s, n_s))
    p_new, q_new = onestep_reachability(p_0, gp, k_ff[0, :, None], l_mu, l_sigm, q_0, k_fb_init,
                                        c_safety, verbose, a, b, t_z_gp)
    p_all[0] = p_new.T
    q_all[0] = q_new
    for i in range(1, n):
        p_new, q_new = onestep_reachability(p_new, gp, k_ff[i, :, None], l_mu, l_sigm, q_new,
                                            k_fb[i - 1, :, :], c_safety, verbose, a, b, t_z_gp)
        p_all[i] = p_new.T
        q_all[i] = q_new
    return p_new, q_new, p_all, q_all


def lin_ellipsoid_safety_distance(p_center, q_shape, h_mat, h_vec, c_safety=1.0):
    """gp_reachability.py:215-250."""
    d_center = h_mat @ p_center
    d_shape = c_safety * np.sqrt(np.sum((q_shape @ h_mat.T) * h_mat.T, axis=0)[:, None])
    return d_center + d_shape - h_vec


# ----------------------------------------------------------------------------- batched
def remainder_batch(q, k_fb, l_mu, l_sigma):
    """compute_remainder_overapproximations vectorised: q (B,n,n); k_fb (n_u,n) or (B,n_u,n)."""
    n_s = q.shape[-1]
    bm = np.eye(n_s) + np.swapaxes(k_fb, -1, -2) @ k_fb
    evals = np.linalg.eigvals(q @ bm)
    r_sqr = np.max(evals.real, axis=-1)
    return l_mu[None, :] * r_sqr[:, None], l_sigma[None, :] * np.sqrt(r_sqr)[:, None]


def onestep_batch(p, gp, k_ff, l_mu, l_sigma, q=None, k_fb=None, c_safety=1., a=None, b=None,
                  t_z_gp=None):
    """onestep_reachability for B independent trajectories.

    p (B,n_s); k_ff (B,n_u); q None or (B,n_s,n_s); k_fb (n_u,n_s) or (B,n_u,n_s);
    gp: object with predict_batch(z (B,D)) -> (mu (B,n_s), var (B,n_s), jac (B,n_s,D)).
    Returns p_1 (B,n_s), q_1 (B,n_s,n_s), var (B,n_s)."""
    bsz, n_s = p.shape
    n_u = k_ff.shape[1]
    if a is None:
        a = np.eye(n_s)
        b = np.zeros((n_s, n_u))
    x_bar = p if t_z_gp is None else p @ t_z_gp.T
    mu, var, jac = gp.predict_batch(np.hstack((x_bar, k_ff)))
    p_lin = p @ a.T + k_ff @ b.T + mu
    sig = np.sqrt(var)
    if q is None:
        d = n_s * (c_safety * sig) ** 2
        q_1 = np.zeros((bsz, n_s, n_s))
        idx = np.arange(n_s)
        q_1[:, idx, idx] = d
        return p_lin, q_1, var
    n_in = n_s if t_z_gp is None else t_z_gp.shape[0]
    a_mu = jac[:, :, :n_in] if t_z_gp is None else jac[:, :, :n_in] @ t_z_gp
    b_mu = jac[:, :, n_in:]
    h = a[None] + a_mu + (b_mu + b[None]) @ k_fb
    q_0 = h @ q @ np.swapaxes(h, 1, 2)
    ub_mean, ub_sigma = remainder_batch(q, k_fb, l_mu, l_sigma)
    d_sig = n_s * (c_safety * (sig + ub_sigma)) ** 2
    d_mu = n_s * ub_mean ** 2
    c1 = np.sqrt(d_sig.sum(1) / d_mu.sum(1))
    d_l = (1.0 + 1.0 / c1)[:, None] * d_sig + (1.0 + c1)[:, None] * d_mu
    c2 = np.sqrt(d_l.sum(1) / np.trace(q_0, axis1=1, axis2=2))
    q_1 = (1.0 + c2)[:, None, None] * q_0
    idx = np.arange(n_s)
    q_1[:, idx, idx] += (1.0 + 1.0 / c2)[:, None] * d_l
    return p_lin, q_1, var


def multistep_batch(p_0, gp, k_fb, k_ff, l_mu, l_sigm, q_0=None, c_safety=1., a=None, b=None,
                    k_fb_init=None, t_z_gp=None):
    """multistep_reachability for B trajectories.

    p_0 (n_s,) | (n_s,1) | (B,n_s); k_ff (B,H,n_u); k_fb (H-1,n_u,n_s) shared or (B,H-1,n_u,n_s);
    q_0 None | (n_s,n_s) | (B,n_s,n_s).  Returns p_all (B,H,n_s), q_all (B,H,n_s,n_s),
    var_all (B,H,n_s)."""
    k_ff = np.asarray(k_ff, dtype=np.float64)
    bsz, hor, n_u = k_ff.shape
    k_fb = np.asarray(k_fb, dtype=np.float64)
    n_s = k_fb.shape[-1]
    p = np.asarray(p_0, dtype=np.float64)
    p = np.broadcast_to(p.reshape(-1, n_s), (bsz, n_s)).copy() if p.size == n_s else p.reshape(bsz, n_s)
    q = None
    if q_0 is not None:
        q = np.broadcast_to(np.asarray(q_0, dtype=np.float64), (bsz, n_s, n_s)).copy()
    p_all = np.empty((bsz, hor, n_s))
    q_all = np.empty((bsz, hor, n_s, n_s))
    var_all = np.empty((bsz, hor, n_s))
    kfb_t = k_fb_init
    for t in range(hor):
        if t > 0:
            kfb_t = k_fb[t - 1] if k_fb.ndim == 3 else k_fb[:, t - 1]
        p, q, var = onestep_batch(p, gp, k_ff[:, t], l_mu, l_sigm, q, kfb_t, c_safety, a, b, t_z_gp)
        p_all[:, t] = p
        q_all[:, t] = q
        var_all[:, t] = var
    return p_all, q_all, var_all


def safety_distance_batch(p_all, q_all, h_mat, h_vec, c_safety=1.0):
    """lin_ellipsoid_safety_distance over leading axes: p (...,n), q (...,n,n) -> (...,m)."""
    d_center = p_all @ h_mat.T
    d_shape = c_safety * np.sqrt(np.einsum("mi,...ij,mj->...m", h_mat, q_all, h_mat))
    return d_center + d_shape - np.reshape(h_vec, (-1,))
