"""Float64 NumPy restatement of the reference's Gaussian uncertainty propagation, vectorised over B trajectories.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

Follows /root/reference/safe_exploration/uncertainty_propagation_casadi.py:
* ``one_step_taylor``            :11-87    first-order Taylor propagation of N(mu_x, sigma_x) through prior + GP
* ``multi_step_taylor_symbolic`` :90-148
* ``one_step_mean_equivalent``   :210-283  same without the linearisation term
* ``mean_equivalent_multistep``  :151-207

The reference builds the joint covariance of (x, u, g) and maps it through [A B I] (:59-87).  With u = K x + k the joint
covariance of z = (x, u) is F Sigma F^T, F = [I; K], and Sigma_zg = Sigma_z J^T, so
    [A B I] Sigma_all [A B I]^T = (A + B K + J F) Sigma (A + B K + J F)^T + diag(sigma_g^2);
this file evaluates that closed form (J = 0 for the mean-equivalent variant) -- tests/test_oracle.py checks it against
the reference's own functions run live through the NumPy-backed CasADi shim, and tests/golden/uncertainty_propagation.npz
holds outputs of those reference functions.  PINNED (reference code, live + golden).
"""
import numpy as np


def onestep_batch(mu_x, gp, k_ff, sigma_x=None, k_fb=None, a=None, b=None, a_gp_inp_x=None, taylor=True):
    """mu_x (B,n_s), k_ff (B,n_u), sigma_x None | (B,n_s,n_s), k_fb (n_u,n_s) | (B,n_u,n_s)
    -> mu_new (B,n_s), sigma_new (B,n_s,n_s), var (B,n_s)"""
    bsz, n_s = mu_x.shape
    n_u = k_ff.shape[1]
    if a is None:
        a = np.eye(n_s)
        b = np.zeros((n_s, n_u))
    x_bar = mu_x if a_gp_inp_x is None else mu_x @ a_gp_inp_x.T
    mu_g, var, jac = gp.predict_batch(np.hstack((x_bar, k_ff)))
    mu_new = mu_x @ a.T + k_ff @ b.T + mu_g
    idx = np.arange(n_s)
    if sigma_x is None:
        sigma_new = np.zeros((bsz, n_s, n_s))
        sigma_new[:, idx, idx] = var
        return mu_new, sigma_new, var
    h = a[None] + b[None] @ k_fb
    if taylor:
        n_in = n_s if a_gp_inp_x is None else a_gp_inp_x.shape[0]
        jx = jac[:, :, :n_in] if a_gp_inp_x is None else jac[:, :, :n_in] @ a_gp_inp_x
        h = h + jx + jac[:, :, n_in:] @ k_fb
    sigma_new = h @ sigma_x @ np.swapaxes(h, 1, 2)
    sigma_new[:, idx, idx] += var
    return mu_new, sigma_new, var


def multistep_batch(mu_0, gp, k_ff, k_fb, a=None, b=None, a_gp_inp_x=None, taylor=True):
    """mu_0 (n_s,) | (B,n_s); k_ff (B,T,n_u); k_fb (T-1,n_u,n_s) | (B,T-1,n_u,n_s)
    -> mu_all (B,T,n_s), sigma_all (B,T,n_s,n_s), var_all (B,T,n_s)"""
    k_ff = np.asarray(k_ff, dtype=np.float64)
    bsz, hor, n_u = k_ff.shape
    k_fb = np.asarray(k_fb, dtype=np.float64)
    n_s = k_fb.shape[-1]
    mu = np.asarray(mu_0, dtype=np.float64)
    mu = np.broadcast_to(mu.reshape(-1, n_s), (bsz, n_s)).copy() if mu.size == n_s else mu.reshape(bsz, n_s)
    sigma = None
    mu_all = np.empty((bsz, hor, n_s))
    sigma_all = np.empty((bsz, hor, n_s, n_s))
    var_all = np.empty((bsz, hor, n_s))
    for t in range(hor):
        kfb_t = None if t == 0 else (k_fb[t - 1] if k_fb.ndim == 3 else k_fb[:, t - 1])
        mu, sigma, var = onestep_batch(mu, gp, k_ff[:, t], sigma, kfb_t, a, b, a_gp_inp_x, taylor)
        mu_all[:, t], sigma_all[:, t], var_all[:, t] = mu, sigma, var
    return mu_all, sigma_all, var_all
