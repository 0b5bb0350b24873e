"""Gaussian uncertainty propagation with the reference's call signatures, evaluated in batch on the GPU.

Mirror of reference safe_exploration/uncertainty_propagation_casadi.py:

* ``one_step_taylor(mu_x, ssm, k_ff, sigma_x=None, k_fb=None, a=None, b=None, a_gp_inp_x=None)``          (:11-87)
* ``multi_step_taylor_symbolic(mu_0, ssm, k_ff, k_fb, sigma_0=None, a=None, b=None, a_gp_inp_x=None)``     (:90-148)
* ``one_step_mean_equivalent(...)``                                                                        (:210-283)
* ``mean_equivalent_multistep(...)``                                                                       (:151-207)

Same names, positional order and return tuples for the reference's un-batched shapes (``mu`` n_s x 1, ``k_ff`` T x n_u,
``k_fb`` (T-1) x n_u x n_s): ``(mu_all T x n_s, sigma_all T x (n_s n_s), gp_sigma_pred_all)``.  With a leading batch
axis (``k_ff`` (B,T,n_u), ``mu_0`` (n_s,) or (B,n_s)) B propagations run in one library call and the results carry
the batch axis: ``(mu_all (B,T,n_s), sigma_all (B,T,n_s,n_s), var_all (B,T,n_s))``.

The GP posterior is the same fused path as the ellipsoid rollouts (segp_multistep with ``propagation`` =
SEGP_PROP_TAYLOR / SEGP_PROP_MEAN_EQUIVALENT); only the n_s x n_s update differs:
``Sigma' = H Sigma H^T + diag(sigma_g^2)`` with ``H = A + J_x T + (B + J_u) K`` (Taylor) or ``H = A + B K``.

Third return value: the reference returns the predictive variances ``pred_var.T`` (1 x n_s) from the first step and from
every mean-equivalent step, but the full ``sigma_g`` matrix (``diag(var) + J Sigma_z J^T``, n_s x n_s) from a Taylor
step with an input covariance (:87) -- stacked as they come by the multi-step functions.  The un-batched functions here
reproduce exactly that; the batched form returns the predictive variances.
"""
import numpy as np

from . import _lib
from .gp_reachability import rollout
from .ssm import BatchedGPSSM

__all__ = ["one_step_taylor", "multi_step_taylor_symbolic", "multi_step_taylor", "one_step_mean_equivalent",
           "mean_equivalent_multistep"]


def _need_gp(ssm):
    if not isinstance(ssm, BatchedGPSSM):
        raise TypeError("the B200 uncertainty propagation needs a BatchedGPSSM")


def _propagate(mu_0, ssm, k_ff, k_fb, sigma_0, k_fb_0, a, b, a_gp_inp_x, mode):
    """k_ff (B,T,n_u) -> RolloutResult with covariances in q_all."""
    n_s = ssm.n_s_out
    zeros = np.zeros(n_s)
    return rollout(ssm, mu_0, k_ff, k_fb, zeros, zeros, sigma_0, k_fb_0, 1.0, a, b, a_gp_inp_x, True, mode)


def _sigma_g_taylor(ssm, mu_x, k_ff, sigma_x, k_fb, a_gp_inp_x):
    """sigma_g of one_step_taylor (:59-73) for one input: diag(var) + J Sigma_z J^T, J Sigma_z J^T = G Sigma G^T with
    G = J_x T + J_u K."""
    n_s = ssm.n_s_out
    x_bar = mu_x.reshape(1, -1) if a_gp_inp_x is None else mu_x.reshape(1, -1) @ np.asarray(a_gp_inp_x).T
    _, var, jac = ssm.predict(x_bar, k_ff.reshape(1, -1), jacobians=True)[:3]
    jac = np.asarray(jac).reshape(n_s, -1)
    n_in = x_bar.shape[1]
    jx = jac[:, :n_in] if a_gp_inp_x is None else jac[:, :n_in] @ np.asarray(a_gp_inp_x)
    g = jx + jac[:, n_in:] @ k_fb
    return np.diag(np.asarray(var).reshape(-1)) + g @ sigma_x @ g.T


def _one_step(mu_x, ssm, k_ff, sigma_x, k_fb, a, b, a_gp_inp_x, mode):
    _need_gp(ssm)
    n_s, n_u = ssm.n_s_out, ssm.n_u
    mu_arr = np.asarray(mu_x, dtype=np.float64)
    k_ff_arr = np.asarray(k_ff, dtype=np.float64)
    batched = k_ff_arr.ndim == 2 and k_ff_arr.shape[-1] == n_u and not (k_ff_arr.shape == (n_u, 1) and mu_arr.shape == (n_s, 1))
    if batched:
        bsz = k_ff_arr.shape[0]
        res = _propagate(mu_arr.reshape(-1, n_s), ssm, k_ff_arr.reshape(bsz, 1, n_u), None, sigma_x, k_fb, a, b,
                         a_gp_inp_x, mode)
        return res.p_all[:, 0], res.q_all[:, 0], res.var_all[:, 0]
    if (sigma_x is None) != (k_fb is None):
        raise ValueError("sigma_x and k_fb must be given together")
    res = _propagate(mu_arr.reshape(n_s), ssm, k_ff_arr.reshape(1, 1, n_u), None,
                     None if sigma_x is None else np.asarray(sigma_x, dtype=np.float64).reshape(n_s, n_s),
                     None if k_fb is None else np.asarray(k_fb, dtype=np.float64).reshape(n_u, n_s), a, b, a_gp_inp_x,
                     mode)
    mu_new = res.p_all[0, 0].reshape(n_s, 1)
    sigma_new = res.q_all[0, 0]
    if sigma_x is None or mode == _lib.PROP_MEAN_EQUIVALENT:
        third = res.var_all[0, 0].reshape(1, n_s)                     # pred_var.T / sigma_g.T of the reference
    else:
        third = _sigma_g_taylor(ssm, mu_arr.reshape(n_s), k_ff_arr.reshape(n_u), np.asarray(sigma_x).reshape(n_s, n_s),
                                np.asarray(k_fb).reshape(n_u, n_s), a_gp_inp_x).T
    return mu_new, sigma_new, third


def one_step_taylor(mu_x, ssm, k_ff, sigma_x=None, k_fb=None, a=None, b=None, a_gp_inp_x=None):
    """uncertainty_propagation_casadi.py:11-87"""
    return _one_step(mu_x, ssm, k_ff, sigma_x, k_fb, a, b, a_gp_inp_x, _lib.PROP_TAYLOR)


def one_step_mean_equivalent(mu_x, ssm, k_ff, sigma_x=None, k_fb=None, a=None, b=None, a_gp_inp_x=None):
    """uncertainty_propagation_casadi.py:210-283"""
    return _one_step(mu_x, ssm, k_ff, sigma_x, k_fb, a, b, a_gp_inp_x, _lib.PROP_MEAN_EQUIVALENT)


def _multi_step(mu_0, ssm, k_ff, k_fb, sigma_0, a, b, a_gp_inp_x, mode):
    _need_gp(ssm)
    if sigma_0 is not None:
        raise NotImplementedError("Still need  to do this")     # as the reference (:122-123, 180-181)
    n_s, n_u = ssm.n_s_out, ssm.n_u
    k_ff_arr = np.asarray(k_ff, dtype=np.float64)
    if k_ff_arr.ndim == 3:
        res = _propagate(mu_0, ssm, k_ff_arr, k_fb, None, None, a, b, a_gp_inp_x, mode)
        return res.p_all, res.q_all, res.var_all
    hor = k_ff_arr.shape[0]
    k_fb_arr = None if hor == 1 else np.asarray(k_fb, dtype=np.float64).reshape(hor - 1, n_u, n_s)
    mu0 = np.asarray(mu_0, dtype=np.float64).reshape(n_s)
    res = _propagate(mu0, ssm, k_ff_arr.reshape(1, hor, n_u), k_fb_arr, None, None, a, b, a_gp_inp_x, mode)
    mu_all = res.p_all[0]
    sigma = res.q_all[0]
    third = [res.var_all[0, 0].reshape(1, n_s)]
    for i in range(hor - 1):
        if mode == _lib.PROP_MEAN_EQUIVALENT:
            third.append(res.var_all[0, i + 1].reshape(1, n_s))
        else:
            third.append(_sigma_g_taylor(ssm, mu_all[i], k_ff_arr[i + 1], sigma[i], k_fb_arr[i], a_gp_inp_x).T)
    return mu_all, sigma.reshape(hor, n_s * n_s), np.vstack(third)


def multi_step_taylor_symbolic(mu_0, ssm, k_ff, k_fb, sigma_0=None, a=None, b=None, a_gp_inp_x=None):
    """uncertainty_propagation_casadi.py:90-148"""
    return _multi_step(mu_0, ssm, k_ff, k_fb, sigma_0, a, b, a_gp_inp_x, _lib.PROP_TAYLOR)


multi_step_taylor = multi_step_taylor_symbolic


def mean_equivalent_multistep(mu_0, ssm, k_ff, k_fb, sigma_0=None, a=None, b=None, a_gp_inp_x=None):
    """uncertainty_propagation_casadi.py:151-207"""
    return _multi_step(mu_0, ssm, k_ff, k_fb, sigma_0, a, b, a_gp_inp_x, _lib.PROP_MEAN_EQUIVALENT)
