"""Float64 NumPy restatement of greedy maximum-predicted-variance data selection.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

What it restates: the selection loop of SimpleGPModel.choose_datapoints_maxvar
(/root/reference/safe_exploration/ssm_gpy/gaussian_process.py:280-345) --

    repeat:  _, pred_var_pool = self.predict(x_pool);  idx = argmax(sum(pred_var_pool, axis=1));
             move x_pool[idx] to the chosen set;  self.gps[j].set_XY(x_chosen, y_chosen[:, j])

i.e. at every step the pool point with the largest predictive variance (summed over the output dimensions) of
the GPs conditioned on the points chosen so far.  With fixed hyper-parameters the predictive variance does not
depend on the targets, so the selection is a function of the inputs alone.

Two things of the reference are NOT restated, on purpose (they make its result non-deterministic and tie it to GPy's
optimiser): the k-means + np.random.choice initial set (:303-318) and the hyper-parameter re-optimisation on the
pool every `chunks` iterations (:327-330).  The selection here starts from the empty set with fixed hyper-parameters.

greedy_maxvar       incremental form (what the CUDA kernels do): per output dimension a partial Cholesky factor of
                    K + noise I restricted to the chosen points, one new column per step, O(n t) per step
brute_force_scores  the defining form, for pinning: GP posterior variance of every pool point given the chosen set,
                    through a fresh Cholesky solve (predict_noiseless)
"""
import numpy as np
import scipy.linalg as sla

from . import gp_oracle


def _kernel_fn(kern_types, lengthscale, variance, prod_linear, linear, d):
    if kern_types[d] in gp_oracle.COMPOSITE:
        return lambda a, b: gp_oracle.k_composite(kern_types[d], a, b, 1.0 / lengthscale[d], prod_linear[d], linear[d],
                                                  variance[d])
    return lambda a, b: gp_oracle.kernel(kern_types[d], a, b, variance[d], lengthscale[d])


def _prior(kern_types, variance, prod_linear, linear, d, x):
    if kern_types[d] in gp_oracle.COMPOSITE:
        return variance[d] * np.sum(prod_linear[d][None, :] * x * x, axis=1) + np.sum(linear[d][None, :] * x * x, axis=1)
    return np.full(x.shape[0], variance[d])


def greedy_maxvar(x, kern_types, lengthscale, variance, noise, m, prod_linear=None, linear=None):
    """Indices (m,) of the selected rows of x in selection order, and the summed predictive variance (m,) each had
    when it was selected.  Ties go to the lowest index."""
    x = np.asarray(x, dtype=np.float64)
    n, n_s = x.shape[0], len(kern_types)
    m = min(int(m), n)
    lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(n_s, -1)
    v = np.stack([_prior(kern_types, variance, prod_linear, linear, d, x) for d in range(n_s)])   # (n_s, n)
    cols = np.zeros((n_s, m, n))
    chosen = np.zeros(n, dtype=bool)
    idx = np.empty(m, dtype=np.int64)
    score = np.empty(m)
    kfn = [_kernel_fn(kern_types, lengthscale, variance, prod_linear, linear, d) for d in range(n_s)]
    for t in range(m):
        s = np.where(chosen, -np.inf, v.sum(axis=0))
        j = int(np.argmax(s))
        idx[t], score[t] = j, s[j]
        chosen[j] = True
        for d in range(n_s):
            c = kfn[d](x, x[j:j + 1])[:, 0]
            c[j] = _prior(kern_types, variance, prod_linear, linear, d, x[j:j + 1])[0]
            c = c - cols[d, :t].T @ cols[d, :t, j]
            l = c / np.sqrt(v[d, j] + noise[d])
            cols[d, t] = l
            v[d] = v[d] - l * l
    return idx, score


def brute_force_scores(x, kern_types, lengthscale, variance, noise, chosen_idx, prod_linear=None, linear=None):
    """sum_d var_d(x_i | observations at chosen_idx) for every row of x, by a direct Cholesky solve."""
    x = np.asarray(x, dtype=np.float64)
    n_s = len(kern_types)
    lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(n_s, -1)
    out = np.zeros(x.shape[0])
    xs = x[np.asarray(chosen_idx, dtype=np.int64)]
    for d in range(n_s):
        kfn = _kernel_fn(kern_types, lengthscale, variance, prod_linear, linear, d)
        prior = _prior(kern_types, variance, prod_linear, linear, d, x)
        if xs.shape[0] == 0:
            out += prior
            continue
        kss = kfn(xs, xs)
        kss = 0.5 * (kss + kss.T)
        kss[np.diag_indices_from(kss)] = _prior(kern_types, variance, prod_linear, linear, d, xs) + noise[d]
        l = np.linalg.cholesky(kss)
        vv = sla.solve_triangular(l, kfn(xs, x), lower=True)
        out += prior - np.sum(vv * vv, axis=0)
    return out
