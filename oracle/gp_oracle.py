"""Float64 NumPy restatement of the GP half of the hot path.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Never imported by the product.

What it restates (all citations relative to /root/reference/safe_exploration):

* kernels            ssm_gpy/gp_models_utils_casadi.py:17-40 (_k_rbf), :43-70 (_k_mat52),
                     :160-174 (_unscaled_dist: r^2 = -2 x y^T + |x|^2 + |y|^2)
* posterior state    ssm_gpy/gaussian_process.py:238-263: per output dimension d an
                     independent GP; ``inv_K[d] = (K_d + noise_d I)^-1`` (GPy
                     ``posterior.woodbury_inv``), ``beta[:, d] = inv_K[d] y_d``
                     (``woodbury_vector``).  ``noise_d`` is the TOTAL diagonal term the
                     caller wants (GPy: Gaussian_noise.variance + noise_diag(1e-5) + 1e-8
                     jitter of ExactGaussianInference); it is an explicit input here
                     because GPy is not installable in this image.
* prediction         ssm_gpy/gp_models_utils_casadi.py:177-197 (gp_pred):
                     mu = k* beta ; var = k(x,x) - sum((k* inv_K) * k*)   (explicit inverse)
                     and GPy's predict_noiseless form var = k(x,x) - |L^-1 k*|^2 (Cholesky),
                     used by SimpleGPModel.predict (ssm_gpy/gaussian_process.py:546-568).
* mean Jacobian      ssm_gpy/gp_models_utils_casadi.py:275-280 (CasADi AD of mu wrt input);
                     restated in closed form and verified against finite differences.
* call surface       SimpleGPModel.__call__ (ssm_gpy/gaussian_process.py:135-144,161-175):
                     (states 1 x n_s, actions 1 x n_u) -> (mu n_s x 1, var n_s x 1, jac n_s x D).

* composite kernels  ssm_gpy/gp_models_utils_casadi.py:73-157 (_k_lin_rbf, _k_lin_mat52, _k_lin) and the GPy kernel
                     objects of ssm_gpy/gaussian_process.py:469-474, as one general form (composite_vectors).

PARITY STATUS: the kernel functions and the predictive formulas (kernel rows, k(x,x), mean, variance) are PINNED by the
reference's own gp_models_utils_casadi.py functions, run numerically through the NumPy-backed CasADi shim
(oracle/ref_loader.load_gp_utils): live in tests/test_oracle.py and as golden vectors
(tests/golden/gp_pred_reference.npz, oracle/make_golden.golden_gp_pred).  What stays a restatement is the posterior
state GPy computes -- (K + noise I)^-1 and (K + noise I)^-1 y, textbook definitions evaluated by LAPACK -- and the
closed-form mean Jacobian (the reference gets it from CasADi AD), which is pinned by finite differences.
"""
import numpy as np
import scipy.linalg as sla

SQRT5 = np.sqrt(5.0)
KERN_RBF = 0
KERN_MAT52 = 1
KERN_LIN_RBF = 2
KERN_LIN_MAT52 = 3
_KERN_IDS = {"rbf": KERN_RBF, "mat52": KERN_MAT52, "lin_rbf": KERN_LIN_RBF, "lin_mat52": KERN_LIN_MAT52}
COMPOSITE = ("lin_rbf", "lin_mat52")


def unscaled_dist(x, y):
    """gp_models_utils_casadi.py:160-174.  (Clipped at 0 like GPy's stationary.py, the
    reference's CasADi copy would return NaN for a -1e-17.)"""
    x1sq = np.sum(x ** 2, axis=1)
    x2sq = np.sum(y ** 2, axis=1)
    r2 = -2.0 * x @ y.T + x1sq[:, None] + x2sq[None, :]
    return np.sqrt(np.maximum(r2, 0.0))


def k_rbf(x, y, variance, lengthscale):
    """gp_models_utils_casadi.py:17-40."""
    r = unscaled_dist(x / lengthscale[None, :], y / lengthscale[None, :])
    return variance * np.exp(-0.5 * r ** 2)


def k_mat52(x, y, variance, lengthscale):
    """gp_models_utils_casadi.py:43-70."""
    r = unscaled_dist(x / lengthscale[None, :], y / lengthscale[None, :])
    return variance * (1.0 + SQRT5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-SQRT5 * r)


def k_lin(x, y, variances):
    """gp_models_utils_casadi.py:132-157: sum_j v_j x_j y_j."""
    sv = np.sqrt(np.asarray(variances, dtype=np.float64)).reshape(1, -1)
    return (x * sv) @ (y * sv).T


def composite_vectors(kern_type, hyp, dim, semantics="casadi"):
    """The composite kernels of the reference, k = k_lin(prod) * k_stat + k_lin(all), as three D-vectors of one general
    form  k(x, y) = (sum_j a_j x_j y_j) s^2 phi(|s * (x - y)|) + sum_j v_j x_j y_j :

      inverse length-scales s_j (0 = the dimension does not enter the stationary factor), product-linear weights a_j,
      linear variances v_j, and the stationary variance s^2.

    semantics="casadi": what the hot path evaluates -- _k_lin_rbf / _k_lin_mat52, gp_models_utils_casadi.py:73-129:
        the product term sees input column 1 ONLY (x[:, 1]) with the scalar prod.* hyper-parameters;
    semantics="gpy":    the kernel object GPy is given -- Linear(D) * RBF(D) + Linear(D, ARD=True),
        ssm_gpy/gaussian_process.py:469-474: all columns, one shared variance / length-scale in the product term.
    The reference mixes the two (K from GPy, k* from CasADi); SURVEY.md section 8 f3 records it.  hyp uses the
    reference's keys (ssm_gpy/gaussian_process.py:523-538)."""
    stat = "rbf" if kern_type == "lin_rbf" else "mat52"
    ls = float(np.asarray(hyp["prod.{}.lengthscale".format(stat)]).reshape(-1)[0])
    var = float(np.asarray(hyp["prod.{}.variance".format(stat)]).reshape(-1)[0])
    plv = float(np.asarray(hyp["prod.linear.variances"]).reshape(-1)[0])
    lv = np.asarray(hyp["linear.variances"], dtype=np.float64).reshape(-1)
    if lv.size == 1:
        lv = np.full(dim, lv[0])
    if semantics == "casadi":
        s = np.zeros(dim)
        s[1] = 1.0 / ls
        a = np.zeros(dim)
        a[1] = plv
    elif semantics == "gpy":
        s = np.full(dim, 1.0 / ls)
        a = np.full(dim, plv)
    else:
        raise ValueError(semantics)
    return s, a, lv, var


def k_composite(kern_type, x, y, s, a, v, variance):
    """General form above; kern_type picks the stationary factor."""
    xs, ys = x * s[None, :], y * s[None, :]
    d = xs[:, None, :] - ys[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    if kern_type in ("lin_rbf", KERN_LIN_RBF):
        stat = variance * np.exp(-0.5 * r ** 2)
    else:
        stat = variance * (1.0 + SQRT5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-SQRT5 * r)
    return ((x * a[None, :]) @ y.T) * stat + (x * v[None, :]) @ y.T


def kernel(kern_type, x, y, variance, lengthscale):
    if kern_type in ("rbf", KERN_RBF):
        return k_rbf(x, y, variance, lengthscale)
    if kern_type in ("mat52", KERN_MAT52):
        return k_mat52(x, y, variance, lengthscale)
    raise ValueError("Unknown kernel {}".format(kern_type))


def _scaled_diff_sq(z, x, lengthscale):
    """Direct (cancellation-free) r^2 between rows of z and rows of x, scaled per input dim."""
    zs = z / lengthscale[None, :]
    xs = x / lengthscale[None, :]
    d = zs[:, None, :] - xs[None, :, :]
    return np.sum(d * d, axis=2)


def vectors_from_reference_hyp(kern_types, hyp, dim, semantics="casadi"):
    """Per-dimension arrays (lengthscale with inf where a dimension is inactive, variance, prod_linear, linear) from
    the reference's hyper-parameter dicts (ssm_gpy/gaussian_process.py:515-538)."""
    n_s = len(kern_types)
    ls = np.ones((n_s, dim))
    var = np.ones(n_s)
    pl = np.zeros((n_s, dim))
    lin = np.zeros((n_s, dim))
    for d, (k, h) in enumerate(zip(kern_types, hyp)):
        if k in COMPOSITE:
            s, a, v, va = composite_vectors(k, h, dim, semantics)
            with np.errstate(divide="ignore"):
                ls[d] = 1.0 / s
            var[d], pl[d], lin[d] = va, a, v
        else:
            l = np.asarray(h["lengthscale"], dtype=np.float64).reshape(-1)
            ls[d] = l if l.size == dim else np.full(dim, l[0])
            var[d] = float(np.asarray(h["variance"]).reshape(-1)[0])
    return ls, var, pl, lin


def golden_gp_case(g, name):
    """(GPOracle, kern_types, hyp dicts in the reference's layout) of one case of tests/golden/gp_pred_reference.npz
    (written by oracle/make_golden.golden_gp_pred)."""
    kerns = [str(k) for k in g[name + "/kern_types"]]
    pre = name + "/hyp"
    hyp = [{k[len(pre) + 2:]: g[k] for k in g.files if k.startswith("%s%d/" % (pre, i))} for i in range(len(kerns))]
    x = g[name + "/x_train"]
    ls, var, pl, lin = vectors_from_reference_hyp(kerns, hyp, x.shape[1])
    ora = GPOracle(x, g[name + "/y_train"], kerns, ls, var, g[name + "/noise"], prod_linear=pl, linear=lin)
    return ora, kerns, hyp


class GPOracle(object):
    """n_s independent exact GPs sharing the training inputs (SimpleGPModel posterior state).

    Parameters
    ----------
    x_train : (N, D) float64          training inputs z_i = [state, action]
    y_train : (N, n_s) float64        training targets
    kern_types : list[str] len n_s    "rbf" | "mat52"
    lengthscale : (n_s, D)            ARD length-scales per output dimension
    variance : (n_s,)                 signal variances sigma_f^2
    noise : (n_s,)                    TOTAL diagonal added to K (see module docstring)
    """

    def __init__(self, x_train, y_train, kern_types, lengthscale, variance, noise, prod_linear=None, linear=None):
        """prod_linear / linear : (n_s, D) or None -- rows a_d, v_d of the general composite form (composite_vectors)
        for the output dimensions with a "lin_*" kernel; there ``lengthscale`` may hold inf (inverse 0)."""
        self.prod_linear = None if prod_linear is None else np.asarray(prod_linear, dtype=np.float64)
        self.linear = None if linear is None else np.asarray(linear, dtype=np.float64)
        self.x_train = np.ascontiguousarray(x_train, dtype=np.float64)
        self.y_train = np.ascontiguousarray(y_train, dtype=np.float64)
        self.n_train, self.dim_in = self.x_train.shape
        self.n_s_out = self.y_train.shape[1]
        self.kern_types = list(kern_types)
        self.lengthscale = np.asarray(lengthscale, dtype=np.float64).reshape(self.n_s_out, self.dim_in)
        self.variance = np.asarray(variance, dtype=np.float64).reshape(self.n_s_out)
        self.noise = np.asarray(noise, dtype=np.float64).reshape(self.n_s_out)
        self.chol = []      # lower Cholesky factors L_d
        self.inv_K = []     # explicit inverses, as the reference stores them
        self.beta = np.empty((self.n_train, self.n_s_out))
        for d in range(self.n_s_out):
            k = self.kstar(d, self.x_train)
            k = 0.5 * (k + k.T)
            k[np.diag_indices_from(k)] = self.prior_var(d, self.x_train) + self.noise[d]
            l = np.linalg.cholesky(k)
            self.chol.append(l)
            self.beta[:, d] = sla.cho_solve((l, True), self.y_train[:, d])
        self._have_inv = False

    # -- explicit inverse exactly as the reference keeps it (dpotri of the Cholesky factor)
    def _ensure_inv(self):
        if not self._have_inv:
            eye = np.eye(self.n_train)
            self.inv_K = [sla.cho_solve((l, True), eye) for l in self.chol]
            self._have_inv = True

    def is_composite(self, d):
        return self.kern_types[d] in COMPOSITE

    def kstar(self, d, z):
        if self.is_composite(d):
            return k_composite(self.kern_types[d], z, self.x_train, 1.0 / self.lengthscale[d], self.prod_linear[d],
                               self.linear[d], self.variance[d])
        return kernel(self.kern_types[d], z, self.x_train, self.variance[d], self.lengthscale[d])

    def prior_var(self, d, z):
        """k_d(z, z) per row of z (the diag_only branch of the kernel functions)."""
        if self.is_composite(d):
            return self.variance[d] * np.sum(self.prod_linear[d][None, :] * z * z, axis=1) + \
                np.sum(self.linear[d][None, :] * z * z, axis=1)
        return np.full(z.shape[0], self.variance[d])

    def predict(self, z, form="chol"):
        """mean (T, n_s), variance (T, n_s) at inputs z (T, D).

        form="explicit": gp_models_utils_casadi.py:186-193 (what __call__ executes);
        form="chol":     GPy predict_noiseless (Cholesky + dtrtrs)."""
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        t = z.shape[0]
        mu = np.empty((t, self.n_s_out))
        var = np.empty((t, self.n_s_out))
        if form == "explicit":
            self._ensure_inv()
        for d in range(self.n_s_out):
            ks = self.kstar(d, z)                       # (T, N)
            mu[:, d] = ks @ self.beta[:, d]
            if form == "explicit":
                var[:, d] = self.prior_var(d, z) - np.sum((ks @ self.inv_K[d]) * ks, axis=1)
            else:
                v = sla.solve_triangular(self.chol[d], ks.T, lower=True, check_finite=False)
                var[:, d] = self.prior_var(d, z) - np.sum(v * v, axis=0)
        return mu, var

    def jacobian(self, z):
        """d mu_d / d z, shape (T, n_s, D).  Closed forms of the AD result at
        gp_models_utils_casadi.py:275-280:
          rbf   : J = -sum_i beta_i k_i (z - x_i) / l^2
          mat52 : J = -(5/3) s2 sum_i beta_i (1 + sqrt5 r_i) exp(-sqrt5 r_i) (z - x_i) / l^2
        """
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        t = z.shape[0]
        jac = np.empty((t, self.n_s_out, self.dim_in))
        for d in range(self.n_s_out):
            ls = self.lengthscale[d]
            if self.is_composite(d):
                jac[:, d, :] = self._jacobian_composite(d, z)
                continue
            r = unscaled_dist(z / ls[None, :], self.x_train / ls[None, :])   # (T, N), BLAS form
            if self.kern_types[d] in ("rbf", KERN_RBF):
                g = self.variance[d] * np.exp(-0.5 * r * r)
            else:
                g = (5.0 / 3.0) * self.variance[d] * (1.0 + SQRT5 * r) * np.exp(-SQRT5 * r)
            w = g * self.beta[None, :, d]                        # (T, N)
            # sum_i w_i (z_j - x_ij) = z_j sum_i w_i - sum_i w_i x_ij
            jac[:, d, :] = -(z * np.sum(w, axis=1, keepdims=True) - w @ self.x_train) / (ls ** 2)[None, :]
        return jac

    def _jacobian_composite(self, d, z):
        """k = lp * stat + ll with lp = sum_j a_j z_j x_j, ll = sum_j v_j z_j x_j:
           dk/dz_j = a_j x_j stat - lp g s_j^2 (z_j - x_j) + v_j x_j,
        g = stat (rbf) or (5/3) s2 (1 + sqrt5 r) exp(-sqrt5 r) (mat52)."""
        s, a, v, var = 1.0 / self.lengthscale[d], self.prod_linear[d], self.linear[d], self.variance[d]
        x, beta = self.x_train, self.beta[:, d]
        diff = z[:, None, :] - x[None, :, :]                                  # (T, N, D)
        r = np.sqrt(np.sum((diff * s[None, None, :]) ** 2, axis=2))
        if self.kern_types[d] == "lin_rbf":
            stat = var * np.exp(-0.5 * r * r)
            g = stat
        else:
            e = var * np.exp(-SQRT5 * r)
            stat = (1.0 + SQRT5 * r + 5.0 / 3.0 * r * r) * e
            g = (5.0 / 3.0) * (1.0 + SQRT5 * r) * e
        lp = (z * a[None, :]) @ x.T                                           # (T, N)
        t1 = ((stat * beta[None, :]) @ x) * a[None, :]
        t2 = -np.einsum("tn,tnj->tj", lp * g * beta[None, :], diff) * (s * s)[None, :]
        t3 = (beta @ x)[None, :] * v[None, :]
        return t1 + t2 + t3

    def jacobian_fd(self, z, eps=1e-6):
        """Central finite differences of the mean (identity check for `jacobian`)."""
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        jac = np.empty((z.shape[0], self.n_s_out, self.dim_in))
        for j in range(self.dim_in):
            dz = np.zeros(self.dim_in)
            dz[j] = eps
            mp, _ = self.predict(z + dz)
            mm, _ = self.predict(z - dz)
            jac[:, :, j] = (mp - mm) / (2 * eps)
        return jac

    # -- the exact call surface onestep_reachability uses (gp_reachability.py:74,101)
    def __call__(self, states, actions):
        """SimpleGPModel.__call__ (ssm_gpy/gaussian_process.py:135-144): single input only."""
        n, _ = np.shape(states)
        if n > 1:
            raise NotImplementedError(
                "Currently do not support multiple state-action pairs to evaluate on.")
        z = np.hstack((np.asarray(states, dtype=np.float64), np.asarray(actions, dtype=np.float64)))
        mu, var = self.predict(z, form="explicit")
        jac = self.jacobian(z)
        return mu.T, var.T, jac[0]

    # -- vectorised triple for the batch oracle
    def predict_batch(self, z):
        mu, var = self.predict(z, form="chol")
        return mu, var, self.jacobian(z)
