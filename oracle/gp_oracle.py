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

PARITY STATUS: GPy / CasADi absent => this half is "parity unpinned by literals"; it is
pinned by mathematical identities in tests/test_oracle.py.
"""
import numpy as np
import scipy.linalg as sla

SQRT5 = np.sqrt(5.0)
KERN_RBF = 0
KERN_MAT52 = 1
_KERN_IDS = {"rbf": KERN_RBF, "mat52": KERN_MAT52}


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

    def __init__(self, x_train, y_train, kern_types, lengthscale, variance, noise):
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
            k = kernel(self.kern_types[d], self.x_train, self.x_train, self.variance[d],
                       self.lengthscale[d])
            k = 0.5 * (k + k.T)
            k[np.diag_indices_from(k)] = self.variance[d] + self.noise[d]
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

    def kstar(self, d, z):
        return kernel(self.kern_types[d], z, self.x_train, self.variance[d], self.lengthscale[d])

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
                var[:, d] = self.variance[d] - np.sum((ks @ self.inv_K[d]) * ks, axis=1)
            else:
                v = sla.solve_triangular(self.chol[d], ks.T, lower=True, check_finite=False)
                var[:, d] = self.variance[d] - np.sum(v * v, axis=0)
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
            r = unscaled_dist(z / ls[None, :], self.x_train / ls[None, :])   # (T, N), BLAS form
            if self.kern_types[d] in ("rbf", KERN_RBF):
                g = self.variance[d] * np.exp(-0.5 * r * r)
            else:
                g = (5.0 / 3.0) * self.variance[d] * (1.0 + SQRT5 * r) * np.exp(-SQRT5 * r)
            w = g * self.beta[None, :, d]                        # (T, N)
            # sum_i w_i (z_j - x_ij) = z_j sum_i w_i - sum_i w_i x_ij
            jac[:, d, :] = -(z * np.sum(w, axis=1, keepdims=True) - w @ self.x_train) / (ls ** 2)[None, :]
        return jac

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
