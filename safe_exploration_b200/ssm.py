"""BatchedGPSSM: the state-space-model plugin of the reference, backed by libsegp.so on a B200.

Mirrors, for the hot path only, the surface of

* ``StateSpaceModel`` (reference safe_exploration/state_space_models.py:14-211):
  ``num_states``, ``num_actions``, ``predict(states, actions, jacobians, full_cov)``, ``__call__``,
  ``linearize_predict``, ``update_model``;
* ``SimpleGPModel`` (reference safe_exploration/ssm_gpy/gaussian_process.py:15-634):
  ``n_s_out / n_s_in / n_u``, ``train``, ``update_model``, GPy-style ``predict(x_new,
  compute_gradients=...)``, ``predictive_gradients``, ``to_dict / from_dict``, ``information_gain``,
  ``choose_datapoints_maxvar``, the subset-of-data mode (``m``), ``sample_from_gp``,
  ``x_train / y_train / z / beta / hyp / kern_types / gp_trained``.

What is different, on purpose:

* every entry point accepts a batch (the reference's ``__call__`` raises for more than one input,
  ssm_gpy/gaussian_process.py:142-143);
* hyper-parameters are explicit inputs (``hyp``), never optimised here: hyper-parameter optimisation is
  GPy's L-BFGS and is out of scope; ``opt_hyp=True`` raises NotImplementedError;
* the model lives on the GPU; there is no CPU path: constructing a model without a CUDA device raises.

The arithmetic is float64 end to end (the reference is float64; the predictive variance cancels 3-4
digits).
"""
import ctypes
import os
import warnings

import numpy as np

from . import _lib

__all__ = ["BatchedGPSSM"]

# Which tensor pipe runs the variance contraction of new models (include/segp.h, option "tri_mode"):
# -1 automatic (int8 digit planes on tcgen05 when the padded training size allows and the factorize-time probe finds
# them precise enough, else fp64 DMMA), 0 fp64 DMMA, 1 the int8 reference kernel (test cross-check), 4 / 5 the
# production tcgen05 kernels (one cluster per tile / persistent).  "i8_digits": 0 automatic, 5 = 15 int8 products,
# 4 = 10 products (diagonal-split digit set), guarded by the a-posteriori error estimate.
DEFAULT_TRI_MODE = int(os.environ.get("SEGP_TRI_MODE", "-1"))
DEFAULT_I8_DIGITS = int(os.environ.get("SEGP_I8_DIGITS", "0"))

_GPY_JITTER = 1e-8        # ExactGaussianInference adds 1e-8 to the diagonal of K
_GPY_DEFAULT_NOISE = 1.0  # GPRegression default Gaussian_noise.variance


def _default_hyp(n_s_out, dim, kern_types=None):
    """GPy defaults (lengthscale 1, variance 1, Gaussian noise 1), i.e. what SimpleGPModel yields for
    train(..., opt_hyp=False) (reference test/test_safempc.py:56-69)."""
    out = []
    for d in range(n_s_out):
        k = kern_types[d] if kern_types is not None else "rbf"
        if k in _lib.COMPOSITE_KERNELS:
            st = "rbf" if k == "lin_rbf" else "mat52"
            out.append({"prod.%s.lengthscale" % st: np.ones(1), "prod.%s.variance" % st: 1.0,
                        "prod.linear.variances": np.ones(1), "linear.variances": np.ones(dim),
                        "noise": _GPY_DEFAULT_NOISE})
        else:
            out.append({"lengthscale": np.ones(dim), "variance": 1.0, "noise": _GPY_DEFAULT_NOISE})
    return out


def _composite_vectors(kern_type, hyp, dim, semantics):
    """Inverse length-scales s_j, product-linear weights a_j, linear variances v_j and the stationary variance of
    the general form  k(x,y) = (sum_j a_j x_j y_j) s_f^2 phi(|s (x - y)|) + sum_j v_j x_j y_j  for the reference's
    composite kernels.  semantics "casadi": the product term sees input column 1 only, what the hot path evaluates
    (_k_lin_rbf / _k_lin_mat52, ssm_gpy/gp_models_utils_casadi.py:73-129); "gpy": the kernel object GPy is given,
    Linear(D) * RBF(D) + Linear(D, ARD) on all columns (ssm_gpy/gaussian_process.py:469-474)."""
    st = "rbf" if kern_type == "lin_rbf" else "mat52"
    ls = float(np.asarray(hyp["prod.%s.lengthscale" % st]).reshape(-1)[0])
    var = float(np.asarray(hyp["prod.%s.variance" % st]).reshape(-1)[0])
    plv = float(np.asarray(hyp["prod.linear.variances"]).reshape(-1)[0])
    lv = np.asarray(hyp["linear.variances"], dtype=np.float64).reshape(-1)
    if lv.size == 1:
        lv = np.full(dim, lv[0])
    if lv.size != dim:
        raise ValueError("linear.variances needs {} entries".format(dim))
    if semantics == "casadi":
        if dim < 2:
            raise ValueError("the CasADi form of the composite kernels uses input column 1")
        s = np.zeros(dim)
        s[1] = 1.0 / ls
        a = np.zeros(dim)
        a[1] = plv
    elif semantics == "gpy":
        s = np.full(dim, 1.0 / ls)
        a = np.full(dim, plv)
    else:
        raise ValueError("composite_semantics must be 'casadi' or 'gpy'")
    return s, a, lv, var


class BatchedGPSSM(object):
    """n_s_out independent exact GPs over inputs [state (n_s_in), action (n_u)], evaluated in batch on the GPU.

    Parameters
    ----------
    n_s_out, n_s_in, n_u : int
        As SimpleGPModel.__init__ (ssm_gpy/gaussian_process.py:32-33).
    X : (N, n_s_in+n_u), y : (N, n_s_out), optional
        Training data; if both are given and ``train`` is true the model is factorised immediately.
    kern_types : list[str], optional
        "rbf" | "mat52" | "lin_rbf" | "lin_mat52" per output dimension (default "rbf").  The composite kernels
         run on the same tensor pipes; on the int8 path their unbounded, signed values are
        scaled per trajectory by an analytic bound before the digit split.
    hyp : list[dict], optional
        Per output dimension ``{"lengthscale": (D,) or scalar, "variance": float, "noise": float}``; for the composite
        kernels ``{"prod.rbf.lengthscale" | "prod.mat52.lengthscale", "prod.*.variance", "prod.linear.variances",
        "linear.variances" (D,), "noise"}`` (the reference's hyp dicts, ssm_gpy/gaussian_process.py:494-538, plus the
        Gaussian noise variance GPy keeps on the likelihood).  Defaults are GPy's defaults.
    composite_semantics : "casadi" | "gpy"
        Which of the reference's two (inconsistent) definitions of the composite kernels to use for BOTH the
        factorised matrix and the kernel rows: the CasADi functions the hot path evaluates (product term on input
        column 1 only; default) or the GPy kernel object (all columns).
    device : int or torch.device, optional
        CUDA device (default: current).
    """

    has_jacobian = True
    has_reverse = False

    def __init__(self, n_s_out, n_s_in, n_u, X=None, y=None, m=None, kern_types=None, hyp=None, train=True,
                 Z=None, device=None, noise_diag=1e-5, tri_mode=None, composite_semantics="casadi", i8_digits=None,
                 guard_rtol=None):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        if Z is not None:
            raise NotImplementedError("inducing inputs Z (sparse GP regression) are not on this path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.n_s_out = int(n_s_out)
        self.n_s_in = int(n_s_in)
        self.n_u = int(n_u)
        self.num_states = self.n_s_out
        self.num_actions = self.n_u
        self.dim_in = self.n_s_in + self.n_u
        self.kern_types = list(kern_types) if kern_types is not None else ["rbf"] * self.n_s_out
        if len(self.kern_types) != self.n_s_out:
            raise ValueError("kern_types needs one entry per output dimension")
        for k in self.kern_types:
            if k not in _lib.KERN_IDS:
                raise ValueError("kernel type '{}' not supported".format(k))     # as gaussian_process.py:476-478
        self.has_composite = any(k in _lib.COMPOSITE_KERNELS for k in self.kern_types)
        if composite_semantics not in ("casadi", "gpy"):
            raise ValueError("composite_semantics must be 'casadi' or 'gpy'")
        self.composite_semantics = composite_semantics
        self.hyp = self._normalise_hyp(hyp)
        self.noise_diag = float(noise_diag)
        self.gp_trained = False
        self.x_train = None
        self.y_train = None
        self.z = None
        self.m = None if m is None else int(m)    # subset-of-data size (ssm_gpy/gaussian_process.py:36, 201-222)
        self._handle = ctypes.c_void_p()
        kern_ids = (ctypes.c_int * self.n_s_out)(*[_lib.KERN_IDS[k] for k in self.kern_types])
        _lib.check(self._lib.segp_create(ctypes.byref(self._handle), self.device.index, self.n_s_out, self.n_s_in,
                                         self.n_u, kern_ids))
        if tri_mode is None:
            tri_mode = DEFAULT_TRI_MODE
        self.set_option("tri_mode", tri_mode)
        self.set_option("i8_digits", DEFAULT_I8_DIGITS if i8_digits is None else i8_digits)
        if guard_rtol is not None:
            self.set_param("guard_rtol", guard_rtol)
        self.last_predict_status = None
        if X is not None and y is not None and train:
            self.train(X, y)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            self._lib.segp_destroy(h)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ hyper-parameters
    def _normalise_hyp(self, hyp):
        if hyp is None:
            return _default_hyp(self.n_s_out, self.dim_in, self.kern_types)
        if len(hyp) != self.n_s_out:
            raise ValueError("hyp needs one dict per output dimension")
        out = []
        for k, h in zip(self.kern_types, hyp):
            if k in _lib.COMPOSITE_KERNELS:
                hc = {key: np.asarray(val, dtype=np.float64).copy() for key, val in h.items() if key != "noise"}
                _composite_vectors(k, hc, self.dim_in, self.composite_semantics)      # validates keys and sizes
                hc["noise"] = float(np.asarray(h.get("noise", _GPY_DEFAULT_NOISE)).reshape(-1)[0])
                out.append(hc)
                continue
            ls = np.asarray(h.get("lengthscale", 1.0), dtype=np.float64).reshape(-1)
            if ls.size == 1:
                ls = np.full(self.dim_in, float(ls[0]))
            if ls.size != self.dim_in:
                raise ValueError("lengthscale needs {} entries".format(self.dim_in))
            out.append({"lengthscale": ls.copy(), "variance": float(np.asarray(h.get("variance", 1.0)).reshape(-1)[0]),
                        "noise": float(np.asarray(h.get("noise", _GPY_DEFAULT_NOISE)).reshape(-1)[0])})
        return out

    def _model_arrays(self):
        """(lengthscale (n_s, D) with inf where a composite kernel ignores a dimension, variance (n_s,),
        prod_linear (n_s, D), linear (n_s, D)) -- the arguments of segp_set_model / segp_set_linear_terms."""
        ls = np.ones((self.n_s_out, self.dim_in))
        var = np.ones(self.n_s_out)
        pl = np.zeros((self.n_s_out, self.dim_in))
        lin = np.zeros((self.n_s_out, self.dim_in))
        for d, (k, h) in enumerate(zip(self.kern_types, self.hyp)):
            if k in _lib.COMPOSITE_KERNELS:
                s, a, v, va = _composite_vectors(k, h, self.dim_in, self.composite_semantics)
                with np.errstate(divide="ignore"):
                    ls[d] = 1.0 / s
                var[d], pl[d], lin[d] = va, a, v
            else:
                ls[d], var[d] = h["lengthscale"], h["variance"]
        return ls, var, pl, lin

    def _upload(self, x_h, y_h):
        ls, var, pl, lin = (_lib.host_f64(v) for v in self._model_arrays())
        noise = _lib.host_f64(self.total_noise())
        _lib.check(self._lib.segp_set_model(self._handle, x_h.shape[0], _lib.dbl_ptr(x_h), _lib.dbl_ptr(y_h),
                                            _lib.dbl_ptr(ls), _lib.dbl_ptr(var), _lib.dbl_ptr(noise)))
        if self.has_composite:
            _lib.check(self._lib.segp_set_linear_terms(self._handle, _lib.dbl_ptr(pl), _lib.dbl_ptr(lin)))

    def total_noise(self):
        """Diagonal term added to K per output dimension: Gaussian noise + noise_diag
        (ssm_gpy/gaussian_process.py:252-253) + GPy's 1e-8 jitter."""
        return np.array([h["noise"] + self.noise_diag + _GPY_JITTER for h in self.hyp])

    # ------------------------------------------------------------------ training (posterior only)
    def train(self, X, y, m=None, opt_hyp=False, noise_diag=None, Z=None, choose_data=True, hyp=None):
        """Posterior half of SimpleGPModel.train (ssm_gpy/gaussian_process.py:238-263): per output dimension
        K_d + noise I -> Cholesky -> beta, L^-1, all on the GPU.  Hyper-parameters stay fixed."""
        if opt_hyp:
            raise NotImplementedError("hyper-parameter optimisation is out of scope for the B200 path; "
                                      "pass fixed hyper-parameters via hyp=")
        if Z is not None:
            raise NotImplementedError("inducing inputs Z (sparse GP regression) are not on this path")
        if hyp is not None:
            self.hyp = self._normalise_hyp(hyp)
        if noise_diag is not None:
            self.noise_diag = float(noise_diag)
        x_h = _lib.host_f64(X)
        y_h = _lib.host_f64(y)
        if x_h.ndim != 2 or x_h.shape[1] != self.dim_in:
            raise ValueError("X must be N x {}".format(self.dim_in))
        if y_h.ndim != 2 or y_h.shape != (x_h.shape[0], self.n_s_out):
            raise ValueError("y must be N x {}".format(self.n_s_out))
        # subset of data (ssm_gpy/gaussian_process.py:201-222): the GP conditions on m of the N points -- the
        # max-variance selection on the device, or a random draw -- and keeps the whole set as x_train / y_train
        m = self.m if m is None else int(m)
        x_z, y_z = x_h, y_h
        if m is not None:
            if x_h.shape[0] < m:
                warnings.warn("The desired number of datapoints is not available. Dataset consist of {} "
                              "Datapoints! ".format(x_h.shape[0]))
            elif x_h.shape[0] > m:
                if choose_data:
                    idx, _ = self.select_maxvar(x_h, m)
                else:
                    idx = np.random.choice(x_h.shape[0], size=m, replace=False)
                x_z, y_z = np.ascontiguousarray(x_h[idx]), np.ascontiguousarray(y_h[idx])
        self.gp_trained = False
        self._upload(x_z, y_z)
        self.x_train = x_h
        self.y_train = y_h
        self.z = x_z
        self._factorize()

    def _factorize(self):
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.segp_factorize(self._handle, _lib.current_stream(self.device)))
        self.gp_trained = True

    def update_model(self, x, y, opt_hyp=False, replace_old=True, noise_diag=None, choose_data=True):
        """SimpleGPModel.update_model (ssm_gpy/gaussian_process.py:347-419) without re-optimisation:
        replace or append the data and refactorise."""
        if opt_hyp:
            raise NotImplementedError("hyper-parameter optimisation is out of scope for the B200 path")
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if not replace_old and self.x_train is not None:
            if self.gp_trained and self.m is None and (noise_diag is None or float(noise_diag) == self.noise_diag):
                return self.append_data(x, y)
            x = np.vstack((self.x_train, x))
            y = np.vstack((self.y_train, y))
        self.train(x, y, noise_diag=noise_diag, choose_data=choose_data)

    def append_data(self, x, y):
        """Append training points to the factorised model (segp_append): while the padded size does not grow only the
        trailing rows of L^-1 are recomputed, O(n_new N^2) instead of O(N^3); ``get_option("append_incremental")``
        tells which path ran.  The first call keeps W dense on the device from then on."""
        x_h = _lib.host_f64(x).reshape(-1, self.dim_in)
        y_h = _lib.host_f64(y).reshape(-1, self.n_s_out)
        if x_h.shape[0] != y_h.shape[0]:
            raise ValueError("x and y need the same number of rows")
        if self.m is not None:      # subset-of-data models re-select from the whole set
            return self.train(np.vstack((self.x_train, x_h)), np.vstack((self.y_train, y_h)))
        with self._torch.cuda.device(self.device):
            try:
                _lib.check(self._lib.segp_append(self._handle, x_h.shape[0], _lib.dbl_ptr(x_h), _lib.dbl_ptr(y_h),
                                                 _lib.current_stream(self.device)))
            except Exception:
                # the library restores the previous model when an update fails; if even that failed the handle is
                # unfactorised and says so
                self.gp_trained = bool(self.get_option("factorized"))
                raise
        self.x_train = np.vstack((self.x_train, x_h))
        self.y_train = np.vstack((self.y_train, y_h))
        self.z = self.x_train

    # ------------------------------------------------------------------ multi-GPU: one broadcast of the factor
    def factor_buffers(self):
        """The device buffers that make up the factorised state, as uint8 torch views (for broadcast)."""
        _lib.check(self._lib.segp_alloc_factor_buffers(self._handle))
        out = []
        for i in range(self._lib.segp_num_factor_buffers(self._handle)):
            ptr = ctypes.c_void_p()
            nbytes = ctypes.c_size_t()
            _lib.check(self._lib.segp_factor_buffer(self._handle, i, ctypes.byref(ptr), ctypes.byref(nbytes)))
            out.append((ptr.value, nbytes.value))
        return out

    def factor_views(self):
        """``factor_buffers`` as uint8 CUDA tensors aliasing the device memory (what a broadcast writes into)."""
        return [_tensor_from_ptr(self._torch, ptr, nbytes, self.device) for ptr, nbytes in self.factor_buffers()]

    def set_data_only(self, X, y):
        """Upload data + hyper-parameters without factorising (non-root ranks before the broadcast)."""
        x_h = _lib.host_f64(X)
        y_h = _lib.host_f64(y)
        self._upload(x_h, y_h)
        self.x_train, self.y_train, self.z = x_h, y_h, x_h
        self.gp_trained = False

    def alloc_fp64_operand(self):
        """Allocate the float64 DMMA operand (factor buffer 1) on a rank that receives it by broadcast."""
        _lib.check(self._lib.segp_alloc_fp64_operand(self._handle))

    def mark_factorized(self):
        _lib.check(self._lib.segp_mark_factorized(self._handle))
        self.gp_trained = True

    # ------------------------------------------------------------------ prediction
    def _as_device_inputs(self, states, actions):
        torch = self._torch
        if actions is None:
            z = states
        else:
            if torch.is_tensor(states) or torch.is_tensor(actions):
                states = torch.as_tensor(states, dtype=torch.float64, device=self.device)
                actions = torch.as_tensor(actions, dtype=torch.float64, device=self.device)
                z = torch.cat((states.reshape(-1, self.n_s_in), actions.reshape(-1, self.n_u)), dim=1)
            else:
                z = np.hstack((np.asarray(states, dtype=np.float64).reshape(-1, self.n_s_in),
                               np.asarray(actions, dtype=np.float64).reshape(-1, self.n_u)))
        was_tensor = torch.is_tensor(z)
        z_d = torch.as_tensor(z, dtype=torch.float64, device=self.device).reshape(-1, self.dim_in).contiguous()
        return z_d, was_tensor

    def predict_device(self, z, jacobians=False):
        """Device-resident predict: z (T, D) float64 CUDA tensor -> (mu (T,n_s), var (T,n_s)[, jac (T,n_s,D)])."""
        torch = self._torch
        if not self.gp_trained:
            raise RuntimeError("model is not trained")
        assert z.is_cuda and z.dtype == torch.float64 and z.dim() == 2 and z.shape[1] == self.dim_in
        z = z.contiguous()
        t = z.shape[0]
        mu = torch.empty((t, self.n_s_out), dtype=torch.float64, device=self.device)
        var = torch.empty((t, self.n_s_out), dtype=torch.float64, device=self.device)
        jac = torch.empty((t, self.n_s_out, self.dim_in), dtype=torch.float64, device=self.device) if jacobians else None
        status = torch.empty((t,), dtype=torch.int32, device=self.device)
        _lib.check(self._lib.segp_predict_ex(self._handle, t, _lib.dev_ptr(z), _lib.dev_ptr(mu), _lib.dev_ptr(var),
                                             _lib.dev_ptr(jac), _lib.dev_ptr(status), _lib.current_stream(self.device)))
        # per-input bits (_lib.STATUS_BAD_VARIANCE, _lib.STATUS_LOW_PRECISION) of the last call, a CUDA int32 tensor
        self.last_predict_status = status
        return (mu, var, jac) if jacobians else (mu, var)

    def predict(self, states, actions=None, jacobians=False, full_cov=False, quantiles=None,
                compute_gradients=False):
        """Both reference spellings:

        * ``predict(states (T,n_s), actions (T,n_u), jacobians=False, full_cov=False)``
          (StateSpaceModel.predict, state_space_models.py:74-104) -> ``(mean (T,n_s), var (T,n_s)[, jac_mean])``;
        * ``predict(x_new (T,D), compute_gradients=False)`` (SimpleGPModel.predict,
          ssm_gpy/gaussian_process.py:546-568) -> ``(mu, var[, grad_mu (T,n_s,D)])``.

        NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out."""
        if full_cov:
            raise NotImplementedError("full covariance between test points is not on this path")
        if quantiles is not None:
            raise NotImplementedError()       # as the reference, ssm_gpy/gaussian_process.py:561
        want_jac = bool(jacobians or compute_gradients)
        z_d, was_tensor = self._as_device_inputs(states, actions)
        out = self.predict_device(z_d, want_jac)
        if was_tensor:
            return out
        return tuple(o.cpu().numpy() for o in out)

    def predictive_gradients(self, x_new, grad_sigma=False):
        """ssm_gpy/gaussian_process.py:570-596: d mean / d input, (T, n_s_out, D)."""
        if grad_sigma:
            raise NotImplementedError("Gradient of sigma not implemented")
        return self.predict(x_new, compute_gradients=True)[2]

    def __call__(self, states, actions):
        """The exact triple onestep_reachability unpacks (gp_reachability.py:74,101) for a single input:
        ``(mu (n_s,1), var (n_s,1), jac (n_s,D))`` (SimpleGPModel.__call__ -> predict_casadi_symbolic,
        ssm_gpy/gaussian_process.py:135-175).  For T > 1 inputs the batch triple
        ``(mu (T,n_s), var (T,n_s), jac (T,n_s,D))`` is returned instead of raising."""
        mu, var, jac = self.predict(states, actions, jacobians=True)
        if mu.shape[0] == 1:
            if self._torch.is_tensor(mu):
                return mu.t(), var.t(), jac[0]
            return mu.T, var.T, jac[0]
        return mu, var, jac

    def linearize_predict(self, states, actions, jacobians=False, full_cov=False):
        raise NotImplementedError("Taylor-linearised prediction is a 'next' row (SURVEY.md section 8 f2)")

    def get_forward_model_casadi(self, linearize_mu=True):
        raise NotImplementedError("the CasADi/IPOPT bridge is what the batched sampling path replaces "
                                  "(SURVEY.md section 2); use safe_exploration_b200.gp_reachability")

    # ------------------------------------------------------------------ introspection
    @property
    def beta(self):
        """(N, n_s_out) = K^-1 y per output dimension (posterior.woodbury_vector)."""
        if not self.gp_trained:
            return None
        host = np.empty((self.n_s_out, self.z.shape[0]))
        self._torch.cuda.synchronize(self.device)
        _lib.check(self._lib.segp_beta(self._handle, _lib.dbl_ptr(host)))
        return np.ascontiguousarray(host.T)

    def log_det_k(self):
        """log det (K_d + noise_d I) per output dimension, from the Cholesky factor."""
        out = np.empty(self.n_s_out)
        _lib.check(self._lib.segp_logdet(self._handle, _lib.dbl_ptr(out)))
        return out

    def information_gain(self, x=None):
        """ssm_gpy/gaussian_process.py:621-634: log det(I + K / noise_var) per output dimension.
        log det(I + K/s) = log det(K + s I) - n log s, from the Cholesky factor on the device.  The diagonal term s is
        the one this model factorises (Gaussian noise + noise_diag + GPy jitter, relative shift 1e-5 / noise against
        the reference's bare Gaussian noise: documented difference).

        x=None (the reference's only working case: it reads the TRAINING kernel matrix ``posterior._K`` whatever x
        is, :631-632): the training inputs, from the model's own factor.  A foreign x (n, D): K(x, x) is built and
        factorised on the device with the same hyper-parameters."""
        if x is None:
            n = self.z.shape[0]
            return list(self.log_det_k() - n * np.log(self.total_noise()))
        x = _lib.host_f64(x)
        if x.ndim != 2 or x.shape[1] != self.dim_in:
            raise ValueError("x must be n x {}".format(self.dim_in))
        tmp = BatchedGPSSM(self.n_s_out, self.n_s_in, self.n_u, None, None, None, self.kern_types, self.hyp, False,
                           None, self.device.index, self.noise_diag, 0, self.composite_semantics)
        try:
            tmp.train(x, np.zeros((x.shape[0], self.n_s_out)))
            return list(tmp.log_det_k() - x.shape[0] * np.log(self.total_noise()))
        finally:
            tmp.close()

    def select_maxvar(self, x, m):
        """Indices (m,) and scores (m,) of the greedy maximum-predicted-variance selection among the rows of x
        (segp_select_maxvar: partial Cholesky factors on the device, fixed hyper-parameters of this model)."""
        x = _lib.host_f64(x)
        if x.ndim != 2 or x.shape[1] != self.dim_in:
            raise ValueError("x must be n x {}".format(self.dim_in))
        m = int(m)
        if m < 1 or m > x.shape[0]:
            raise ValueError("1 <= m <= n required")
        ls, var, pl, lin = (_lib.host_f64(v) for v in self._model_arrays())
        noise = _lib.host_f64(self.total_noise())
        kern_ids = (ctypes.c_int * self.n_s_out)(*[_lib.KERN_IDS[k] for k in self.kern_types])
        idx = np.empty(m, dtype=np.int32)
        score = np.empty(m)
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.segp_select_maxvar(
                self.device.index, x.shape[0], self.n_s_out, self.dim_in, kern_ids, _lib.dbl_ptr(x), _lib.dbl_ptr(ls),
                _lib.dbl_ptr(var), _lib.dbl_ptr(noise), _lib.dbl_ptr(pl) if self.has_composite else None,
                _lib.dbl_ptr(lin) if self.has_composite else None, m,
                idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _lib.dbl_ptr(score),
                _lib.current_stream(self.device)))
        return idx.astype(np.int64), score

    def choose_datapoints_maxvar(self, x, y, m, k=10, min_ratio_k=0.25, n_reopt_gp=1):
        """SimpleGPModel.choose_datapoints_maxvar (ssm_gpy/gaussian_process.py:280-345): pick m of the n data points
        by the maximum-predicted-variance criterion; returns (x_chosen, y_chosen) in selection order.

        Same criterion and loop (:320-343), run on the device.  Differences, on purpose: the initial set is empty
        instead of k random representatives of a k-means clustering (:303-318, non-deterministic), and the
        hyper-parameters stay fixed (the reference re-optimises them on the pool every (m-k)/(n_reopt_gp+1) steps,
        :327-330; GPy's optimiser is out of scope).  k, min_ratio_k, n_reopt_gp are accepted and ignored."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if x.shape[0] <= m:          # less data than the subset: the whole set (:295-296)
            return x, y
        idx, _ = self.select_maxvar(x, m)
        return x[idx], y[idx]

    def sample_from_gp(self, inp, size=10):
        """ssm_gpy/gaussian_process.py:598-619: ``size`` samples of the (independent, full_cov=False) predictive
        distribution per test input, (n, size, n_s_out); NumPy's global random state, like GPy."""
        mu, var = self.predict(np.asarray(inp, dtype=np.float64))
        std = np.sqrt(np.maximum(var, 0.0))
        return mu[:, None, :] + std[:, None, :] * np.random.standard_normal((mu.shape[0], int(size), self.n_s_out))

    def to_dict(self):
        """ssm_gpy/gaussian_process.py:177-187 (inv_K is not materialised on this path)."""
        return {"x": self.x_train, "y": self.y_train, "kern_types": self.kern_types, "hyp": self.hyp,
                "beta": self.beta, "inv_K": None, "n_s_in": self.n_s_in, "n_s_out": self.n_s_out, "n_u": self.n_u}

    @classmethod
    def from_dict(cls, gp_dict, device=None):
        """ssm_gpy/gaussian_process.py:72-133."""
        x = gp_dict.get("x")
        y = gp_dict.get("y")
        if x is None or y is None:
            warnings.warn("no data in gp_dict: the model is instantiated untrained")
        if "prior_model" in gp_dict and x is not None:
            y = y - gp_dict["prior_model"](x)
        return cls(gp_dict["n_s_out"], gp_dict["n_s_in"], gp_dict["n_u"], x, y, None, gp_dict.get("kern_types"),
                   gp_dict.get("hyp"), gp_dict.get("train", True), None, device)

    def set_option(self, name, value):
        _lib.check(self._lib.segp_set_option(self._handle, name.encode(), int(value)))

    def get_option(self, name):
        v = ctypes.c_long()
        _lib.check(self._lib.segp_get_option(self._handle, name.encode(), ctypes.byref(v)))
        return v.value

    def set_param(self, name, value):
        """Real-valued parameters of the precision guard (include/segp.h: "guard_rtol", "guard_kappa")."""
        _lib.check(self._lib.segp_set_param(self._handle, name.encode(), float(value)))

    def get_param(self, name):
        v = ctypes.c_double()
        _lib.check(self._lib.segp_get_param(self._handle, name.encode(), ctypes.byref(v)))
        return v.value

    def precision_report(self):
        """What the factorize-time probe measured and decided: digit set of the first
        contraction pass, whether automatic mode runs float64, and the probe statistics."""
        names = ("probe_ran", "probe_frac4", "probe_frac5", "probe_err4", "probe_err5", "probe_rel4", "probe_rel5",
                 "probe_ratio4", "probe_ratio5", "probe_rho4", "probe_rho5", "probe_min_var_ratio", "probe_margin4",
                 "guard_rtol", "guard_kappa")
        out = {n: self.get_param(n) for n in names}
        out["tri_mode_effective"] = self.get_option("tri_mode_effective")
        out["i8_digits_effective"] = self.get_option("i8_digits_effective")
        out["fallback_panels"] = self.get_option("fallback_panels")
        out["unguarded"] = self.get_option("unguarded")
        out["demoted"] = self.get_option("demoted")
        return out


def _cudart_memcpy_d2h(torch, host, dev_ptr, nbytes, device):
    """Copy raw device memory into a NumPy array through a uint8 torch view of the pointer."""
    view = _tensor_from_ptr(torch, dev_ptr, nbytes, device)
    host.view(np.uint8).reshape(-1)[:] = view.cpu().numpy()


def _tensor_from_ptr(torch, dev_ptr, nbytes, device):
    """uint8 CUDA tensor aliasing [dev_ptr, dev_ptr+nbytes) via the CUDA array interface."""

    class _Holder(object):
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(dev_ptr), False),
                                  "version": 2}
    return torch.as_tensor(h, device=device)
