"""Ellipsoid reachability with the reference's call signatures, evaluated in batch on the GPU.

Drop-in mirror of reference safe_exploration/gp_reachability.py:

* ``onestep_reachability(p_center, ssm, k_ff, l_mu, l_sigma, q_shape=None, k_fb=None, c_safety=1.,
  verbose=1, a=None, b=None)``                                        (gp_reachability.py:19-156)
* ``multistep_reachability(p_0, gp, k_fb, k_ff, L_mu, L_sigm, q_0=None, c_safety=1., verbose=1,
  a=None, b=None, k_fb_init=None)``                                   (gp_reachability.py:159-212)
* ``lin_ellipsoid_safety_distance(p_center, q_shape, h_mat, h_vec, c_safety=1.)``  (:215-250)

Same names, same positional order, same return tuples.  Un-batched inputs (the reference's shapes)
return un-batched NumPy arrays, so the reference's own call sites (test/test_gp_reachability_casadi.py:
87-88, 133-137) run unchanged; arrays may also carry a leading batch axis B (``k_ff`` of shape
(B, H, n_u)), in which case B independent recursions run in one library call.  One extra optional
keyword, ``t_z_gp`` (the GP-input transform of gp_reachability_casadi.py:60-98).

If ``ssm`` is a BatchedGPSSM the GP posterior and the ellipsoid step are fused inside libsegp
(segp_multistep); any other callable with the reference's plugin signature
``ssm(states 1 x n_s, actions 1 x n_u) -> (mu, var, jac)`` is evaluated by the caller and only the
ellipsoid algebra runs on the GPU (segp_ellipsoid_step).

Per-trajectory failures (non-positive variance, a zero box bound, overflow) do not raise in batch mode:
they are reported in the int32 status mask of ``rollout`` (bits in _lib.STATUS_*).  For an un-batched
call a zero box bound raises AssertionError as the reference does (utils_ellipsoid.py:226-228).
"""
import collections

import numpy as np

from . import _lib
from .ssm import BatchedGPSSM

__all__ = ["onestep_reachability", "multistep_reachability", "lin_ellipsoid_safety_distance", "rollout", "pinned_result",
           "RolloutResult"]

RolloutResult = collections.namedtuple("RolloutResult", ["p_all", "q_all", "var_all", "status"])


def _is_tensor(x):
    try:
        import torch
        return torch.is_tensor(x)
    except ImportError:       # pragma: no cover
        return False


def pinned_result(gp, batch, horizon, want_var=True):
    """Page-locked host buffers for `rollout(..., out=...)`: a RolloutResult of NumPy views of pinned memory, to be
    allocated once and reused across calls (the device->host copy of a result into pageable NumPy memory costs ~5 %
    of a C4 call; page-locking per call would cost as much again)."""
    torch = gp._torch
    n_s = gp.n_s_out

    def pin(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()

    return RolloutResult(pin((batch, horizon, n_s), torch.float64), pin((batch, horizon, n_s, n_s), torch.float64),
                         pin((batch, horizon, n_s), torch.float64) if want_var else None,
                         pin((batch,), torch.int32))


def rollout(gp, p_0, k_ff, k_fb, l_mu, l_sigma, q_0=None, k_fb_init=None, c_safety=1., a=None, b=None,
            t_z_gp=None, want_var=True, propagation=0, out=None):
    """B independent H-step reachability recursions in one call (the batched core behind
    multistep_reachability).

    gp        BatchedGPSSM
    p_0       (n_s,) | (n_s,1) shared, or (B,n_s)
    k_ff      (B,H,n_u)
    k_fb      (H-1,n_u,n_s) shared or (B,H-1,n_u,n_s); may be None when H == 1
    q_0       None | (n_s,n_s) shared | (B,n_s,n_s);  k_fb_init (n_u,n_s) | (B,n_u,n_s), needed with q_0
    NumPy inputs use host buffers through segp_multistep_host and return NumPy; CUDA float64 tensors stay
    on the device (segp_multistep, asynchronous on the current stream) and return tensors.
    propagation  0 ellipsoid reachability (default); 1 / 2: q_all holds Gaussian covariances propagated by the
              first-order Taylor / mean-equivalent scheme (see uncertainty_propagation.py); l_mu, l_sigma, c_safety unused
    out       optional RolloutResult to write into: host arrays on the host path (see pinned_result), CUDA tensors on
              the device path.  Re-using the same buffers lets the library replay the call as a CUDA graph (one
              handle serves one call at a time: calls on the same model must not overlap on different streams)
    Returns RolloutResult(p_all (B,H,n_s), q_all (B,H,n_s,n_s), var_all (B,H,n_s) | None, status (B,) int32).
    """
    if not isinstance(gp, BatchedGPSSM):
        raise TypeError("rollout needs a BatchedGPSSM")
    if not gp.gp_trained:
        raise RuntimeError("model is not trained")
    lib = gp._lib
    n_s, n_u, n_in = gp.n_s_out, gp.n_u, gp.n_s_in
    on_device = _is_tensor(k_ff)
    prm, keep = _lib.make_reach_params(l_mu, l_sigma, c_safety, a, b, t_z_gp, n_s, n_u, n_in, propagation)
    if on_device:
        torch = gp._torch
        dev = gp.device

        def prep(x):
            if x is None:
                return None
            return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

        k_ff_d = prep(k_ff)
        if k_ff_d.dim() != 3 or k_ff_d.shape[2] != n_u:
            raise ValueError("k_ff must be (B, H, n_u)")
        bsz, hor = int(k_ff_d.shape[0]), int(k_ff_d.shape[1])
        p0_d = prep(p_0).reshape(-1)
        q0_d = prep(q_0)
        kfb_d = prep(k_fb) if hor > 1 else None
        kfbi_d = prep(k_fb_init)
        per = max(hor - 1, 0) * n_u * n_s
        # the kernels derive every address from these sizes: a mis-shaped tensor must fail here, not read out of bounds
        if p0_d.numel() not in (n_s, bsz * n_s):
            raise ValueError("p_0 must be (n_s,), (n_s,1) or (B,n_s)")
        if q0_d is not None and q0_d.numel() not in (n_s * n_s, bsz * n_s * n_s):
            raise ValueError("q_0 must be (n_s,n_s) or (B,n_s,n_s)")
        if hor > 1 and kfb_d is None:
            raise ValueError("k_fb is required for H > 1")
        if kfb_d is not None and kfb_d.numel() not in (per, bsz * per):
            raise ValueError("k_fb must be (H-1,n_u,n_s) or (B,H-1,n_u,n_s)")
        if kfbi_d is not None and kfbi_d.numel() not in (n_u * n_s, bsz * n_u * n_s):
            raise ValueError("k_fb_init must be (n_u,n_s) or (B,n_u,n_s)")
        if q0_d is not None and kfbi_d is None:
            raise ValueError("k_fb_init is required when q_0 is given")
        p0_stride = 0 if (p0_d.numel() == n_s or bsz == 1) else n_s
        q0_stride = 0 if (q0_d is None or q0_d.numel() == n_s * n_s) else n_s * n_s
        kfb_stride = 0 if (kfb_d is None or kfb_d.numel() == per) else per
        kfbi_stride = 0 if (kfbi_d is None or kfbi_d.numel() == n_u * n_s) else n_u * n_s
        if out is not None:
            p_all, q_all, var_all, status = out.p_all, out.q_all, (out.var_all if want_var else None), out.status
            for arr, shape, dt in ((p_all, (bsz, hor, n_s), torch.float64), (q_all, (bsz, hor, n_s, n_s), torch.float64),
                                   (var_all, (bsz, hor, n_s), torch.float64), (status, (bsz,), torch.int32)):
                if arr is not None and (not _is_tensor(arr) or not arr.is_cuda or tuple(arr.shape) != shape or
                                        arr.dtype != dt or not arr.is_contiguous()):
                    raise ValueError("out buffers must be contiguous CUDA tensors of shape {}".format(shape))
            if want_var and var_all is None:
                raise ValueError("out.var_all is required when want_var is true")
        else:
            p_all = torch.empty((bsz, hor, n_s), dtype=torch.float64, device=dev)
            q_all = torch.empty((bsz, hor, n_s, n_s), dtype=torch.float64, device=dev)
            var_all = torch.empty((bsz, hor, n_s), dtype=torch.float64, device=dev) if want_var else None
            status = torch.empty((bsz,), dtype=torch.int32, device=dev)
        _lib.check(lib.segp_multistep(gp._handle, bsz, hor, _lib.dev_ptr(p0_d), p0_stride, _lib.dev_ptr(q0_d),
                                      q0_stride, _lib.dev_ptr(k_ff_d), _lib.dev_ptr(kfb_d), kfb_stride,
                                      _lib.dev_ptr(kfbi_d), kfbi_stride, prm, _lib.dev_ptr(p_all),
                                      _lib.dev_ptr(q_all), _lib.dev_ptr(var_all), _lib.dev_ptr(status),
                                      _lib.current_stream(dev)))
        return RolloutResult(p_all, q_all, var_all, status)

    k_ff_h = _lib.host_f64(k_ff)
    if k_ff_h.ndim != 3 or k_ff_h.shape[2] != n_u:
        raise ValueError("k_ff must be (B, H, n_u)")
    bsz, hor = k_ff_h.shape[0], k_ff_h.shape[1]
    p0_h = _lib.host_f64(p_0).reshape(-1)
    if p0_h.size not in (n_s, bsz * n_s):
        raise ValueError("p_0 must be (n_s,), (n_s,1) or (B,n_s)")
    p0_stride = 0 if p0_h.size == n_s else n_s
    if bsz == 1:
        p0_stride = 0
    q0_h = _lib.host_f64(q_0).reshape(-1) if q_0 is not None else None
    q0_stride = 0 if (q0_h is None or q0_h.size == n_s * n_s) else n_s * n_s
    per = max(hor - 1, 0) * n_u * n_s
    kfb_h = _lib.host_f64(k_fb).reshape(-1) if (k_fb is not None and per > 0) else None
    if kfb_h is not None and kfb_h.size not in (per, bsz * per):
        raise ValueError("k_fb must be (H-1,n_u,n_s) or (B,H-1,n_u,n_s)")
    kfb_stride = 0 if (kfb_h is None or kfb_h.size == per) else per
    kfbi_h = _lib.host_f64(k_fb_init).reshape(-1) if k_fb_init is not None else None
    kfbi_stride = 0 if (kfbi_h is None or kfbi_h.size == n_u * n_s) else n_u * n_s
    if out is not None:
        p_all, q_all, var_all, status = out.p_all, out.q_all, (out.var_all if want_var else None), out.status
        for arr, shape, dt in ((p_all, (bsz, hor, n_s), np.float64), (q_all, (bsz, hor, n_s, n_s), np.float64),
                               (var_all, (bsz, hor, n_s), np.float64), (status, (bsz,), np.int32)):
            if arr is not None and (arr.shape != shape or arr.dtype != dt or not arr.flags["C_CONTIGUOUS"]):
                raise ValueError("out buffers must be C-contiguous {} arrays of shape {}".format(np.dtype(dt), shape))
        if want_var and var_all is None:
            raise ValueError("out.var_all is required when want_var is true")
    else:
        p_all = np.empty((bsz, hor, n_s))
        q_all = np.empty((bsz, hor, n_s, n_s))
        var_all = np.empty((bsz, hor, n_s)) if want_var else None
        status = np.zeros((bsz,), dtype=np.int32)

    def hp(x):
        return None if x is None else x.ctypes.data

    _lib.check(lib.segp_multistep_host(gp._handle, bsz, hor, hp(p0_h), p0_stride, hp(q0_h), q0_stride, hp(k_ff_h),
                                       hp(kfb_h), kfb_stride, hp(kfbi_h), kfbi_stride, prm, hp(p_all), hp(q_all),
                                       hp(var_all), hp(status)))
    return RolloutResult(p_all, q_all, var_all, status)


def _raise_on_status(status):
    st = int(status)
    if st & (_lib.STATUS_ZERO_BOUND | _lib.STATUS_BAD_VARIANCE):
        # the reference's ellipsoid_from_rectangle assertion (utils_ellipsoid.py:226-228)
        raise AssertionError("all elements of u_b need to be greater than zero!")


# ------------------------------------------------------------------------------------------ foreign ssm
def _ellipsoid_step_foreign(p, ssm, k_ff, l_mu, l_sigma, q, k_fb, c_safety, a, b, t_z_gp):
    """One step for B trajectories with a foreign state-space model: the caller-side ssm is evaluated one
    trajectory at a time (the reference's plugin contract is single-point), the ellipsoid algebra runs in
    segp_ellipsoid_step.  p (B,n_s), k_ff (B,n_u), q None|(B,n_s,n_s), k_fb (n_u,n_s)|(B,n_u,n_s)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    bsz, n_s = p.shape
    n_u = k_ff.shape[1]
    t_mat = None if t_z_gp is None else np.asarray(t_z_gp, dtype=np.float64)
    n_in = n_s if t_mat is None else t_mat.shape[0]
    dim = n_in + n_u
    mu = np.empty((bsz, n_s))
    var = np.empty((bsz, n_s))
    jac = np.zeros((bsz, n_s, dim))
    for i in range(bsz):
        x_bar = p[i:i + 1] if t_mat is None else p[i:i + 1] @ t_mat.T
        out = ssm(x_bar, k_ff[i:i + 1])
        mu[i] = np.real(np.asarray(out[0], dtype=np.float64)).reshape(n_s)
        var[i] = np.real(np.asarray(out[1], dtype=np.float64)).reshape(n_s)
        if q is not None:
            jac[i] = np.real(np.asarray(out[2], dtype=np.float64)).reshape(n_s, dim)
    dev = torch.device("cuda", torch.cuda.current_device())
    prm, keep = _lib.make_reach_params(l_mu, l_sigma, c_safety, a, b, t_mat, n_s, n_u, n_in)

    def up(x):
        return None if x is None else torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)

    mu_d, var_d, jac_d, p_d, q_d, kff_d = up(mu), up(var), up(jac), up(p), up(q), up(k_ff)
    kfb_d = up(k_fb) if q is not None else None
    kfb_stride = 0
    if kfb_d is not None and kfb_d.numel() != n_u * n_s:
        kfb_stride = n_u * n_s
    p1 = torch.empty((bsz, n_s), dtype=torch.float64, device=dev)
    q1 = torch.empty((bsz, n_s, n_s), dtype=torch.float64, device=dev)
    status = torch.empty((bsz,), dtype=torch.int32, device=dev)
    _lib.check(lib.segp_ellipsoid_step(dev.index, bsz, n_s, n_in, n_u, _lib.dev_ptr(mu_d), _lib.dev_ptr(var_d),
                                       _lib.dev_ptr(jac_d), _lib.dev_ptr(p_d), _lib.dev_ptr(q_d), _lib.dev_ptr(kff_d),
                                       _lib.dev_ptr(kfb_d), kfb_stride, prm, _lib.dev_ptr(p1), _lib.dev_ptr(q1),
                                       _lib.dev_ptr(status), _lib.current_stream(dev)))
    return p1.cpu().numpy(), q1.cpu().numpy(), status.cpu().numpy()


# ------------------------------------------------------------------------------------------ reference API
def onestep_reachability(p_center, ssm, k_ff, l_mu, l_sigma, q_shape=None, k_fb=None, c_safety=1., verbose=1,
                         a=None, b=None, t_z_gp=None):
    """Over-approximate the one-step reachable set of an ellipsoid E(p_center, q_shape) under the GP dynamics
    and the affine feedback law (reference gp_reachability.py:19-156; same arguments and return values).

    Un-batched (reference shapes): p_center (n_s,1), k_ff (n_u,1), q_shape None|(n_s,n_s), k_fb (n_u,n_s)
    -> (p_1 (n_s,1), q_1 (n_s,n_s)).
    Batched: p_center (B,n_s), k_ff (B,n_u), q_shape None|(n_s,n_s)|(B,n_s,n_s), k_fb (n_u,n_s)|(B,n_u,n_s)
    -> (p_1 (B,n_s), q_1 (B,n_s,n_s))."""
    tensors = _is_tensor(p_center) or _is_tensor(k_ff)
    if tensors:
        if not isinstance(ssm, BatchedGPSSM):
            raise TypeError("tensor inputs need a BatchedGPSSM")
        n_s, n_u = ssm.n_s_out, ssm.n_u
        p_t = p_center.reshape(-1, n_s)
        res = rollout(ssm, p_t, k_ff.reshape(-1, 1, n_u), None, l_mu, l_sigma, q_shape, k_fb, c_safety, a, b,
                      t_z_gp, want_var=False)
        return res.p_all[:, 0], res.q_all[:, 0]
    p_np = np.asarray(p_center, dtype=np.float64)
    kff_np = np.asarray(k_ff, dtype=np.float64)
    n_s_known = getattr(ssm, "num_states", None)
    unbatched = (p_np.ndim == 2 and p_np.shape[1] == 1 and kff_np.ndim == 2 and kff_np.shape[1] == 1 and
                 not (n_s_known == 1 and p_np.shape[0] > 1))
    n_u = kff_np.shape[0] if unbatched else kff_np.shape[-1]
    if unbatched:
        n_s = p_np.shape[0]
        p_b = p_np.reshape(1, n_s)
        kff_b = kff_np.reshape(1, n_u)
    else:
        p_b = np.atleast_2d(p_np)
        n_s = p_b.shape[1]
        kff_b = kff_np.reshape(p_b.shape[0], -1)
        n_u = kff_b.shape[1]
    q_np = None if q_shape is None else np.asarray(q_shape, dtype=np.float64)
    if isinstance(ssm, BatchedGPSSM):
        res = rollout(ssm, p_b, kff_b[:, None, :], None, l_mu, l_sigma, q_np, k_fb, c_safety, a, b, t_z_gp,
                      want_var=False)
        p1, q1, status = res.p_all[:, 0], res.q_all[:, 0], res.status
    else:
        q_b = None if q_np is None else np.broadcast_to(q_np, (p_b.shape[0], n_s, n_s))
        kfb_np = None if k_fb is None else np.asarray(k_fb, dtype=np.float64)
        p1, q1, status = _ellipsoid_step_foreign(p_b, ssm, kff_b, l_mu, l_sigma, q_b, kfb_np, c_safety, a, b, t_z_gp)
    if unbatched:
        _raise_on_status(status[0])
        return p1[0].reshape(n_s, 1), q1[0]
    return p1, q1


def multistep_reachability(p_0, gp, k_fb, k_ff, L_mu, L_sigm, q_0=None, c_safety=1., verbose=1, a=None, b=None,
                           k_fb_init=None, t_z_gp=None):
    """Ellipsoidal over-approximation of the multi-step reachable sets (reference gp_reachability.py:159-212;
    same arguments and return values).

    Un-batched (reference shapes): p_0 (n_s,1), k_fb (H-1,n_u,n_s), k_ff (H,n_u)
    -> (p_new (n_s,1), q_new (n_s,n_s), p_all (H,n_s), q_all (H,n_s,n_s)).
    Batched: k_ff (B,H,n_u), p_0 shared or (B,n_s), k_fb shared or (B,H-1,n_u,n_s)
    -> (p_new (B,n_s), q_new (B,n_s,n_s), p_all (B,H,n_s), q_all (B,H,n_s,n_s))."""
    tensors = _is_tensor(k_ff)
    if tensors:
        res = rollout(gp, p_0, k_ff, k_fb, L_mu, L_sigm, q_0, k_fb_init, c_safety, a, b, t_z_gp, want_var=False)
        return res.p_all[:, -1], res.q_all[:, -1], res.p_all, res.q_all
    kff_np = np.asarray(k_ff, dtype=np.float64)
    unbatched = kff_np.ndim == 2
    kff_b = kff_np[None] if unbatched else kff_np
    bsz, hor, n_u = kff_b.shape
    if isinstance(gp, BatchedGPSSM):
        res = rollout(gp, p_0, kff_b, k_fb, L_mu, L_sigm, q_0, k_fb_init, c_safety, a, b, t_z_gp, want_var=False)
        p_all, q_all, status = res.p_all, res.q_all, res.status
    else:
        kfb_np = np.asarray(k_fb, dtype=np.float64)
        n_s = kfb_np.shape[-1]
        p = np.broadcast_to(np.asarray(p_0, dtype=np.float64).reshape(-1, n_s), (bsz, n_s)).copy()
        q = None if q_0 is None else np.broadcast_to(np.asarray(q_0, dtype=np.float64), (bsz, n_s, n_s)).copy()
        p_all = np.empty((bsz, hor, n_s))
        q_all = np.empty((bsz, hor, n_s, n_s))
        status = np.zeros(bsz, dtype=np.int32)
        kfb_t = None if k_fb_init is None else np.asarray(k_fb_init, dtype=np.float64)
        for t in range(hor):
            if t > 0:
                kfb_t = kfb_np[t - 1] if kfb_np.ndim == 3 else kfb_np[:, t - 1]
            p, q, st = _ellipsoid_step_foreign(p, gp, kff_b[:, t], L_mu, L_sigm, q, kfb_t, c_safety, a, b, t_z_gp)
            p_all[:, t], q_all[:, t] = p, q
            status |= st
    if unbatched:
        _raise_on_status(status[0])
        n_s = p_all.shape[-1]
        return p_all[0, -1].reshape(n_s, 1), q_all[0, -1], p_all[0], q_all[0]
    return p_all[:, -1], q_all[:, -1], p_all, q_all


def lin_ellipsoid_safety_distance(p_center, q_shape, h_mat, h_vec, c_safety=1.0):
    """Distance between ellipsoid(s) and the polytope h_mat x <= h_vec (reference gp_reachability.py:215-250).

    Un-batched: p_center (n_s,1), q_shape (n_s,n_s) -> d (m,1).
    Batched: p_center (...,n_s), q_shape (...,n_s,n_s) -> d (...,m).  d < 0 elementwise == inside (safe)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    h_mat_h = _lib.host_f64(h_mat)
    m, n_s = h_mat_h.shape
    h_vec_h = _lib.host_f64(h_vec).reshape(-1)
    assert h_vec_h.size == m, "h_vec has to have shape m x 1"
    tensors = _is_tensor(p_center)
    if tensors:
        dev = p_center.device
        p_d = p_center.to(torch.float64).contiguous()
        q_d = q_shape.to(torch.float64).contiguous()
        lead = tuple(q_d.shape[:-2])
        unbatched = False
    else:
        p_np = np.asarray(p_center, dtype=np.float64)
        q_np = np.asarray(q_shape, dtype=np.float64)
        unbatched = q_np.ndim == 2
        if unbatched:
            assert p_np.shape == (n_s, 1), "p_center has to have shape n_s x 1"
            assert q_np.shape == (n_s, n_s), "q_shape has to have shape n_s x n_s"
        lead = tuple(q_np.shape[:-2])
        dev = torch.device("cuda", torch.cuda.current_device())
        p_d = torch.as_tensor(np.ascontiguousarray(p_np.reshape(-1, n_s)), device=dev)
        q_d = torch.as_tensor(np.ascontiguousarray(q_np.reshape(-1, n_s, n_s)), device=dev)
    n_items = int(np.prod(lead)) if lead else 1
    dist = torch.empty((n_items, m), dtype=torch.float64, device=dev)
    _lib.check(lib.segp_safety_distance(dev.index, n_items, n_s, m, _lib.dev_ptr(p_d), _lib.dev_ptr(q_d),
                                        _lib.dbl_ptr(h_mat_h), _lib.dbl_ptr(h_vec_h), float(c_safety),
                                        _lib.dev_ptr(dist), _lib.current_stream(dev)))
    if tensors:
        return dist.reshape(lead + (m,))
    out = dist.cpu().numpy()
    if unbatched:
        return out.reshape(m, 1)
    return out.reshape(lead + (m,))
