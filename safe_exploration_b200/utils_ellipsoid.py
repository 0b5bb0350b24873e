"""Batched mirrors of the ellipsoid helpers on the hot path (reference safe_exploration/utils_ellipsoid.py).

* ``sum_two_ellipsoids(p_1, q_1, p_2, q_2, c=None)``   utils_ellipsoid.py:63-94
* ``ellipsoid_from_rectangle(u_b)``                    utils_ellipsoid.py:197-233

Reference shapes return reference shapes; a leading batch axis is accepted and evaluated in one kernel
launch through the C ABI (segp_sum_two_ellipsoids, segp_ellipsoid_from_rectangle).  Inside
onestep/multistep_reachability these steps are fused into the ellipsoid_step kernel; the stand-alone
entry points exist for callers that use the helpers directly and for the parity tests.
"""
import numpy as np

from . import _lib

__all__ = ["sum_two_ellipsoids", "ellipsoid_from_rectangle"]


def _dev(torch):
    return torch.device("cuda", torch.cuda.current_device())


def sum_two_ellipsoids(p_1, q_1, p_2, q_2, c=None):
    """Trace-minimal outer ellipsoid of the Minkowski sum of E(p_1,q_1) and E(p_2,q_2)
    (utils_ellipsoid.py:63-94).  p (n,1) / q (n,n), or batched p (B,n) / q (B,n,n)."""
    q1 = np.asarray(q_1, dtype=np.float64)
    q2 = np.asarray(q_2, dtype=np.float64)
    unbatched = q1.ndim == 2
    n = q1.shape[-1]
    if c is not None:
        # explicit scaling: plain arithmetic, nothing to accelerate (utils_ellipsoid.py:91-93)
        p_new = np.asarray(p_1, dtype=np.float64) + np.asarray(p_2, dtype=np.float64)
        return p_new, (1 + (1. / c)) * q1 + (1 + c) * q2
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = _dev(torch)
    bsz = 1 if unbatched else q1.shape[0]

    def up(x, shape):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(shape)), device=dev)

    p1_d, p2_d = up(p_1, (bsz, n)), up(p_2, (bsz, n))
    q1_d, q2_d = up(q1, (bsz, n, n)), up(q2, (bsz, n, n))
    p_d = torch.empty_like(p1_d)
    q_d = torch.empty_like(q1_d)
    _lib.check(lib.segp_sum_two_ellipsoids(dev.index, bsz, n, _lib.dev_ptr(p1_d), _lib.dev_ptr(q1_d),
                                           _lib.dev_ptr(p2_d), _lib.dev_ptr(q2_d), _lib.dev_ptr(p_d),
                                           _lib.dev_ptr(q_d), _lib.current_stream(dev)))
    p_out, q_out = p_d.cpu().numpy(), q_d.cpu().numpy()
    if unbatched:
        return p_out.reshape(np.shape(p_1)), q_out[0]
    return p_out, q_out


def ellipsoid_from_rectangle(u_b):
    """Smallest-volume ellipsoid around the box [-u_b, u_b]: diag(n u_b^2) (utils_ellipsoid.py:197-233).
    u_b (n,) -> (n,n); batched (B,n) -> (B,n,n).  Raises AssertionError like the reference for a 2-D
    un-batched input of the wrong kind or non-positive bounds."""
    ub = np.asarray(u_b, dtype=np.float64)
    unbatched = ub.ndim == 1
    assert ub.ndim in (1, 2), "lb and ub need to be 1-dimensional (1darrays)!"
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = _dev(torch)
    ub2 = np.ascontiguousarray(ub.reshape(-1, ub.shape[-1]))
    bsz, n = ub2.shape
    ub_d = torch.as_tensor(ub2, device=dev)
    q_d = torch.empty((bsz, n, n), dtype=torch.float64, device=dev)
    st_d = torch.empty((bsz,), dtype=torch.int32, device=dev)
    _lib.check(lib.segp_ellipsoid_from_rectangle(dev.index, bsz, n, _lib.dev_ptr(ub_d), _lib.dev_ptr(q_d),
                                                 _lib.dev_ptr(st_d), _lib.current_stream(dev)))
    assert not bool((st_d != 0).any().item()), "all elements of u_b need to be greater than zero!"
    q = q_d.cpu().numpy()
    return q[0] if unbatched else q
