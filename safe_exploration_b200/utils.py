"""Batched mirror of the one hot-path function of reference safe_exploration/utils.py.

* ``compute_remainder_overapproximations(q, k_fb, l_mu, l_sigma)``   utils.py:108-144
"""
import numpy as np

from . import _lib

__all__ = ["compute_remainder_overapproximations"]


def compute_remainder_overapproximations(q, k_fb, l_mu, l_sigma):
    """Lagrange-remainder boxes of the mean / std-deviation linearisation (utils.py:108-144):
    r^2 = max eig(q (I + k_fb^T k_fb));  u_mu = l_mu r^2;  u_sigma = l_sigma sqrt(r^2).

    q (n_s,n_s), k_fb (n_u,n_s) -> two (n_s,) arrays; batched q (B,n_s,n_s) with k_fb shared or (B,n_u,n_s)
    -> two (B,n_s) arrays.  Returns real float64 (the reference leaks complex128 with zero imaginary part
    from scipy.linalg.eig)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    q_np = np.asarray(q, dtype=np.float64)
    unbatched = q_np.ndim == 2
    n_s = q_np.shape[-1]
    kfb_np = np.asarray(k_fb, dtype=np.float64)
    n_u = kfb_np.shape[-2]
    q_d = torch.as_tensor(np.ascontiguousarray(q_np.reshape(-1, n_s, n_s)), device=dev)
    bsz = q_d.shape[0]
    kfb_d = torch.as_tensor(np.ascontiguousarray(kfb_np), device=dev)
    kfb_stride = 0 if kfb_np.ndim == 2 else n_u * n_s
    l_mu_h = _lib.host_f64(l_mu, (n_s,))
    l_sig_h = _lib.host_f64(l_sigma, (n_s,))
    u_mu = torch.empty((bsz, n_s), dtype=torch.float64, device=dev)
    u_sig = torch.empty((bsz, n_s), dtype=torch.float64, device=dev)
    _lib.check(lib.segp_remainder_overapproximations(dev.index, bsz, n_s, n_u, _lib.dev_ptr(q_d), _lib.dev_ptr(kfb_d),
                                                     kfb_stride, _lib.dbl_ptr(l_mu_h), _lib.dbl_ptr(l_sig_h),
                                                     _lib.dev_ptr(u_mu), _lib.dev_ptr(u_sig),
                                                     _lib.current_stream(dev)))
    um, us = u_mu.cpu().numpy(), u_sig.cpu().numpy()
    if unbatched:
        return um[0], us[0]
    return um, us
