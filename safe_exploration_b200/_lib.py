"""ctypes binding of libsegp.so (include/segp.h).  No CPU fallback: a missing or unloadable library raises.

PyTorch tensors appear here only as device-memory handles (``tensor.data_ptr()``) and for the
current CUDA stream; no torch type crosses the C ABI.
"""
import ctypes
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libsegp.so")

SEGP_OK = 0
SEGP_ERR_INVALID = 1
SEGP_ERR_CUDA = 2
SEGP_ERR_NOT_POSDEF = 3
SEGP_ERR_NOT_TRAINED = 4
SEGP_ERR_UNSUPPORTED = 5

KERN_IDS = {"rbf": 0, "mat52": 1, "lin_rbf": 2, "lin_mat52": 3}
COMPOSITE_KERNELS = ("lin_rbf", "lin_mat52")

STATUS_NONFINITE = 1
STATUS_BAD_VARIANCE = 2
STATUS_ZERO_BOUND = 4
STATUS_LOW_PRECISION = 8

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_long = ctypes.c_long
_int = ctypes.c_int
_dbl = ctypes.c_double


class ReachParams(ctypes.Structure):
    """struct segp_reach_params (include/segp.h)."""
    _fields_ = [("h_l_mu", _c_double_p), ("h_l_sigma", _c_double_p), ("c_safety", _dbl),
                ("h_a", _c_double_p), ("h_b", _c_double_p), ("h_t_z_gp", _c_double_p), ("propagation", _int)]


PROP_ELLIPSOID = 0
PROP_TAYLOR = 1
PROP_MEAN_EQUIVALENT = 2


class ScoreParams(ctypes.Structure):
    """struct segp_score_params (include/segp.h)."""
    _fields_ = [("h_u_min", _c_double_p), ("h_u_max", _c_double_p), ("m_obs", _int), ("h_mat_obs", _c_double_p),
                ("h_obs", _c_double_p), ("m_safe", _int), ("h_mat_safe", _c_double_p), ("h_safe", _c_double_p),
                ("c_safety", _dbl), ("eps_constraints", _dbl), ("cost_type", _int), ("eps_noise", _dbl),
                ("h_wx", _c_double_p), ("h_wu", _c_double_p), ("h_x_ref", _c_double_p), ("layout", _int),
                ("h_q0", _c_double_p), ("h_k_fb_0", _c_double_p)]


COST_EXPLORATION = 0
COST_QUADRATIC = 1
SCORE_SAFEMPC = 0
SCORE_CAUTIOUS = 1

# name -> (restype, argtypes); kept in one table so tests can check it against include/segp.h
PROTOTYPES = {
    "segp_abi_version": (_int, []),
    "segp_last_error": (ctypes.c_char_p, []),
    "segp_create": (_int, [ctypes.POINTER(_vp), _int, _int, _int, _int, _c_int_p]),
    "segp_destroy": (_int, [_vp]),
    "segp_set_model": (_int, [_vp, _int, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p]),
    "segp_set_linear_terms": (_int, [_vp, _c_double_p, _c_double_p]),
    "segp_factorize": (_int, [_vp, _vp]),
    "segp_append": (_int, [_vp, _int, _c_double_p, _c_double_p, _vp]),
    "segp_alloc_factor_buffers": (_int, [_vp]),
    "segp_alloc_fp64_operand": (_int, [_vp]),
    "segp_num_factor_buffers": (_int, [_vp]),
    "segp_factor_buffer": (_int, [_vp, _int, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "segp_mark_factorized": (_int, [_vp]),
    "segp_beta": (_int, [_vp, _c_double_p]),
    "segp_logdet": (_int, [_vp, _c_double_p]),
    "segp_select_maxvar": (_int, [_int, _int, _int, _int, _c_int_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                  _c_double_p, _c_double_p, _int, _c_int_p, _c_double_p, _vp]),
    "segp_predict": (_int, [_vp, _long, _vp, _vp, _vp, _vp, _vp]),
    "segp_predict_ex": (_int, [_vp, _long, _vp, _vp, _vp, _vp, _vp, _vp]),
    "segp_multistep": (_int, [_vp, _long, _int, _vp, _long, _vp, _long, _vp, _vp, _long, _vp, _long,
                              ctypes.POINTER(ReachParams), _vp, _vp, _vp, _vp, _vp]),
    "segp_multistep_host": (_int, [_vp, _long, _int, _vp, _long, _vp, _long, _vp, _vp, _long, _vp, _long,
                                   ctypes.POINTER(ReachParams), _vp, _vp, _vp, _vp]),
    "segp_ellipsoid_step": (_int, [_int, _long, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _long,
                                   ctypes.POINTER(ReachParams), _vp, _vp, _vp, _vp]),
    "segp_remainder_overapproximations": (_int, [_int, _long, _int, _int, _vp, _vp, _long, _c_double_p,
                                                 _c_double_p, _vp, _vp, _vp]),
    "segp_sum_two_ellipsoids": (_int, [_int, _long, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "segp_ellipsoid_from_rectangle": (_int, [_int, _long, _int, _vp, _vp, _vp, _vp]),
    "segp_safety_distance": (_int, [_int, _long, _int, _int, _vp, _vp, _c_double_p, _c_double_p, _dbl, _vp, _vp]),
    "segp_score_num_constraints": (_int, [_int, _int, ctypes.POINTER(ScoreParams)]),
    "segp_score_rollouts": (_int, [_int, _long, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _long, _vp,
                                   ctypes.POINTER(ScoreParams), _vp, _vp, _vp, _vp, _vp]),
    "segp_argbest": (_int, [_int, _long, _vp, _vp, _vp, ctypes.POINTER(_long), ctypes.POINTER(_dbl),
                            ctypes.POINTER(_dbl), ctypes.POINTER(_int), _vp]),
    "segp_dmma_peak": (_int, [_int, _int, ctypes.POINTER(_dbl)]),
    "segp_i8_peak": (_int, [_int, _int, _int, ctypes.POINTER(_dbl)]),
    "segp_i8_peak_pattern": (_int, [_int, _int, _int, _int, ctypes.POINTER(_dbl)]),
    "segp_i8_selftest": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp]),
    "segp_i8_gemm_selftest": (_int, [_int, _int, _int, _int, _vp, _vp, _vp, _dbl, _dbl, _int, _int]),
    "segp_set_option": (_int, [_vp, ctypes.c_char_p, _long]),
    "segp_get_option": (_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_long)]),
    "segp_set_param": (_int, [_vp, ctypes.c_char_p, _dbl]),
    "segp_get_param": (_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_dbl)]),
}

_lib = None


def load():
    """Load libsegp.so (once).  Raises RuntimeError when it is missing: there is no other backend."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsegp.so is not built ({}). Run `python -m safe_exploration_b200.build` (needs nvcc). "
            "safe_exploration_b200 has no CPU or PyTorch fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError here == library/header mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    msg = load().segp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Map a status code to the exception type the reference raises in the same situation
    (SURVEY.md section 8b 'Errors')."""
    if rc == SEGP_OK:
        return
    msg = last_error()
    if rc == SEGP_ERR_INVALID:
        raise ValueError(msg)
    if rc == SEGP_ERR_NOT_POSDEF:
        raise np.linalg.LinAlgError(msg)
    if rc == SEGP_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg or "libsegp error {}".format(rc))


def dbl_ptr(arr):
    """Host float64 C-contiguous ndarray -> double* (the caller keeps `arr` alive)."""
    if arr is None:
        return None
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(_c_double_p)


def host_f64(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if shape is not None:
        a = np.ascontiguousarray(a.reshape(shape))
    return a


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("safe_exploration_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    return torch


def dev_ptr(t):
    """torch CUDA tensor (contiguous) -> void* device pointer."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def current_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_reach_params(l_mu, l_sigma, c_safety, a, b, t_z_gp, n_s, n_u, n_in, propagation=0):
    """Build the shared-parameter struct; returns (struct, keepalive list)."""
    l_mu = host_f64(l_mu, (n_s,))
    l_sigma = host_f64(l_sigma, (n_s,))
    keep = [l_mu, l_sigma]
    a_h = b_h = t_h = None
    if a is not None:
        a_h = host_f64(a, (n_s, n_s))
        b_h = host_f64(b if b is not None else np.zeros((n_s, n_u)), (n_s, n_u))
        keep += [a_h, b_h]
    elif b is not None:
        a_h = host_f64(np.eye(n_s))
        b_h = host_f64(b, (n_s, n_u))
        keep += [a_h, b_h]
    if t_z_gp is not None:
        t_h = host_f64(t_z_gp, (n_in, n_s))
        keep.append(t_h)
    prm = ReachParams(dbl_ptr(l_mu), dbl_ptr(l_sigma), float(c_safety), dbl_ptr(a_h), dbl_ptr(b_h), dbl_ptr(t_h),
                      int(propagation))
    return prm, keep
