"""Minimal NumPy-backed stand-in for the ``casadi`` module (ORACLE / TEST INFRASTRUCTURE).

CasADi is not installable in this image.  ``oracle/ref_loader.py`` puts this directory on ``sys.path`` so that the
reference's own modules import UNMODIFIED and run on NumPy arrays:

* ``safe_exploration/utils.py:14`` does ``from casadi import reshape`` at import time (none of the NumPy hot-path
  functions use it): ``reshape`` (column-major, like CasADi);
* ``safe_exploration/uncertainty_propagation_casadi.py:8`` does ``from casadi import *`` and then uses ``mtimes``,
  ``vertcat``, ``horzcat``, ``diag``, ``MX.eye``, ``MX.zeros`` and -- through CasADi's own star export -- ``np``.

* ``safe_exploration/ssm_gpy/gp_models_utils_casadi.py:13-14`` imports ``mtimes, exp, sum2, repmat, Function, sqrt,
  vertcat, horzcat, SX, reshape``; its kernels, ``_unscaled_dist`` and ``gp_pred`` then evaluate on NumPy arrays
  (``gp_pred_function`` builds a symbolic ``Function`` and is not usable).

Only numeric evaluation is provided; nothing symbolic.
"""
import numpy as np
import numpy as _np


def reshape(x, *shape):
    if len(shape) == 1:
        shape = shape[0]
    return _np.reshape(_np.asarray(x), shape, order="F")


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    out = _np.asarray(args[0])
    for m in args[1:]:
        out = _np.dot(out, _np.asarray(m))
    return out


def vertcat(*args):
    return _np.vstack([_np.atleast_2d(_np.asarray(a)) for a in args])


def horzcat(*args):
    return _np.hstack([_np.atleast_2d(_np.asarray(a)) for a in args])


def diag(x):
    """CasADi semantics: a vector (n x 1 or 1 x n) gives the diagonal matrix, a matrix gives its diagonal (n x 1)."""
    x = _np.asarray(x)
    if x.ndim <= 1 or 1 in x.shape:
        return _np.diag(x.reshape(-1))
    return _np.diag(x).reshape(-1, 1)


class MX(object):
    """``MX.eye(n)`` / ``MX.zeros(n[, m])`` / ``SX(n)`` (an n x 1 matrix of zeros, ssm_gpy/gp_models_utils_casadi.py:24)."""

    def __new__(cls, *shape):
        return cls.zeros(*shape)

    @staticmethod
    def eye(n):
        return _np.eye(n)

    @staticmethod
    def zeros(*shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if len(shape) == 1:
            shape = (shape[0], 1)
        return _np.zeros(shape)


SX = MX


# ---- what ssm_gpy/gp_models_utils_casadi.py:13-14 imports (kernels, _unscaled_dist, gp_pred), numerically
def exp(x):
    return _np.exp(_np.asarray(x, dtype=float))


def sqrt(x):
    return _np.sqrt(_np.asarray(x, dtype=float))


def sum2(x):
    """Row sums as a column (CasADi: sum over the second dimension)."""
    return _np.sum(_np.atleast_2d(_np.asarray(x)), axis=1, keepdims=True)


def sum1(x):
    return _np.sum(_np.atleast_2d(_np.asarray(x)), axis=0, keepdims=True)


def repmat(x, n, m=1):
    return _np.tile(_np.atleast_2d(_np.asarray(x)), (int(n), int(m)))


class Function(object):
    """Symbolic function objects cannot be built numerically: gp_pred_function is not usable through this shim."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("casadi.Function is not available in the NumPy-backed shim")
