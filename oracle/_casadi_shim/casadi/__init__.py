"""One-function stand-in for the ``casadi`` module (ORACLE / TEST INFRASTRUCTURE).

The reference's ``safe_exploration/utils.py:14`` does ``from casadi import reshape``
at import time although none of the NumPy hot-path functions use it.  CasADi is not
installable in this image, so ``oracle/ref_loader.py`` puts this directory on
``sys.path`` to let the reference's own ``gp_reachability.py`` / ``utils.py`` /
``utils_ellipsoid.py`` import unmodified.  CasADi reshapes are column-major.
"""
import numpy as _np


def reshape(x, *shape):
    if len(shape) == 1:
        shape = shape[0]
    return _np.reshape(_np.asarray(x), shape, order="F")
