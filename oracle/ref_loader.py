"""Import the reference's OWN modules of the path, unmodified.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

In the build container the modules are imported from where they lie under /root/reference; on the GPU box (no
/root/reference) from oracle/_ref, the byte-for-byte placement oracle/build_ref.py makes at build time (git-ignored,
travels with the snapshot).  Callers must check ``available()`` and skip otherwise.  oracle/_casadi_shim on sys.path
satisfies ``from casadi import reshape`` (reference safe_exploration/utils.py:14).
"""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "_casadi_shim")


def _root():
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "safe_exploration", "gp_reachability.py")):
            return cand
    return None


REFERENCE_ROOT = _root() or "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "safe_exploration", "gp_reachability.py"))


def origin():
    """'reference tree' (/root/reference) or 'oracle/_ref' (its byte-for-byte placement), or None."""
    if not available():
        return None
    return "reference tree" if REFERENCE_ROOT == "/root/reference" else "oracle/_ref"


def _prepare_path():
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    have_casadi = True
    try:
        import casadi  # noqa: F401  (a real CasADi wins if one is ever installed)
    except ImportError:
        have_casadi = False
    if not have_casadi and _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)


def load():
    """Returns (gp_reachability, utils, utils_ellipsoid) modules of the reference."""
    _prepare_path()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from safe_exploration import gp_reachability, utils, utils_ellipsoid
    return gp_reachability, utils, utils_ellipsoid


def load_uncertainty_propagation():
    """The reference's uncertainty_propagation_casadi module (one_step_taylor, multi_step_taylor_symbolic,
    one_step_mean_equivalent, mean_equivalent_multistep), evaluated numerically through the NumPy-backed shim."""
    _prepare_path()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from safe_exploration import uncertainty_propagation_casadi
    return uncertainty_propagation_casadi


def load_gp_utils():
    """The reference's ssm_gpy/gp_models_utils_casadi.py (kernel functions _k_rbf / _k_mat52 / _k_lin / _k_lin_rbf /
    _k_lin_mat52, _unscaled_dist, gp_pred), loaded BY FILE -- the ssm_gpy package __init__ needs GPy, the module itself
    only NumPy and CasADi -- and evaluated numerically through the NumPy-backed shim."""
    import importlib.util
    _prepare_path()
    path = os.path.join(REFERENCE_ROOT, "safe_exploration", "ssm_gpy", "gp_models_utils_casadi.py")
    spec = importlib.util.spec_from_file_location("_ref_gp_models_utils_casadi", path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod
