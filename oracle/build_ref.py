"""Recipe for oracle/_ref: the reference's OWN files of the path, byte for byte, so that the GPU box (which has no
/root/reference) can time and check the reference itself.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

The reference is pure Python with no native sources, so "building" it is placing its unmodified modules where the
loader finds them.  Run in the build container (`python -m oracle.build_ref`, also called by
__graft_entry__.build()); outputs go ONLY to oracle/_ref/, which is git-ignored (the reference's sources never
enter this repository's history) but travels to the GPU box with the snapshot, like a compiled library would.

    oracle/_ref/safe_exploration/{gp_reachability,utils,utils_ellipsoid,uncertainty_propagation_casadi}.py
    oracle/_ref/safe_exploration/ssm_gpy/gp_models_utils_casadi.py
    oracle/_ref/MANIFEST.json     sha256 of every file and of its source
"""
import hashlib
import json
import os
import shutil

REFERENCE_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [
    "safe_exploration/gp_reachability.py",
    "safe_exploration/utils.py",
    "safe_exploration/utils_ellipsoid.py",
    "safe_exploration/uncertainty_propagation_casadi.py",
    "safe_exploration/ssm_gpy/gp_models_utils_casadi.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True):
    """Returns True if oracle/_ref is in place (built now or before), False if the reference tree is absent and
    nothing was built earlier."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "safe_exploration")):
        return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {"sha256": _sha(dst), "source": src}
        assert _sha(src) == manifest[rel]["sha256"]
    # package markers (generated, empty: the reference's own __init__ files import GPy / CasADi-only modules)
    for pkg in ("safe_exploration", "safe_exploration/ssm_gpy"):
        open(os.path.join(OUT, pkg, "__init__.py"), "w").close()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print("[oracle/_ref] {} reference files placed under {}".format(len(FILES), OUT))
    return True


if __name__ == "__main__":
    build()
