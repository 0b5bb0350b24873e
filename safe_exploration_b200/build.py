"""Build libsegp.so (the C-ABI CUDA library of this package) in-tree with nvcc for sm_100a.

    python -m safe_exploration_b200.build [--force]

No GPU is needed to build (nvcc cross-compiles).  The product never falls back to anything else:
``safe_exploration_b200._lib`` raises if the library is missing or fails to load.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libsegp.so")

SOURCES = ["api.cu", "predict.cu", "setup.cu", "ellipsoid.cu", "diag.cu", "tri_i8.cu", "fact_i8.cu", "score.cu", "select.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libsegp.so cannot be built")
    return nvcc


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(nvcc, src, obj, log):
    extra = os.environ.get("SEGP_EXTRA_NVCC_FLAGS", "").split()   # tuning experiments only (e.g. -DSEGP_KS_UNROLL=16)
    cmd = [nvcc] + ARCH_FLAGS + NVCC_FLAGS + extra + ["-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for {}:\n{}".format(src, res.stderr[-4000:]))
    return obj


def build_library(force=False, verbose=True):
    """Compile every .cu under csrc/ for sm_100a and link libsegp.so.  Returns the library path."""
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    headers = [os.path.join(INCLUDE, "segp.h"), os.path.join(CSRC, "segp_internal.cuh"), os.path.join(CSRC, "tc_i8.cuh")]
    jobs = []
    objs = []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(BUILD_DIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj, os.path.join(BUILD_DIR, s.replace(".cu", ".log"))))
    if jobs:
        if verbose:
            print("[segp build] nvcc sm_100a:", ", ".join(os.path.basename(j[0]) for j in jobs), flush=True)
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            futs = [pool.submit(_compile_one, nvcc, *j) for j in jobs]
            for f in futs:
                f.result()
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stderr[-4000:])
        if verbose:
            print("[segp build] linked", LIB_PATH, flush=True)
    return LIB_PATH


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
