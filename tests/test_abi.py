"""CPU tests of the drop-in boundary: libsegp.so builds for sm_100a, loads without a GPU, exports every
symbol include/segp.h declares, and the ctypes table of the Python mirror matches the header."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "segp.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(segp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from safe_exploration_b200 import build, _lib
    build.build_library(verbose=False)
    return _lib.load()


def test_header_declares_the_path():
    names = _declared_functions()
    for must in ("segp_create", "segp_set_model", "segp_factorize", "segp_predict", "segp_multistep",
                 "segp_multistep_host", "segp_ellipsoid_step", "segp_safety_distance", "segp_destroy"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), "libsegp.so does not export " + name


def test_ctypes_table_matches_header(lib):
    from safe_exploration_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_functions()
    assert lib.segp_abi_version() == 2


def test_no_cpp_or_torch_types_cross_the_boundary():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    for banned in ("std::", "torch", "at::Tensor", "template"):
        assert banned not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def test_library_is_sm100a_sass_with_tcgen05_dmma_and_bulk_copies():
    """The shipped binary holds sm_100a SASS: tcgen05 int8 MMAs (UTCIMMA) with TMEM loads (LDTM) for the variance
    contraction, the FP64 tensor-pipe instruction (DMMA) of its float64 variant and of the factorisation, and TMA
    bulk copies (UBLKCP) with mbarrier transactions (SYNCS)."""
    from safe_exploration_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "UTCIMMA" in out.stdout and "LDTM" in out.stdout
    assert "DMMA" in out.stdout
    assert "UBLKCP" in out.stdout and "SYNCS" in out.stdout
    # per kernel: the contraction AND the factorisation GEMM issue tcgen05 MMAs and read TMEM
    per, cur = {}, None
    for line in out.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = ""
        elif cur is not None:
            per[cur] += line
    for frag in ("tri_i8m_kernel", "tri_i8mp_kernel", "gemm_i8d_kernel"):
        hits = [k for k in per if frag in k]
        assert hits, frag
        for k in hits:
            assert "UTCIMMA" in per[k] and "LDTM" in per[k] and "UBLKCP" in per[k], k


def test_product_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import safe_exploration_b200 as se
    with pytest.raises(RuntimeError):
        se.BatchedGPSSM(2, 2, 1)
    with pytest.raises(RuntimeError):
        se.lin_ellipsoid_safety_distance(np.zeros((2, 1)), np.eye(2), np.eye(2), np.ones((2, 1)))
    # the C ABI itself reports the missing device instead of computing anything on the host
    h = ctypes.c_void_p()
    kern = (ctypes.c_int * 2)(0, 0)
    rc = lib.segp_create(ctypes.byref(h), 0, 2, 2, 1, kern)
    assert rc != 0 and not h.value
    assert lib.segp_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "safe_exploration_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "/root/reference" not in src, fn


def test_workloads_are_frozen_and_sized_like_baseline():
    from safe_exploration_b200 import workloads
    w1 = workloads.make("C2", batch=8)
    w2 = workloads.make("C2", batch=8)
    assert np.array_equal(w1.x_train, w2.x_train) and np.array_equal(w1.k_ff, w2.k_ff)
    assert (w1.n_s, w1.n_u, w1.n_train, w1.horizon) == (2, 1, 500, 10)
    w4 = workloads.make("C4", batch=2, n_train=64)
    assert (w4.n_s, w4.n_u, w4.horizon) == (4, 1, 20) and w4.k_fb.shape == (19, 1, 4)
    cfg = workloads.CONFIGS
    assert cfg["C4"][3:6] == (5000, 20, 65536) and cfg["C5"][3:6] == (10000, 30, 131072)
    assert workloads.flop_per_step(4, 1, 5000) == 4 * (5000 ** 2 + 5000 * 34)
    # closed loop of the linear prior is stable for every system
    for name in ("C2", "C3", "C5"):
        w = workloads.make(name, batch=1, n_train=16)
        assert np.max(np.abs(np.linalg.eigvals(w.a + w.b @ w.k_fb[0]))) < 1.0


def test_ctypes_structs_match_the_c_layout(tmp_path):
    """The ctypes mirrors of segp_reach_params / segp_score_params must have the size and field offsets the C compiler
    gives the structs of include/segp.h (a field added on one side only would silently shift every later field)."""
    import ctypes
    import shutil
    import subprocess
    from safe_exploration_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"segp_reach_params": _lib.ReachParams, "segp_score_params": _lib.ScoreParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "segp.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run([gcc, "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
