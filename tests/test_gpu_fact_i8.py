"""GPU tests of the factorisation GEMMs on tcgen05 (csrc/fact_i8.cu: 8 int8 digit planes per operand, 36 products),
through the C ABI.

* the split + GEMM kernel against NumPy float64 on host matrices (every flag combination the factorisation uses),
  including rows of very different magnitude: the error must stay at float64 level relative to |A| |B|^T;
* segp_factorize with the tensor-core GEMMs (option fact_i8 = 1) against the float64 DMMA path (fact_i8 = 0) on the same
  model: beta, log-determinant, predictive mean / variance / Jacobian, and both against the float64 oracle.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A_LOWER, B_LOWER, C_LOWER = 1, 2, 4


def _tol(a, b_nk, extra=0.0):
    """Error bound of the digit-plane product, entry (i, j): every operand row is fixed-point with 63 bits below its own
    max-abs (each element first rounded to 53 bits), and the digit pairs a + c >= 8 are dropped, so next to the
    float64-level term relative to |a_i|^T |b_j| (scaling by 1 / max, rounding to the digit grid and the epilogue's three
    multiplications: <= 1e-15 in the worst case, far below the K * 1.1e-16 of a float64 dot product) there is a NORM-wise
    one, ~ 2^-61 K max|a_i| max|b_j|."""
    k = a.shape[1]
    amax = np.abs(a).max(axis=1)[:, None]
    bmax = np.abs(b_nk).max(axis=1)[None, :]
    return 2.0 ** -61 * k * amax * bmax + 2e-15 * (np.abs(a) @ np.abs(b_nk).T + extra)


@pytest.fixture(scope="module")
def se():
    import safe_exploration_b200 as pkg
    pkg._lib.load()
    return pkg


def _gemm(se, a, b, c, alpha, beta, trans_b, flags):
    lib = se._lib.load()
    m, k = a.shape
    n = b.shape[0] if trans_b else b.shape[1]
    out = np.ascontiguousarray(c, dtype=np.float64).copy()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    se._lib.check(lib.segp_i8_gemm_selftest(0, m, n, k, a.ctypes.data_as(ctypes.c_void_p),
                                            b.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                                            float(alpha), float(beta), int(trans_b), int(flags)))
    return out


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 384, 320), (384, 256, 1024)])
def test_gemm_general(se, m, n, k):
    rng = np.random.default_rng(m + n + k)
    # rows spanning nine orders of magnitude: the scale is per operand row, so every row keeps its 55 bits
    a = rng.standard_normal((m, k)) * 10.0 ** rng.uniform(-6, 3, size=(m, 1))
    b = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-6, 3, size=(n, 1))
    c0 = rng.standard_normal((m, n))
    got = _gemm(se, a, b, c0, -1.0, 1.0, 1, 0)
    want = c0 - a @ b.T
    assert np.all(np.abs(got - want) <= _tol(a, b, np.abs(c0)))
    got0 = _gemm(se, a, b, c0, 2.5, 0.0, 1, 0)          # beta = 0: C is not read (NaN there must not leak)
    got0n = _gemm(se, a, b, np.full((m, n), np.nan), 2.5, 0.0, 1, 0)
    assert np.array_equal(got0, got0n)
    assert np.all(np.abs(got0 - 2.5 * (a @ b.T)) <= 2.5 * _tol(a, b))
    # and it IS float64-grade where it matters: relative to the row norms the error sits at 1e-16
    nrm = np.linalg.norm(a, axis=1)[:, None] * np.linalg.norm(b, axis=1)[None, :]
    assert np.max(np.abs(got0 - 2.5 * (a @ b.T)) / nrm) < 1e-14


def test_gemm_syrk_lower(se):
    """The potrf trailing update: C -= P P^T, tiles on or below the diagonal only, the rest untouched."""
    rng = np.random.default_rng(7)
    m, k = 512, 256
    p = rng.standard_normal((m, k)) * 10.0 ** rng.uniform(-3, 1, size=(m, 1))
    c0 = rng.standard_normal((m, m))
    got = _gemm(se, p, p, c0, -1.0, 1.0, 1, C_LOWER)
    want = c0 - p @ p.T
    r, c = np.indices((m, m))
    computed = (c // 64) * 64 <= (r // 128) * 128 + 127     # 128 x 64 tiles that touch the lower triangle
    assert np.all((np.abs(got - want) <= _tol(p, p, np.abs(c0)))[computed])
    assert np.array_equal(got[~computed], c0[~computed])
    assert np.all(computed[r >= c])


def test_gemm_triangular_operands(se):
    """The two products of a trtri level: T = L21 W11 (W11 lower, k x n) and W21 = -W22 T (W22 lower)."""
    rng = np.random.default_rng(11)
    s = 512
    l21 = rng.standard_normal((s, s))
    w11 = np.tril(rng.standard_normal((s, s)))
    w11_dirty = w11 + np.triu(np.full((s, s), np.nan), 1)        # the upper part must never be read
    t_got = _gemm(se, l21, w11_dirty, np.zeros((s, s)), 1.0, 0.0, 0, B_LOWER)
    t_want = l21 @ w11
    assert np.all(np.abs(t_got - t_want) <= _tol(l21, w11.T))
    w22 = np.tril(rng.standard_normal((s, s)))
    w22_dirty = w22 + np.triu(np.full((s, s), np.nan), 1)
    got = _gemm(se, w22_dirty, t_want, np.zeros((s, s)), -1.0, 0.0, 0, A_LOWER)
    want = -w22 @ t_want
    assert np.all(np.abs(got - want) <= _tol(w22, t_want.T))


@pytest.mark.parametrize("n_train,kern", [(1500, "rbf"), (2000, "mat52"), (1100, "rbf"), (5000, "rbf")])
def test_factorize_on_tensor_cores_matches_dmma_path(se, n_train, kern):
    import torch
    from oracle.gp_oracle import GPOracle
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=256, n_train=n_train, kern=kern)
    res = {}
    for mode in (0, 1):
        gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, None, None, kern_types=w.kern_types, hyp=w.hyp, device=0)
        gp.set_option("fact_i8", mode)
        gp.set_option("tri_mode", 0)          # float64 contraction: what is compared is the factorisation
        gp.train(w.x_train, w.y_train)
        assert gp.get_option("fact_i8_effective") == mode
        z = np.concatenate([w.p0 + 0.05 * np.random.default_rng(3).standard_normal((256, w.n_s)),
                            w.k_ff[:, 0, :]], axis=1)
        mu, var = gp.predict(z)
        jac = gp.predictive_gradients(z)
        res[mode] = (gp.beta.copy(), gp.log_det_k().copy(), np.array(mu), np.array(var), np.array(jac))
        noise = gp.total_noise()
        gp.close()
        torch.cuda.synchronize()
    b0, ld0, mu0, var0, jac0 = res[0]
    b1, ld1, mu1, var1, jac1 = res[1]
    print("n=%d %s: beta %.2e logdet %.2e mu %.2e var %.2e jac %.2e" % (
        n_train, kern, np.abs(b1 - b0).max() / np.abs(b0).max(), np.abs(ld1 - ld0).max() / np.abs(ld0).max(),
        np.abs(mu1 - mu0).max() / np.abs(mu0).max(), np.max(np.abs(var1 - var0) / var0),
        np.abs(jac1 - jac0).max() / np.abs(jac0).max()))
    assert np.abs(b1 - b0).max() <= 1e-9 * np.abs(b0).max()
    assert np.allclose(ld1, ld0, rtol=1e-12, atol=0)
    assert np.allclose(mu1, mu0, rtol=1e-7, atol=1e-8 * np.abs(mu0).max())   # k*^T beta cancels: cond(K) x 1e-16
    assert np.allclose(var1, var0, rtol=1e-7, atol=0)
    assert np.allclose(jac1, jac0, rtol=1e-8, atol=1e-10 * np.abs(jac0).max())
    ora = GPOracle(w.x_train, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in w.hyp]),
                   [h["variance"] for h in w.hyp], noise)
    mu_o, var_o, _ = ora.predict_batch(z)
    assert np.allclose(mu1, mu_o, rtol=1e-6, atol=1e-9 * np.abs(mu_o).max())
    assert np.allclose(var1, var_o, rtol=1e-6, atol=0)


def test_scratch_cache_reuses_and_frees(se):
    """segp_factorize keeps its scratch for the next same-size factorisation (option scratch_cache): same bits, and the
    memory goes back when the option is switched off."""
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=96, n_train=700)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, None, None, kern_types=w.kern_types, hyp=w.hyp, device=0)
    gp.set_option("fact_i8", 1)
    gp.train(w.x_train, w.y_train)
    beta = gp.beta.copy()
    assert gp.get_option("scratch_cached_bytes") > 0
    gp.train(w.x_train, w.y_train)                       # from the cache
    assert np.array_equal(beta, gp.beta)
    gp.set_option("scratch_cache", 0)
    assert gp.get_option("scratch_cached_bytes") == 0
    gp.train(w.x_train, w.y_train)                       # fresh allocations, freed again
    assert gp.get_option("scratch_cached_bytes") == 0 and np.array_equal(beta, gp.beta)
    gp.set_option("fact_i8", 0)
    gp.train(w.x_train, w.y_train)                       # float64 DMMA path: same factor to rounding
    assert np.abs(gp.beta - beta).max() <= 1e-10 * np.abs(beta).max()
    gp.close()
