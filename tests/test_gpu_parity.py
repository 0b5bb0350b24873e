"""GPU parity tests: the CUDA path (through the Python mirror -> ctypes -> C ABI of libsegp.so) against

* the golden vectors the reference's own functions produced (tests/golden/, oracle/make_golden.py),
* the float64 CPU oracle (oracle/) on the same seeded inputs, at sizes the oracle finishes in seconds,
* size-independent properties at the benchmark sizes.

Tolerance: BASELINE.json's rtol 1e-4 on predictive mean/variance and on (p, Q), with the reference's own
atol (1e-6 on GP outputs, test/test_gp_models.py:22-23; 1e-5 on ellipsoids, test_gp_reachability_casadi.py:25-26)
scaled to the data.  The float64 pipeline is in practice ~1e-9 or better, and the tests assert a much tighter
bound (RTOL_TIGHT) where conditioning allows so regressions show up long before the 1e-4 gate.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4          # the gate of BASELINE.json
RTOL_TIGHT = 1e-7    # what the float64 path is expected to deliver on (p, Q)


@pytest.fixture(scope="module")
def se():
    import safe_exploration_b200 as pkg
    pkg._lib.load()
    return pkg


@pytest.fixture(autouse=True, params=[(0, 0), (4, 5), (4, 0)], ids=["fp64-dmma", "int8-15products", "int8-auto"])
def _tri_mode(request, se):
    """Every parity test runs on the float64 pipe, on the tcgen05 pipe with the 15-product digit set, and on the
    tcgen05 pipe in automatic mode (the factorize-time probe picks the digit set; the precision guard recomputes
    what the 10-product set cannot resolve) -- include/segp.h, "tri_mode" / "i8_digits"."""
    old = se.ssm.DEFAULT_TRI_MODE, se.ssm.DEFAULT_I8_DIGITS
    se.ssm.DEFAULT_TRI_MODE, se.ssm.DEFAULT_I8_DIGITS = request.param
    yield request.param[0]
    se.ssm.DEFAULT_TRI_MODE, se.ssm.DEFAULT_I8_DIGITS = old


def _tight(gp, tight):
    """Regression tolerance: `tight` on the float64 / 15-product paths, the gate itself where the 10-product set runs
    (its variance error is a few 1e-6 .. 1e-5 of sigma^2 by design)."""
    return RTOL if gp.get_option("i8_digits_effective") == 4 else tight


def _make_models(se, x, y, n_s_in, n_u, kern_types, ls, var, noise_total):
    from oracle.gp_oracle import GPOracle
    # the product adds noise_diag + jitter itself; subtract them so both sides factorise the same matrix
    hyp = [{"lengthscale": ls[d], "variance": float(var[d]), "noise": float(noise_total[d]) - 1e-5 - 1e-8}
           for d in range(len(var))]
    gp = se.BatchedGPSSM(y.shape[1], n_s_in, n_u, x, y, kern_types=kern_types, hyp=hyp)
    ora = GPOracle(x, y, kern_types, ls, var, gp.total_noise())
    return gp, ora


def _gp_from_golden(se, g, n_u=1):
    kern = [str(k) for k in g["kern_types"]]
    n_s = g["y_train"].shape[1]
    return _make_models(se, g["x_train"], g["y_train"], n_s, g["x_train"].shape[1] - n_s, kern, g["lengthscale"],
                        g["variance"], g["noise"])


def _assert_close(got, want, rtol, atol_scale=1e-9, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    atol = atol_scale * max(1.0, float(np.max(np.abs(want))))
    err = np.max(np.abs(got - want) / (atol / rtol + np.abs(want))) if want.size else 0.0
    assert np.allclose(got, want, rtol=rtol, atol=atol), "{}: max scaled rel err {:.3e}".format(what, err)


# =========================================================================== golden vectors (reference outputs)
def test_golden_invpend_c1(se, golden_dir):
    """BASELINE config C1: N=50, H=5, reference fixture data, point start."""
    g = np.load(os.path.join(golden_dir, "invpend_c1.npz"))
    gp, _ = _gp_from_golden(se, g)
    c = float(g["c_safety"])
    # batched call
    p_new, q_new, p_all, q_all = se.multistep_reachability(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"],
                                                           g["l_sigma"], None, c)
    _assert_close(p_all, g["p_all"], RTOL_TIGHT, what="p_all")
    _assert_close(q_all, g["q_all"], RTOL_TIGHT, what="q_all")
    assert np.array_equal(p_new, p_all[:, -1]) and np.array_equal(q_new, q_all[:, -1])
    # the reference's un-batched call shape
    for i in range(g["k_ff"].shape[0]):
        p_n, q_n, p_a, q_a = se.multistep_reachability(g["p0"][:, None], gp, g["k_fb"], g["k_ff"][i], g["l_mu"],
                                                       g["l_sigma"], None, c)
        assert p_n.shape == (2, 1) and q_n.shape == (2, 2) and p_a.shape == (5, 2) and q_a.shape == (5, 2, 2)
        _assert_close(p_a, g["p_all"][i], RTOL_TIGHT)
        _assert_close(q_a, g["q_all"][i], RTOL_TIGHT)


def test_golden_invpend_reach_test(se, golden_dir):
    """The reference's test_gp_reachability_casadi.py recipe: onestep (both branches, with and without the
    linear prior), multistep T=3 with per-step gains, safety distance."""
    g = np.load(os.path.join(golden_dir, "invpend_reach_test.npz"))
    gp, _ = _gp_from_golden(se, g)
    c = float(g["c_safety"])
    for tag, a, b in (("lin", g["a"], g["b"]), ("nolin", None, None)):
        p1, q1 = se.onestep_reachability(g["p"], gp, g["k_ff"], g["l_mu"], g["l_sigma"], g["q"], g["k_fb"], c, 0,
                                         a=a, b=b)
        assert p1.shape == (2, 1) and q1.shape == (2, 2)
        _assert_close(p1, g["p1_set_" + tag], RTOL_TIGHT, what="p1 set " + tag)
        _assert_close(q1, g["q1_set_" + tag], RTOL_TIGHT, what="q1 set " + tag)
        p1, q1 = se.onestep_reachability(g["p"], gp, g["k_ff"], g["l_mu"], g["l_sigma"], None, g["k_fb"], c, 0,
                                         a=a, b=b)
        _assert_close(p1, g["p1_point_" + tag], RTOL_TIGHT, what="p1 point " + tag)
        _assert_close(q1, g["q1_point_" + tag], RTOL_TIGHT, what="q1 point " + tag)
    _, _, p_all, q_all = se.multistep_reachability(g["p"], gp, g["k_fb_multi"], g["k_ff_multi"], g["l_mu"],
                                                   g["l_sigma"], None, c, 0, g["a"], g["b"], None)
    _assert_close(p_all, g["p_all"], RTOL_TIGHT, what="multistep p_all")
    _assert_close(q_all, g["q_all"], RTOL_TIGHT, what="multistep q_all")
    dist = se.lin_ellipsoid_safety_distance(g["p_all"][-1][:, None], g["q_all"][-1], g["h_mat"], g["h_vec"], c)
    assert dist.shape == g["dist"].shape
    _assert_close(dist, g["dist"], 1e-10, what="safety distance")
    dist_b = se.lin_ellipsoid_safety_distance(g["p_all"], g["q_all"], g["h_mat"], g["h_vec"], c)
    _assert_close(dist_b[-1], g["dist"][:, 0], 1e-10)


def test_golden_cartpole(se, golden_dir):
    """Cart-pole fixture, mixed rbf / mat52 kernels, per-trajectory gains; with and without q_0."""
    g = np.load(os.path.join(golden_dir, "cartpole.npz"))
    gp, _ = _gp_from_golden(se, g)
    _, _, p_all, q_all = se.multistep_reachability(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"], g["l_sigma"], None,
                                                   2.0, 0, g["a"], g["b"])
    _assert_close(p_all, g["p_all"], RTOL_TIGHT, what="p_all")
    _assert_close(q_all, g["q_all"], RTOL_TIGHT, what="q_all")
    _, _, p_all, q_all = se.multistep_reachability(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"], g["l_sigma"],
                                                   g["q0"], 1.5, 0, g["a"], g["b"], g["k_fb_init"])
    _assert_close(p_all, g["p_all_q0"], RTOL_TIGHT, what="p_all q0")
    _assert_close(q_all, g["q_all_q0"], RTOL_TIGHT, what="q_all q0")


def test_golden_ellipsoid_algebra(se, golden_dir):
    """Reference outputs of compute_remainder_overapproximations (test_utils_casadi.py inputs, n_s up to 8),
    sum_two_ellipsoids and ellipsoid_from_rectangle, plus the literal cases of test_utils_ellipsoid.py."""
    from safe_exploration_b200.utils import compute_remainder_overapproximations
    from safe_exploration_b200.utils_ellipsoid import ellipsoid_from_rectangle, sum_two_ellipsoids
    g = np.load(os.path.join(golden_dir, "ellipsoid_algebra.npz"))
    for tag in ("t_1", "t_2", "t_3", "t_4"):
        u_mu, u_sig = compute_remainder_overapproximations(g[tag + "_q"], g[tag + "_k_fb"], g[tag + "_l_mu"],
                                                           g[tag + "_l_sigma"])
        _assert_close(u_mu, g[tag + "_u_mu"], 1e-11, what=tag + " u_mu")
        _assert_close(u_sig, g[tag + "_u_sigma"], 1e-11, what=tag + " u_sigma")
    for i in range(3):
        p, q = sum_two_ellipsoids(g["s%d_p1" % i], g["s%d_q1" % i], g["s%d_p2" % i], g["s%d_q2" % i])
        assert p.shape == g["s%d_p" % i].shape
        _assert_close(p, g["s%d_p" % i], 1e-13)
        _assert_close(q, g["s%d_q" % i], 1e-13)
        _assert_close(ellipsoid_from_rectangle(g["s%d_ub" % i]), g["s%d_qrect" % i], 1e-14)
    for tag in ("rectangle", "cube"):
        _assert_close(ellipsoid_from_rectangle(g["rect_" + tag + "_ub"]), g["rect_" + tag + "_q"], 1e-14)
    with pytest.raises(AssertionError):          # reference test/test_utils_ellipsoid.py:28-33
        ellipsoid_from_rectangle([0.6, -0.3])
    # batched leaves
    qs = np.stack([g["t_2_q"], 2.0 * g["t_2_q"]])
    um, us = compute_remainder_overapproximations(qs, g["t_2_k_fb"], g["t_2_l_mu"], g["t_2_l_sigma"])
    _assert_close(um[0], g["t_2_u_mu"], 1e-11)
    _assert_close(um[1], 2.0 * g["t_2_u_mu"], 1e-11)
    _assert_close(us[1], np.sqrt(2.0) * g["t_2_u_sigma"], 1e-11)


# =========================================================================== GP posterior vs the oracle
@pytest.mark.parametrize("n,n_s,n_u,kerns,t", [
    (50, 2, 1, ["rbf", "mat52"], 1),
    (128, 3, 2, ["mat52", "rbf", "rbf"], 127),
    (300, 4, 1, ["rbf", "mat52", "mat52", "rbf"], 129),
    (700, 2, 1, ["rbf", "rbf"], 1000),
])
def test_predict_matches_oracle(se, n, n_s, n_u, kerns, t):
    rng = np.random.RandomState(n + t)
    dim = n_s + n_u
    x = rng.uniform(-1, 1, size=(n, dim))
    y = np.tanh(x @ rng.randn(dim, n_s)) + 0.05 * rng.randn(n, n_s)
    ls = rng.uniform(0.7, 2.0, size=(n_s, dim))
    var = rng.uniform(0.5, 1.5, size=n_s)
    noise = rng.uniform(0.01, 0.05, size=n_s)
    gp, ora = _make_models(se, x, y, n_s, n_u, kerns, ls, var, noise)
    z = rng.uniform(-1, 1, size=(t, dim))
    mu, sig2, jac = gp.predict(z, compute_gradients=True)
    mu_o, var_o, jac_o = ora.predict_batch(z)
    assert mu.shape == (t, n_s) and sig2.shape == (t, n_s) and jac.shape == (t, n_s, dim)
    _assert_close(mu, mu_o, RTOL, atol_scale=1e-6, what="mean")       # the gate
    _assert_close(sig2, var_o, RTOL, atol_scale=1e-6, what="variance")
    _assert_close(mu, mu_o, 1e-6, atol_scale=1e-9, what="mean (tight)")
    _assert_close(sig2, var_o, 1e-6, atol_scale=1e-10, what="variance (tight)")
    _assert_close(jac, jac_o, 1e-6, atol_scale=1e-8, what="jacobian")
    # ABC spelling and the single-point __call__ triple (gp_reachability.py:74,101)
    mu2, var2 = gp.predict(z[:, :n_s], z[:, n_s:])
    assert np.array_equal(mu2, mu) and np.array_equal(var2, sig2)
    m1, v1, j1 = gp(z[:1, :n_s], z[:1, n_s:])
    assert m1.shape == (n_s, 1) and v1.shape == (n_s, 1) and j1.shape == (n_s, dim)
    _assert_close(m1[:, 0], mu_o[0], 1e-6, atol_scale=1e-9)
    # beta and log-determinant of the factorisation
    _assert_close(gp.beta, ora.beta, 1e-6, atol_scale=1e-7, what="beta")
    logdet_o = np.array([2.0 * np.sum(np.log(np.diag(l))) for l in ora.chol])
    _assert_close(gp.log_det_k(), logdet_o, 1e-10, what="logdet")
    gp.close()


def test_predict_at_training_inputs_identity(se):
    """Size-independent property that checks potrf + trtri without an oracle: at a training input,
    mean_i = y_i - noise * beta_i (K beta = y - noise beta)."""
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=1)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    idx = np.arange(0, w.n_train, 7)
    mu, var = gp.predict(w.x_train[idx])
    want = w.y_train[idx] - gp.total_noise()[None, :] * gp.beta[idx]
    _assert_close(mu, want, 1e-6, atol_scale=1e-8, what="mean at training inputs")
    assert np.all(var > 0) and np.all(var < gp.total_noise()[None, :])
    gp.close()


def test_model_beyond_the_int8_size_limit_runs_in_float64(se, _tri_mode):
    """More than 16384 (padded) training points: the int32 accumulators of the digit-plane products would overflow,
    so the model must select the float64 contraction on its own, refuse a tcgen05 mode loudly, and still satisfy the
    training-input identity mean_i = y_i - noise beta_i (checks potrf / trtri / beta / K* / mean at that size)."""
    if _tri_mode != 0:
        pytest.skip("one run is enough: the model picks its pipe itself")
    rng = np.random.default_rng(21)
    n, dim = 16500, 3
    x = rng.uniform(-1, 1, (n, dim))
    y = np.sin(x @ rng.standard_normal((dim, 1))) + 0.05 * rng.standard_normal((n, 1))
    hyp = [{"lengthscale": np.array([0.7, 0.9, 1.1]), "variance": 1.0, "noise": 5e-2}]
    gp = se.BatchedGPSSM(1, 2, 1, x, y, kern_types=["rbf"], hyp=hyp, tri_mode=-1)
    assert gp.get_option("n_train_padded") > 16384 and gp.get_option("tri_mode_effective") == 0
    with pytest.raises(NotImplementedError):
        gp.set_option("tri_mode", 4)
    idx = np.arange(0, n, 157)
    mu, var = gp.predict(x[idx])
    want = y[idx] - gp.total_noise()[None, :] * gp.beta[idx]
    _assert_close(mu, want, 1e-6, atol_scale=1e-8, what="mean at training inputs")
    assert np.all(var > 0) and np.all(var < gp.total_noise()[None, :])
    gp.close()


@pytest.mark.parametrize("kerns", [["rbf", "mat52"], ["lin_mat52", "rbf"]])
def test_append_data_equals_refactorisation(se, kerns):
    """update_model(replace_old=False) -> segp_append: appending points (1, a few, across a 64-row block, across the
    128-row padding) must give the model a from-scratch factorisation of the concatenated data gives -- predictions,
    Jacobians, beta, log-determinant, rollouts -- and take the incremental path whenever the padded size stays."""
    from safe_exploration_b200 import workloads
    rng = np.random.default_rng(33)
    n_s, n_u = 2, 1
    dim = n_s + n_u
    n_all = 420
    x = rng.uniform(-1, 1, (n_all, dim))
    y = np.sin(x @ rng.standard_normal((dim, n_s))) + 0.05 * rng.standard_normal((n_all, n_s))
    hyp = []
    for k in kerns:
        if k.startswith("lin_"):
            hyp.append({"prod.mat52.lengthscale": np.array([0.9]), "prod.mat52.variance": 1.1,
                        "prod.linear.variances": np.array([0.6]), "linear.variances": rng.uniform(0.1, 0.4, dim),
                        "noise": 1e-2})
        else:
            hyp.append({"lengthscale": rng.uniform(0.7, 1.5, dim), "variance": 0.9, "noise": 2e-2})
    z = rng.uniform(-1, 1, (150, dim))
    gp = se.BatchedGPSSM(n_s, n_s, n_u, x[:290], y[:290], kern_types=kerns, hyp=hyp)
    n = 290
    # (points to append, incremental expected): the first call switches dense-W keeping on and refactorises
    for add, incremental in ((1, False), (1, True), (5, True), (60, True), (27, True), (1, False), (35, True)):
        gp.update_model(x[n:n + add], y[n:n + add], replace_old=False)
        n += add
        assert gp.get_option("n_train") == n and gp.x_train.shape == (n, dim)
        assert bool(gp.get_option("append_incremental")) == incremental, (n, add)
        ref = se.BatchedGPSSM(n_s, n_s, n_u, x[:n], y[:n], kern_types=kerns, hyp=hyp)
        for a, b, what in zip(gp.predict(z, compute_gradients=True), ref.predict(z, compute_gradients=True),
                              ("mean", "variance", "jacobian")):
            _assert_close(a, b, 1e-7, atol_scale=1e-9, what="%s after appending to %d" % (what, n))
        _assert_close(gp.beta, ref.beta, 1e-6, atol_scale=1e-8, what="beta")
        _assert_close(gp.log_det_k(), ref.log_det_k(), 1e-10, what="log det")
        ref.close()
    assert n == n_all
    # the rollouts run on the refreshed packed operands (both tensor pipes via the fixture)
    w = workloads.make("C2", batch=64, n_train=50, horizon=3)
    ref = se.BatchedGPSSM(n_s, n_s, n_u, x, y, kern_types=kerns, hyp=hyp)
    args = (w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    r1 = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    r0 = se.rollout(ref, w.p0, w.k_ff, w.k_fb, *args)
    _assert_close(r1.q_all, r0.q_all, 1e-6, what="q_all after appends")
    _assert_close(r1.var_all, r0.var_all, 1e-6, atol_scale=1e-10, what="var_all after appends")
    gp.close()
    ref.close()


def test_subset_of_data_and_sampling(se):
    """SimpleGPModel's subset-of-data mode (ssm_gpy/gaussian_process.py:201-222, 372-392): with m given, the GP
    conditions on m points chosen by the max-variance criterion (or at random) and keeps the whole set as x_train;
    sample_from_gp draws from the independent predictive distributions (:598-619)."""
    from oracle import gp_oracle, select_oracle
    rng = np.random.default_rng(17)
    n, n_s, n_u, m = 260, 2, 1, 90
    dim = n_s + n_u
    x = rng.uniform(-1, 1, (n, dim))
    y = np.sin(x @ rng.standard_normal((dim, n_s))) + 0.05 * rng.standard_normal((n, n_s))
    hyp = [{"lengthscale": rng.uniform(0.6, 1.4, dim), "variance": 1.0, "noise": 2e-2} for _ in range(n_s)]
    gp = se.BatchedGPSSM(n_s, n_s, n_u, x, y, m=m, kern_types=["rbf", "mat52"], hyp=hyp)
    ls, var, _, _ = gp_oracle.vectors_from_reference_hyp(["rbf", "mat52"], hyp, dim)
    idx, _ = select_oracle.greedy_maxvar(x, ["rbf", "mat52"], ls, var, gp.total_noise(), m)
    assert gp.x_train.shape == (n, dim) and gp.z.shape == (m, dim) and np.array_equal(gp.z, x[idx])
    ref = se.BatchedGPSSM(n_s, n_s, n_u, x[idx], y[idx], kern_types=["rbf", "mat52"], hyp=hyp)
    z = rng.uniform(-1, 1, (40, dim))
    for a, b in zip(gp.predict(z), ref.predict(z)):
        assert np.array_equal(a, b)
    assert gp.beta.shape == (m, n_s) and len(gp.information_gain()) == n_s
    # appended data: the subset is re-selected from the whole set
    gp.update_model(x[:30] + 0.3, y[:30], replace_old=False)
    assert gp.x_train.shape == (n + 30, dim) and gp.z.shape == (m, dim)
    # random subset (choose_data=False): m rows of the data, reproducible through NumPy's global seed
    np.random.seed(3)
    gp.train(x, y, choose_data=False)
    np.random.seed(3)
    pick = np.random.choice(n, size=m, replace=False)
    assert np.array_equal(gp.z, x[pick])
    with pytest.warns(UserWarning):
        gp.train(x[:50], y[:50])                      # fewer points than m: all of them, with the reference's warning
    assert gp.z.shape == (50, dim)
    np.random.seed(0)
    smp = gp.sample_from_gp(z, size=4000)
    mu, var_p = gp.predict(z)
    assert smp.shape == (40, 4000, n_s)
    assert np.allclose(smp.mean(axis=1), mu, atol=5 * np.sqrt(var_p.max() / 4000))
    assert np.allclose(smp.var(axis=1), var_p, rtol=0.15)
    gp.close()
    ref.close()


def test_empty_and_single_candidate_batches(se):
    """Ragged ends of the batch axis: no candidates at all, one candidate, one more than a 96-trajectory panel."""
    from safe_exploration_b200 import workloads
    w = workloads.make("C2", batch=97, n_train=130, horizon=3)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    args = (w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    full = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    none = se.rollout(gp, w.p0, w.k_ff[:0], w.k_fb, *args)
    assert none.p_all.shape == (0, 3, w.n_s) and none.q_all.shape == (0, 3, w.n_s, w.n_s) and none.status.shape == (0,)
    one = se.rollout(gp, w.p0, w.k_ff[96:97], w.k_fb, *args)
    assert np.array_equal(one.q_all[0], full.q_all[96]) and np.array_equal(one.p_all[0], full.p_all[96])
    mu0, var0 = gp.predict(np.zeros((0, w.n_s + w.n_u)))
    assert mu0.shape == (0, w.n_s) and var0.shape == (0, w.n_s)
    gp.close()


@pytest.mark.parametrize("name", ["pend_rbf_mat52", "pend_composite", "cart_mixed"])
def test_golden_gp_pred_reference(se, golden_dir, name):
    """Predictive mean / variance the reference's OWN gp_models_utils_casadi.py functions produced (kernels incl. the
    composite lin_rbf / lin_mat52 ones, gp_pred; oracle/make_golden.golden_gp_pred) against segp_predict, and the
    mean Jacobian against the oracle's closed form."""
    from oracle import gp_oracle
    g = np.load(os.path.join(golden_dir, "gp_pred_reference.npz"))
    ora, kerns, hyp = gp_oracle.golden_gp_case(g, name)
    x, y, z = g[name + "/x_train"], g[name + "/y_train"], g[name + "/z"]
    n_s = y.shape[1]
    hyp = [dict(h, noise=float(g[name + "/noise"][d]) - 1e-5 - 1e-8) for d, h in enumerate(hyp)]
    gp = se.BatchedGPSSM(n_s, n_s, x.shape[1] - n_s, x, y, kern_types=kerns, hyp=hyp)
    assert np.allclose(gp.total_noise(), g[name + "/noise"], rtol=1e-13)
    mu, var, jac = gp.predict(z, compute_gradients=True)
    _assert_close(mu, g[name + "/mu"], RTOL, atol_scale=1e-6, what="mean (gate)")
    _assert_close(var, g[name + "/var"], RTOL, atol_scale=1e-6, what="variance (gate)")
    _assert_close(mu, g[name + "/mu"], 1e-7, what="mean")
    _assert_close(var, g[name + "/var"], 1e-6, atol_scale=1e-9, what="variance")
    _assert_close(jac, ora.jacobian(z), 1e-7, what="jacobian")
    if any(k.startswith("lin_") for k in kerns):
        # composite kernels run on whichever pipe the fixture selects (per-trajectory scale on the int8 path)
        assert gp.get_option("tri_mode_effective") == se.ssm.DEFAULT_TRI_MODE
    gp.close()


def test_rollout_composite_kernels_vs_oracle(se):
    """H-step reachability with the journal configs' kernel structure (lin_mat52 / lin_rbf mixed with plain ones),
    both definitions of the composite kernels (CasADi: product term on input column 1; GPy object: all columns),
    against the batch oracle; the two definitions must differ."""
    from oracle import gp_oracle, reach_oracle
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=150, n_train=300, horizon=4)
    dim = w.n_s + w.n_u
    rng = np.random.RandomState(4)
    kerns = ["lin_mat52", "rbf", "lin_rbf", "mat52"]
    hyp = []
    for d, k in enumerate(kerns):
        if k.startswith("lin_"):
            st = k[4:]
            hyp.append({"prod.%s.lengthscale" % st: np.array([rng.uniform(0.8, 1.6)]), "prod.%s.variance" % st: 0.9,
                        "prod.linear.variances": np.array([rng.uniform(0.4, 1.0)]),
                        "linear.variances": rng.uniform(0.05, 0.3, dim), "noise": 1e-2})
        else:
            hyp.append(dict(w.hyp[d]))
    results = []
    for sem in ("casadi", "gpy"):
        gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=kerns, hyp=hyp,
                             composite_semantics=sem)
        ls, var, pl, lin = gp_oracle.vectors_from_reference_hyp(kerns, hyp, dim, semantics=sem)
        ora = gp_oracle.GPOracle(w.x_train, w.y_train, kerns, ls, var, gp.total_noise(), prod_linear=pl, linear=lin)
        res = se.rollout(gp, w.p0, w.k_ff, w.k_fb, w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
        p_o, q_o, v_o = reach_oracle.multistep_batch(w.p0, ora, w.k_fb, w.k_ff, w.l_mu, w.l_sigma, None, w.c_safety,
                                                     w.a, w.b)
        assert np.all(res.status == 0) and np.all(np.isfinite(q_o))
        _assert_close(res.var_all, v_o, 1e-6, atol_scale=1e-10, what="variance")
        _assert_close(res.p_all, p_o, 1e-7, what="p_all")
        _assert_close(res.q_all, q_o, 1e-6, what="q_all")
        results.append(res.p_all)
        gp.close()
    assert np.max(np.abs(results[0] - results[1])) > 1e-6


def test_select_maxvar_and_information_gain(se):
    """SURVEY 8 f4 on the device: greedy max-predicted-variance selection against the oracle (same index sequence,
    same scores), choose_datapoints_maxvar's surface, and information_gain for the training set and for foreign
    inputs against numpy's slogdet."""
    from oracle import gp_oracle, select_oracle
    rng = np.random.default_rng(9)
    n, n_s, n_u = 700, 3, 1
    dim = n_s + n_u
    x = rng.uniform(-1, 1, (n, dim))
    y = rng.standard_normal((n, n_s))
    kerns = ["rbf", "lin_mat52", "mat52"]
    hyp = [{"lengthscale": rng.uniform(0.6, 1.5, dim), "variance": 1.1, "noise": 2e-2},
           {"prod.mat52.lengthscale": np.array([0.8]), "prod.mat52.variance": 0.9,
            "prod.linear.variances": np.array([0.7]), "linear.variances": rng.uniform(0.05, 0.3, dim), "noise": 1e-2},
           {"lengthscale": rng.uniform(0.6, 1.5, dim), "variance": 0.6, "noise": 3e-2}]
    gp = se.BatchedGPSSM(n_s, n_s, n_u, x[:200], y[:200], kern_types=kerns, hyp=hyp)
    ls, var, pl, lin = gp_oracle.vectors_from_reference_hyp(kerns, hyp, dim)
    m = 120
    idx_o, score_o = select_oracle.greedy_maxvar(x, kerns, ls, var, gp.total_noise(), m, pl, lin)
    idx, score = gp.select_maxvar(x, m)
    assert np.array_equal(idx, idx_o)
    _assert_close(score, score_o, 1e-9, what="selection scores")
    xc, yc = gp.choose_datapoints_maxvar(x, y, m)
    assert np.array_equal(xc, x[idx_o]) and np.array_equal(yc, y[idx_o])
    xa, ya = gp.choose_datapoints_maxvar(x[:50], y[:50], 60)           # fewer points than m: everything
    assert xa.shape == (50, dim) and ya.shape == (50, n_s)
    with pytest.raises(ValueError):
        gp.select_maxvar(x, n + 1)

    def ig_ref(xx):
        ora = gp_oracle.GPOracle(xx, np.zeros((xx.shape[0], n_s)), kerns, ls, var, gp.total_noise(), pl, lin)
        return np.array([2.0 * np.sum(np.log(np.diag(l))) - xx.shape[0] * np.log(s)
                         for l, s in zip(ora.chol, gp.total_noise())])

    _assert_close(gp.information_gain(), ig_ref(x[:200]), 1e-9, what="information gain (training set)")
    _assert_close(gp.information_gain(x[300:520]), ig_ref(x[300:520]), 1e-9, what="information gain (foreign x)")
    gp.close()


@pytest.mark.parametrize("kerns", [["rbf", "mat52"], ["lin_rbf", "mat52"]])
def test_factor_buffer_transplant_equals_own_factorisation(se, kerns):
    """The multi-GPU setup path on one device: a model that only received the data (set_data_only) plus a copy of
    another model's factor buffers (what the NCCL broadcast delivers) and mark_factorized predicts bit-identically
    to the model that factorised -- including the composite kernels, whose X^T beta term is rebuilt from the
    transplanted beta."""
    from safe_exploration_b200.ssm import _tensor_from_ptr
    import torch
    rng = np.random.default_rng(12)
    n, n_s, n_u = 300, 2, 1
    dim = n_s + n_u
    x = rng.uniform(-1, 1, (n, dim))
    y = rng.standard_normal((n, n_s))
    hyp = []
    for k in kerns:
        if k.startswith("lin_"):
            hyp.append({"prod.rbf.lengthscale": np.array([0.9]), "prod.rbf.variance": 1.1,
                        "prod.linear.variances": np.array([0.6]), "linear.variances": rng.uniform(0.1, 0.4, dim),
                        "noise": 1e-2})
        else:
            hyp.append({"lengthscale": rng.uniform(0.7, 1.5, dim), "variance": 0.9, "noise": 2e-2})
    gp1 = se.BatchedGPSSM(n_s, n_s, n_u, x, y, kern_types=kerns, hyp=hyp)
    gp2 = se.BatchedGPSSM(n_s, n_s, n_u, kern_types=kerns, hyp=hyp)
    gp2.set_data_only(x, y)
    with pytest.raises(RuntimeError):
        gp2.predict(x[:3])
    b1 = gp1.factor_buffers()
    if len(b1) == 2:        # float64 contraction (tri_mode 0 / composite kernels): the DMMA operand travels too
        gp2.alloc_fp64_operand()
    b2 = gp2.factor_buffers()
    assert [nb for _, nb in b1] == [nb for _, nb in b2]
    for (p1, nb), (p2, _) in zip(b1, b2):
        _tensor_from_ptr(torch, p2, nb, gp2.device).copy_(_tensor_from_ptr(torch, p1, nb, gp1.device))
    torch.cuda.synchronize()
    gp2.mark_factorized()
    z = rng.uniform(-1, 1, (200, dim))
    for a, b in zip(gp1.predict(z, compute_gradients=True), gp2.predict(z, compute_gradients=True)):
        assert np.array_equal(a, b)
    assert np.array_equal(gp1.beta, gp2.beta)
    gp1.close()
    gp2.close()


# =========================================================================== rollouts vs the batch oracle
def _rollout_vs_oracle(se, w, t_z_gp=None, q0=None, k_fb_init=None, per_traj_kfb=False, rtol=RTOL_TIGHT):
    from oracle import reach_oracle
    from oracle.gp_oracle import GPOracle
    n_in = w.n_s if t_z_gp is None else t_z_gp.shape[0]
    x = w.x_train if t_z_gp is None else np.hstack((w.x_train[:, :w.n_s] @ t_z_gp.T, w.x_train[:, w.n_s:]))
    hyp = w.hyp if t_z_gp is None else [
        {"lengthscale": np.concatenate((h["lengthscale"][:n_in], h["lengthscale"][w.n_s:])), "variance": h["variance"],
         "noise": h["noise"]} for h in w.hyp]
    gp = se.BatchedGPSSM(w.n_s, n_in, w.n_u, x, w.y_train, kern_types=w.kern_types, hyp=hyp)
    ora = GPOracle(x, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in hyp]),
                   [h["variance"] for h in hyp], gp.total_noise())
    k_fb = w.k_fb
    if per_traj_kfb:
        rng = np.random.RandomState(5)
        k_fb = w.k_fb[None] * (1.0 + 0.05 * rng.randn(w.batch, 1, 1, 1))
    res = se.rollout(gp, w.p0, w.k_ff, k_fb, w.l_mu, w.l_sigma, q0, k_fb_init, w.c_safety, w.a, w.b, t_z_gp)
    p_o, q_o, v_o = reach_oracle.multistep_batch(w.p0, ora, k_fb, w.k_ff, w.l_mu, w.l_sigma, q0, w.c_safety, w.a, w.b,
                                                 k_fb_init, t_z_gp)
    assert np.all(res.status == 0)
    assert np.all(np.isfinite(q_o))
    _assert_close(res.var_all, v_o, RTOL, atol_scale=1e-6, what="variance (gate)")
    _assert_close(res.p_all, p_o, RTOL, atol_scale=1e-5, what="p_all (gate)")
    _assert_close(res.q_all, q_o, RTOL, atol_scale=1e-5, what="q_all (gate)")
    # regression bounds far inside the gate; where the 10-product digit set runs (variance error a few 1e-6 of sigma^2
    # by design, amplified on the small entries of Q over the horizon) the bound is the gate with SURVEY 8d's atol
    ten = gp.get_option("i8_digits_effective") == 4
    _assert_close(res.var_all, v_o, _tight(gp, 1e-6), atol_scale=1e-10, what="variance (tight)")
    _assert_close(res.p_all, p_o, _tight(gp, rtol), what="p_all (tight)")
    _assert_close(res.q_all, q_o, _tight(gp, rtol), atol_scale=1e-6 if ten else 1e-9, what="q_all (tight)")
    return gp, res


def test_rollout_c2_full_size(se):
    """BASELINE config C2 (pendulum, N=500, H=10) on 512 of its candidates, every step of every trajectory."""
    from safe_exploration_b200 import workloads
    gp, _ = _rollout_vs_oracle(se, workloads.make("C2", batch=512))
    gp.close()


def test_rollout_c3_model_size(se):
    """BASELINE config C3 (cart-pole, Matern-5/2, N=2000, H=15) on 192 candidates (ragged vs the 128 tile)."""
    from safe_exploration_b200 import workloads
    gp, _ = _rollout_vs_oracle(se, workloads.make("C3", batch=192), rtol=1e-6)
    gp.close()


def test_rollout_c5_shape_reduced(se):
    """The 10-D / 3-action shape of C5 (generic n_s path of the ellipsoid kernel) at N=600, H=6."""
    from safe_exploration_b200 import workloads
    gp, _ = _rollout_vs_oracle(se, workloads.make("C5", batch=70, n_train=600, horizon=6), rtol=1e-6)
    gp.close()


def test_rollout_with_q0_per_trajectory_gains_and_input_transform(se):
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=33, n_train=400, horizon=5)
    q0 = 1e-3 * np.array([[2., .3, 0., .1], [.3, 1., .2, 0.], [0., .2, 1.5, .4], [.1, 0., .4, 1.]])
    gp, _ = _rollout_vs_oracle(se, w, q0=q0, k_fb_init=w.k_fb[0], per_traj_kfb=True, rtol=1e-6)
    gp.close()
    # GP sees only (vel, theta, omega): t_z_gp drops the cart position (reference defaultconfig_episode.py:44)
    t = np.eye(4)[1:]
    gp, _ = _rollout_vs_oracle(se, w, t_z_gp=t, rtol=1e-6)
    gp.close()


def test_rollout_is_deterministic_chunk_and_order_invariant(se):
    """Bit-exact properties: same call twice; chunk size 128 vs default; reversed candidate order; device
    tensors vs host buffers."""
    import torch
    from safe_exploration_b200 import workloads
    w = workloads.make("C2", batch=300)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    args = (w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    r1 = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    r2 = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    assert np.array_equal(r1.q_all, r2.q_all) and np.array_equal(r1.p_all, r2.p_all)
    gp.set_option("chunk", 128)
    r3 = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    assert np.array_equal(r1.q_all, r3.q_all) and np.array_equal(r1.p_all, r3.p_all)
    gp.set_option("chunk", 8192)
    r4 = se.rollout(gp, w.p0, w.k_ff[::-1].copy(), w.k_fb, *args)
    assert np.array_equal(r1.q_all, r4.q_all[::-1]) and np.array_equal(r1.var_all, r4.var_all[::-1])
    dev = gp.device
    r5 = se.rollout(gp, torch.as_tensor(w.p0, device=dev), torch.as_tensor(w.k_ff, device=dev),
                    torch.as_tensor(w.k_fb, device=dev), *args)
    assert np.array_equal(r5.q_all.cpu().numpy(), r1.q_all) and np.array_equal(r5.p_all.cpu().numpy(), r1.p_all)
    assert int(r5.status.abs().sum().item()) == 0
    assert gp.get_option("launches") > 0
    # caller-owned (page-locked) result buffers
    out = se.pinned_result(gp, w.k_ff.shape[0], w.k_ff.shape[1])
    r6 = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args, out=out)
    assert r6.q_all is out.q_all and np.array_equal(r6.q_all, r1.q_all) and np.array_equal(r6.var_all, r1.var_all)
    assert np.array_equal(r6.status, r1.status)
    with pytest.raises(ValueError):
        se.rollout(gp, w.p0, w.k_ff[:10], w.k_fb, *args, out=out)
    gp.close()


def test_multistep_first_step_equals_onestep_and_horizon_prefix(se):
    """Structural property at any size: the first H' steps of an H-step rollout equal an H'-step rollout, and
    step t+1 equals onestep_reachability applied to step t's ellipsoid."""
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=40, n_train=500, horizon=6)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    _, _, p_all, q_all = se.multistep_reachability(w.p0, gp, w.k_fb, w.k_ff, w.l_mu, w.l_sigma, None, w.c_safety, 0,
                                                   w.a, w.b)
    _, _, p3, q3 = se.multistep_reachability(w.p0, gp, w.k_fb[:2], w.k_ff[:, :3], w.l_mu, w.l_sigma, None,
                                             w.c_safety, 0, w.a, w.b)
    assert np.array_equal(p3, p_all[:, :3]) and np.array_equal(q3, q_all[:, :3])
    p1, q1 = se.onestep_reachability(p_all[:, 2], gp, w.k_ff[:, 3], w.l_mu, w.l_sigma, q_all[:, 2], w.k_fb[2],
                                     w.c_safety, 0, w.a, w.b)
    assert np.array_equal(p1, p_all[:, 3]) and np.array_equal(q1, q_all[:, 3])
    gp.close()


def test_foreign_ssm_uses_ellipsoid_step(se):
    """Any callable with the reference's plugin signature works: here the CPU oracle GP is the 'foreign' model and
    only the ellipsoid algebra runs on the GPU."""
    from oracle import reach_oracle
    from oracle.gp_oracle import GPOracle
    rng = np.random.RandomState(3)
    n_s, n_u, n = 3, 2, 40
    x = rng.uniform(-1, 1, size=(n, n_s + n_u))
    y = np.tanh(x @ rng.randn(n_s + n_u, n_s))
    ora = GPOracle(x, y, ["rbf", "mat52", "rbf"], rng.uniform(0.7, 2, size=(n_s, n_s + n_u)),
                   rng.uniform(0.5, 1.5, size=n_s), rng.uniform(0.01, 0.05, size=n_s))
    m = rng.randn(n_s, n_s)
    q = 0.05 * (m @ m.T + 0.1 * np.eye(n_s))
    p = 0.1 * rng.randn(n_s, 1)
    k_fb = 0.5 * rng.randn(n_u, n_s)
    k_ff = 0.2 * rng.randn(n_u, 1)
    l_mu = rng.uniform(1e-3, 5e-2, n_s)
    l_sig = rng.uniform(1e-3, 5e-2, n_s)
    a = np.eye(n_s) + 0.1 * rng.randn(n_s, n_s)
    b = rng.randn(n_s, n_u)
    for qq in (q, None):
        p1, q1 = se.onestep_reachability(p, ora, k_ff, l_mu, l_sig, qq, k_fb, 1.7, 0, a, b)
        po, qo = reach_oracle.onestep_reachability(p, ora, k_ff, l_mu, l_sig, qq, k_fb, 1.7, 0, a, b)
        _assert_close(p1, po, 1e-12)
        _assert_close(q1, qo, 1e-11)
    hor = 4
    kfb = 0.5 * rng.randn(hor - 1, n_u, n_s)
    kff = 0.2 * rng.randn(hor, n_u)
    _, _, pa, qa = se.multistep_reachability(p, ora, kfb, kff, l_mu, l_sig, None, 2.0, 0, a, b, None)
    _, _, pa_o, qa_o = reach_oracle.multistep_reachability(p, ora, kfb, kff, l_mu, l_sig, None, 2.0, 0, a, b, None)
    _assert_close(pa, pa_o, 1e-10)
    _assert_close(qa, qa_o, 1e-10)


@pytest.mark.parametrize("n_s,n_u", [(5, 2), (7, 3), (8, 1), (10, 3), (16, 8)])
def test_generic_step_kernel_sizes(se, n_s, n_u):
    """State dimensions beyond the register-resident instances (n_s > 4) take the generic step kernel, whose
    lambda_max(Q (I + K^T K)) is a Jacobi iteration spread over the block's warps (round-robin ordering; odd n_s plays
    with a dummy): multi-step rollouts of a foreign model against the oracle, batched so that several lanes iterate on
    different matrices."""
    from oracle import reach_oracle
    from oracle.gp_oracle import GPOracle
    rng = np.random.RandomState(10 * n_s + n_u)
    n = 30
    x = rng.uniform(-1, 1, size=(n, n_s + n_u))
    y = np.tanh(x @ rng.randn(n_s + n_u, n_s))
    ora = GPOracle(x, y, ["rbf"] * n_s, rng.uniform(0.7, 2, size=(n_s, n_s + n_u)),
                   rng.uniform(0.5, 1.5, size=n_s), rng.uniform(0.01, 0.05, size=n_s))
    a = 0.7 * np.eye(n_s) + 0.05 * rng.randn(n_s, n_s)
    b = 0.3 * rng.randn(n_s, n_u)
    l_mu = rng.uniform(1e-3, 1e-2, n_s)
    l_sig = rng.uniform(1e-3, 1e-2, n_s)
    hor = 4
    kfb = 0.3 * rng.randn(hor - 1, n_u, n_s)
    for trial in range(3):
        p = 0.1 * rng.randn(n_s, 1)
        kff = 0.2 * rng.randn(hor, n_u)
        _, _, pa, qa = se.multistep_reachability(p, ora, kfb, kff, l_mu, l_sig, None, 2.0, 0, a, b, None)
        _, _, pa_o, qa_o = reach_oracle.multistep_reachability(p, ora, kfb, kff, l_mu, l_sig, None, 2.0, 0, a, b, None)
        _assert_close(pa, pa_o, 1e-10)
        _assert_close(qa, qa_o, 1e-9)
    # one step from a given ellipsoid, including a rank-deficient shape matrix (zero rotations must be no-ops)
    m = rng.randn(n_s, n_s)
    for q in (0.05 * (m @ m.T + 0.1 * np.eye(n_s)), 0.05 * np.outer(m[0], m[0]) + 1e-9 * np.eye(n_s)):
        p = 0.1 * rng.randn(n_s, 1)
        k1 = 0.5 * rng.randn(n_u, n_s)
        kf = 0.2 * rng.randn(n_u, 1)
        p1, q1 = se.onestep_reachability(p, ora, kf, l_mu, l_sig, q, k1, 1.7, 0, a, b)
        po, qo = reach_oracle.onestep_reachability(p, ora, kf, l_mu, l_sig, q, k1, 1.7, 0, a, b)
        _assert_close(p1, po, 1e-12)
        _assert_close(q1, qo, 1e-10)


# =========================================================================== error behaviour
def test_error_mapping_and_status_flags(se):
    from safe_exploration_b200 import workloads
    rng = np.random.RandomState(0)
    x = rng.uniform(-1, 1, size=(30, 3))
    y = rng.randn(30, 2)
    # a non-finite kernel matrix has no Cholesky factor -> LinAlgError, as numpy/LAPACK raise in the reference
    xd = x.copy()
    xd[7, 1] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        se.BatchedGPSSM(2, 2, 1, xd, y)
    with pytest.raises(ValueError):          # unknown kernel names, as gaussian_process.py:476-478
        se.BatchedGPSSM(2, 2, 1, x, y, kern_types=["lin", "rbf"])
    with pytest.raises(KeyError):            # composite kernels need the reference's prod.* / linear.* hyper-parameters
        se.BatchedGPSSM(2, 2, 1, x, y, kern_types=["lin_rbf", "rbf"], hyp=[{"lengthscale": 1.0}, {"lengthscale": 1.0}])
    with pytest.raises(ValueError):
        se.BatchedGPSSM(2, 2, 1, x, y, kern_types=["rbf", "nope"])
    gp = se.BatchedGPSSM(2, 2, 1)
    with pytest.raises(RuntimeError):
        gp.predict(x)
    with pytest.raises(NotImplementedError):
        gp.update_model(x, y, opt_hyp=True)
    gp.update_model(x, y)
    gp.update_model(x[:5] + 0.3, y[:5], replace_old=False)
    assert gp.x_train.shape[0] == 35
    # zero Lipschitz constant -> the reference's ellipsoid_from_rectangle assertion; batch mode flags it instead
    w = workloads.make("C2", batch=5, n_train=60, horizon=3)
    gp2 = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    res = se.rollout(gp2, w.p0, w.k_ff, w.k_fb, np.zeros(2), w.l_sigma, None, None, w.c_safety, w.a, w.b)
    assert np.all(res.status & se._lib.STATUS_ZERO_BOUND)
    with pytest.raises(AssertionError):
        se.multistep_reachability(w.p0[:, None], gp2, w.k_fb, w.k_ff[0], np.zeros(2), w.l_sigma, None, w.c_safety)
    with pytest.raises(ValueError):
        se.rollout(gp2, w.p0, w.k_ff[:, :, :0], w.k_fb, w.l_mu, w.l_sigma)
    # empty batch is a no-op
    res = se.rollout(gp2, w.p0, w.k_ff[:0], w.k_fb, w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    assert res.p_all.shape == (0, 3, 2)
    gp.close()
    gp2.close()


# =========================================================================== Gaussian uncertainty propagation (8 f2)
def test_golden_uncertainty_propagation(se, golden_dir):
    """multi_step_taylor_symbolic / mean_equivalent_multistep against outputs of the reference's own functions."""
    from safe_exploration_b200 import uncertainty_propagation as up
    g = np.load(os.path.join(golden_dir, "uncertainty_propagation.npz"))
    kern = [str(k) for k in g["kern_types"]]
    n_s = g["y_train"].shape[1]
    n_u = g["x_train"].shape[1] - n_s

    def model(reduced):
        x = g["x_train"][:, 1:] if reduced else g["x_train"]
        ls = g["lengthscale"][:, 1:] if reduced else g["lengthscale"]
        gp, _ = _make_models(se, x, g["y_train"], n_s - 1 if reduced else n_s, n_u, kern, ls, g["variance"], g["noise"])
        return gp

    for tag, fn, one in (("taylor", up.multi_step_taylor_symbolic, up.one_step_taylor),
                         ("meaneq", up.mean_equivalent_multistep, up.one_step_mean_equivalent)):
        for pr, (a, b, red, tm) in (("lin", (g["a"], g["b"], False, None)), ("nolin", (None, None, False, None)),
                                    ("trafo", (g["a"], g["b"], True, g["t_mat"]))):
            gp = model(red)
            mu, sig, var = fn(g["mu0"], gp, g["k_ff"], g["k_fb"], None, a, b, tm)      # batched call
            _assert_close(mu, g["mu_%s_%s" % (tag, pr)], RTOL_TIGHT, what="mu " + tag + pr)
            _assert_close(sig, g["sigma_%s_%s" % (tag, pr)], RTOL_TIGHT, atol_scale=1e-10, what="sigma " + tag + pr)
            assert var.shape == mu.shape
            # the reference's un-batched call shape, one trajectory
            hor = g["k_ff"].shape[1]
            m1, s1, third = fn(g["mu0"][0][:, None], gp, g["k_ff"][0], g["k_fb"][0], None, a, b, tm)
            assert m1.shape == (hor, n_s) and s1.shape == (hor, n_s * n_s)
            assert third.shape == ((1 + (hor - 1) * n_s, n_s) if tag == "taylor" else (hor, n_s))
            _assert_close(m1, g["mu_%s_%s" % (tag, pr)][0], RTOL_TIGHT)
            _assert_close(s1.reshape(hor, n_s, n_s), g["sigma_%s_%s" % (tag, pr)][0], RTOL_TIGHT, atol_scale=1e-10)
            # one step with an input covariance equals the second step of the trajectory
            mu_n, sig_n, _ = one(m1[0][:, None], gp, g["k_ff"][0][1][:, None], s1[0].reshape(n_s, n_s), g["k_fb"][0][0],
                                 a, b, tm)
            assert mu_n.shape == (n_s, 1) and sig_n.shape == (n_s, n_s)
            _assert_close(mu_n[:, 0], m1[1], 1e-12)
            _assert_close(sig_n, s1[1].reshape(n_s, n_s), 1e-12)
            gp.close()
    with pytest.raises(NotImplementedError):
        up.multi_step_taylor_symbolic(g["mu0"][0][:, None], model(False), g["k_ff"][0], g["k_fb"][0], np.eye(n_s))
