"""CPU tests that pin the oracle (oracle/) before anything is compared against it.

* ellipsoid half: against the reference's OWN functions when /root/reference is present (build
  container), against the committed golden vectors the reference produced (everywhere), and against
  the reference's known-answer tests (reference test/test_utils_ellipsoid.py:13-94).
* GP half: kernel rows, k(x,x), predictive mean and variance against the reference's OWN
  ssm_gpy/gp_models_utils_casadi.py functions (kernels incl. the composite lin_* ones, gp_pred), run through the
  NumPy-backed CasADi shim -- live in the build container and as golden vectors everywhere; the posterior state GPy
  would compute (K^-1, K^-1 y) and the closed-form Jacobian by identities (explicit-inverse form == Cholesky form,
  analytic Jacobian == finite differences, K K^-1 = I), with the reference's own acceptance tolerances
  r_tol=1e-4 / a_tol=1e-6 (reference test/test_gp_models.py:22-23) as the loosest bound.
"""
import os

import numpy as np
import pytest

from oracle import gp_oracle, reach_oracle, ref_loader
from oracle.gp_oracle import GPOracle


def _gp_from_golden(d):
    return GPOracle(d["x_train"], d["y_train"], [str(k) for k in d["kern_types"]], d["lengthscale"],
                    d["variance"], d["noise"])


def _rand_gp(rng, n=40, n_s=3, n_u=2, kerns=None):
    dim = n_s + n_u
    x = rng.uniform(-1, 1, size=(n, dim))
    y = np.tanh(x @ rng.randn(dim, n_s)) + 0.05 * rng.randn(n, n_s)
    kerns = kerns or (["rbf", "mat52", "rbf", "mat52"] * 3)[:n_s]
    return GPOracle(x, y, kerns, rng.uniform(0.7, 2.0, size=(n_s, dim)), rng.uniform(0.5, 1.5, size=n_s),
                    rng.uniform(0.01, 0.05, size=n_s))


# ------------------------------------------------------------------ GP half: identities
def test_gp_explicit_inverse_equals_cholesky_form():
    rng = np.random.RandomState(0)
    gp = _rand_gp(rng)
    z = rng.uniform(-1, 1, size=(17, gp.dim_in))
    mu_e, var_e = gp.predict(z, form="explicit")
    mu_c, var_c = gp.predict(z, form="chol")
    assert np.allclose(mu_e, mu_c, rtol=1e-12, atol=1e-12)
    assert np.allclose(var_e, var_c, rtol=1e-8, atol=1e-11)
    assert np.all(var_c > 0)


def test_gp_inverse_is_inverse_and_beta_solves():
    rng = np.random.RandomState(1)
    gp = _rand_gp(rng)
    gp._ensure_inv()
    for d in range(gp.n_s_out):
        k = gp_oracle.kernel(gp.kern_types[d], gp.x_train, gp.x_train, gp.variance[d], gp.lengthscale[d])
        k = k + gp.noise[d] * np.eye(gp.n_train)
        assert np.allclose(k @ gp.inv_K[d], np.eye(gp.n_train), atol=1e-8)
        assert np.allclose(k @ gp.beta[:, d], gp.y_train[:, d], atol=1e-9)


def test_gp_jacobian_matches_finite_differences():
    rng = np.random.RandomState(2)
    gp = _rand_gp(rng, n=60, n_s=4, n_u=1)
    z = rng.uniform(-1, 1, size=(9, gp.dim_in))
    assert np.allclose(gp.jacobian(z), gp.jacobian_fd(z), rtol=1e-5, atol=1e-7)


def test_gp_kernel_values():
    """Literal kernel values: k(x,x)=variance; rbf at unit scaled distance = s2*exp(-1/2);
    mat52 at r=1 = s2*(1+sqrt5+5/3)*exp(-sqrt5)."""
    x = np.zeros((1, 3))
    y = np.array([[2.0, 0.0, 0.0]])
    ls = np.array([2.0, 1.0, 5.0])
    assert np.allclose(gp_oracle.k_rbf(x, x, 1.7, ls), 1.7)
    assert np.allclose(gp_oracle.k_rbf(x, y, 1.7, ls), 1.7 * np.exp(-0.5))
    assert np.allclose(gp_oracle.k_mat52(x, y, 0.3, ls), 0.3 * (1 + np.sqrt(5) + 5. / 3) * np.exp(-np.sqrt(5)))


def test_gp_call_surface_shapes_and_single_point_only():
    rng = np.random.RandomState(3)
    gp = _rand_gp(rng, n_s=2, n_u=1)
    mu, var, jac = gp(np.zeros((1, 2)), np.zeros((1, 1)))
    assert mu.shape == (2, 1) and var.shape == (2, 1) and jac.shape == (2, 3)
    with pytest.raises(NotImplementedError):   # reference ssm_gpy/gaussian_process.py:142-143
        gp(np.zeros((2, 2)), np.zeros((2, 1)))


# ------------------------------------------------------------------ ellipsoid half: known answers
@pytest.mark.parametrize("ub,pts", [([0.1, 0.3, 0.5], [[-0.1, -0.3, 0.5], [-0.1, 0.3, -0.5], [0.1, 0.3, 0.5]]),
                                    ([0.1] * 3, [[-0.1, 0.1, 0.1], [-0.1, -0.1, -0.1], [0.1, 0.1, 0.1]])])
def test_ellipsoid_from_rectangle_known_answer(ub, pts):
    """reference test/test_utils_ellipsoid.py:13-65: the box corners lie on the ellipsoid."""
    q = reach_oracle.ellipsoid_from_rectangle(ub)
    pts = np.array(pts)
    d = np.sum(pts * np.linalg.solve(q, pts.T).T, axis=1)
    assert np.all(np.abs(d - 1) <= 1e-5)
    assert np.all(np.linalg.eigvals(q) > 0)


def test_ellipsoid_from_rectangle_negative_bound_raises():
    with pytest.raises(Exception):             # reference test/test_utils_ellipsoid.py:28-33
        reach_oracle.ellipsoid_from_rectangle([0.6, -0.3])


def test_golden_ellipsoid_algebra(golden_dir):
    g = np.load(os.path.join(golden_dir, "ellipsoid_algebra.npz"))
    for tag in ("t_1", "t_2", "t_3", "t_4"):
        u_mu, u_sig = reach_oracle.compute_remainder_overapproximations(g[tag + "_q"], g[tag + "_k_fb"],
                                                                        g[tag + "_l_mu"], g[tag + "_l_sigma"])
        assert np.allclose(u_mu, g[tag + "_u_mu"], rtol=1e-12)
        assert np.allclose(u_sig, g[tag + "_u_sigma"], rtol=1e-12)
        um_b, us_b = reach_oracle.remainder_batch(g[tag + "_q"][None], g[tag + "_k_fb"], g[tag + "_l_mu"],
                                                  g[tag + "_l_sigma"])
        assert np.allclose(um_b[0], g[tag + "_u_mu"], rtol=1e-12)
        assert np.allclose(us_b[0], g[tag + "_u_sigma"], rtol=1e-12)
    for i in range(3):
        p, q = reach_oracle.sum_two_ellipsoids(g["s%d_p1" % i], g["s%d_q1" % i], g["s%d_p2" % i], g["s%d_q2" % i])
        assert np.allclose(p, g["s%d_p" % i], rtol=1e-14) and np.allclose(q, g["s%d_q" % i], rtol=1e-14)
        assert np.allclose(reach_oracle.ellipsoid_from_rectangle(g["s%d_ub" % i]), g["s%d_qrect" % i], rtol=1e-14)


# ------------------------------------------------------------------ ellipsoid half: golden from the reference
def test_golden_invpend_c1(golden_dir):
    g = np.load(os.path.join(golden_dir, "invpend_c1.npz"))
    gp = _gp_from_golden(g)
    for i in range(g["k_ff"].shape[0]):
        _, _, p_all, q_all = reach_oracle.multistep_reachability(g["p0"][:, None], gp, g["k_fb"], g["k_ff"][i],
                                                                 g["l_mu"], g["l_sigma"], None, float(g["c_safety"]))
        assert np.allclose(p_all, g["p_all"][i], rtol=1e-10, atol=1e-12)
        assert np.allclose(q_all, g["q_all"][i], rtol=1e-10, atol=1e-12)
    p_b, q_b, _ = reach_oracle.multistep_batch(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"], g["l_sigma"], None,
                                               float(g["c_safety"]))
    assert np.allclose(p_b, g["p_all"], rtol=1e-8, atol=1e-10)
    assert np.allclose(q_b, g["q_all"], rtol=1e-8, atol=1e-10)


def test_golden_invpend_reach_test(golden_dir):
    g = np.load(os.path.join(golden_dir, "invpend_reach_test.npz"))
    gp = _gp_from_golden(g)
    c = float(g["c_safety"])
    for tag, a, b in (("lin", g["a"], g["b"]), ("nolin", None, None)):
        p1, q1 = reach_oracle.onestep_reachability(g["p"], gp, g["k_ff"], g["l_mu"], g["l_sigma"], g["q"], g["k_fb"],
                                                   c, 0, a, b)
        assert np.allclose(p1, g["p1_set_" + tag], rtol=1e-10) and np.allclose(q1, g["q1_set_" + tag], rtol=1e-10)
        p1, q1 = reach_oracle.onestep_reachability(g["p"], gp, g["k_ff"], g["l_mu"], g["l_sigma"], None, g["k_fb"],
                                                   c, 0, a, b)
        assert np.allclose(p1, g["p1_point_" + tag], rtol=1e-10) and np.allclose(q1, g["q1_point_" + tag], rtol=1e-10)
    _, _, p_all, q_all = reach_oracle.multistep_reachability(g["p"], gp, g["k_fb_multi"], g["k_ff_multi"], g["l_mu"],
                                                             g["l_sigma"], None, c, 0, g["a"], g["b"], None)
    assert np.allclose(p_all, g["p_all"], rtol=1e-10) and np.allclose(q_all, g["q_all"], rtol=1e-10)
    dist = reach_oracle.lin_ellipsoid_safety_distance(p_all[-1][:, None], q_all[-1], g["h_mat"], g["h_vec"], c)
    assert np.allclose(dist, g["dist"], rtol=1e-10)
    assert np.allclose(reach_oracle.safety_distance_batch(p_all, q_all, g["h_mat"], g["h_vec"], c)[-1],
                       g["dist"][:, 0], rtol=1e-10)


def test_golden_cartpole_batch_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "cartpole.npz"))
    gp = _gp_from_golden(g)
    p_b, q_b, _ = reach_oracle.multistep_batch(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"], g["l_sigma"], None, 2.0,
                                               g["a"], g["b"])
    assert np.allclose(p_b, g["p_all"], rtol=1e-7, atol=1e-10)
    assert np.allclose(q_b, g["q_all"], rtol=1e-7, atol=1e-10)
    p_b, q_b, _ = reach_oracle.multistep_batch(g["p0"], gp, g["k_fb"], g["k_ff"], g["l_mu"], g["l_sigma"], g["q0"],
                                               1.5, g["a"], g["b"], g["k_fb_init"])
    assert np.allclose(p_b, g["p_all_q0"], rtol=1e-7, atol=1e-10)
    assert np.allclose(q_b, g["q_all_q0"], rtol=1e-7, atol=1e-10)


# ------------------------------------------------------------------ GP half pinned by the reference's own functions
GP_GOLDEN_CASES = ("pend_rbf_mat52", "pend_composite", "cart_mixed")


@pytest.mark.parametrize("name", GP_GOLDEN_CASES)
def test_golden_gp_pred_reference(golden_dir, name):
    """Kernel rows, k(z,z), predictive mean and variance produced by the reference's gp_models_utils_casadi.py
    (_k_rbf/_k_mat52/_k_lin_rbf/_k_lin_mat52 + gp_pred) -- the oracle must reproduce them."""
    g = np.load(os.path.join(golden_dir, "gp_pred_reference.npz"))
    ora, kerns, _ = gp_oracle.golden_gp_case(g, name)
    z = g[name + "/z"]
    for d in range(len(kerns)):
        assert np.allclose(ora.kstar(d, z), g[name + "/kstar"][d], rtol=1e-12, atol=1e-14)
        assert np.allclose(ora.prior_var(d, z), g[name + "/prior"][:, d], rtol=1e-13)
    for form in ("explicit", "chol"):
        mu, var = ora.predict(z, form=form)
        assert np.allclose(mu, g[name + "/mu"], rtol=1e-8, atol=1e-10)
        assert np.allclose(var, g[name + "/var"], rtol=1e-7, atol=1e-11)
    assert np.allclose(ora.jacobian(z), ora.jacobian_fd(z), rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_gp_oracle_equals_reference_gp_utils_live():
    """The same comparison live, on fresh random inputs, for every kernel type and both composite semantics' building
    blocks (_k_lin on all columns is the GPy Linear(ARD) kernel)."""
    ref = ref_loader.load_gp_utils()
    rng = np.random.RandomState(3)
    n, dim, t = 35, 4, 9
    x = rng.uniform(-1.5, 1.5, size=(n, dim))
    z = rng.uniform(-1.5, 1.5, size=(t, dim))
    y = rng.randn(n, 4)
    kerns = ["rbf", "mat52", "lin_rbf", "lin_mat52"]
    hyp = [{"lengthscale": rng.uniform(0.5, 2, dim), "variance": 1.3},
           {"lengthscale": rng.uniform(0.5, 2, dim), "variance": 0.6},
           {"prod.rbf.lengthscale": np.array([0.9]), "prod.rbf.variance": 1.1, "prod.linear.variances": np.array([0.7]),
            "linear.variances": rng.uniform(0.1, 0.5, dim)},
           {"prod.mat52.lengthscale": np.array([1.4]), "prod.mat52.variance": 0.8,
            "prod.linear.variances": np.array([1.2]), "linear.variances": rng.uniform(0.1, 0.5, dim)}]
    ls, var, pl, lin = gp_oracle.vectors_from_reference_hyp(kerns, hyp, dim)
    ora = gp_oracle.GPOracle(x, y, kerns, ls, var, np.full(4, 0.02), prod_linear=pl, linear=lin)
    ora._ensure_inv()
    assert np.allclose(gp_oracle.unscaled_dist(z, x), ref._unscaled_dist(z, x), rtol=1e-12)
    assert np.allclose(gp_oracle.k_lin(z, x, lin[2]), ref._k_lin(z, x, lin[2]), rtol=1e-13, atol=1e-15)
    for d, k in enumerate(kerns):
        kfun = ref._get_kernel_function(k, hyp[d])
        assert np.allclose(ora.kstar(d, z), kfun(z, y=x), rtol=1e-12, atol=1e-14)
        assert np.allclose(ora.prior_var(d, z), np.asarray(kfun(z, diag_only=True)).reshape(-1), rtol=1e-13)
        m_r, v_r = ref.gp_pred(z, kfun, ora.beta[:, d:d + 1], x, ora.inv_K[d])
        m_o, v_o = ora.predict(z, form="explicit")
        assert np.allclose(m_o[:, d], np.asarray(m_r).reshape(-1), rtol=1e-11, atol=1e-13)
        assert np.allclose(v_o[:, d], np.asarray(v_r).reshape(-1), rtol=1e-9, atol=1e-13)


# ------------------------------------------------------------------ greedy max-variance selection
def test_select_oracle_incremental_equals_brute_force():
    """The incremental partial-Cholesky form picks, at every step, the arg-max of the GP predictive variance given
    the points chosen so far (the criterion of choose_datapoints_maxvar, ssm_gpy/gaussian_process.py:332-333),
    computed by a fresh Cholesky solve."""
    from oracle import select_oracle
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (150, 3))
    for kerns, pl, lin in ((["rbf", "mat52"], None, None),
                           (["lin_rbf", "mat52"], np.array([[0, .8, 0], [0, 0, 0]]), np.array([[.1, .2, .3], [0, 0, 0]]))):
        ls = rng.uniform(0.5, 1.5, (2, 3))
        if pl is not None:
            ls[0] = [np.inf, 0.9, np.inf]
        var, noise = [1.2, 0.7], [1e-2, 2e-2]
        idx, score = select_oracle.greedy_maxvar(x, kerns, ls, var, noise, 30, pl, lin)
        assert len(set(idx.tolist())) == 30
        for t in (0, 1, 2, 7, 29):
            s = select_oracle.brute_force_scores(x, kerns, ls, var, noise, idx[:t], pl, lin)
            s[idx[:t]] = -np.inf
            assert int(np.argmax(s)) == idx[t]
            assert np.isclose(s[idx[t]], score[t], rtol=1e-8)
        assert np.all(np.diff(score) <= 1e-12)          # greedy scores never increase


# ------------------------------------------------------------------ live against the reference (build container only)
@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_equals_reference_functions_live():
    reach, utils, uell = ref_loader.load()
    rng = np.random.RandomState(11)
    for n_s, n_u in ((2, 1), (4, 1), (5, 2)):
        gp = _rand_gp(rng, n=30, n_s=n_s, n_u=n_u)
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
            pr, qr = reach.onestep_reachability(p, gp, k_ff, l_mu, l_sig, qq, k_fb, 1.7, 0, a, b)
            po, qo = reach_oracle.onestep_reachability(p, gp, k_ff, l_mu, l_sig, qq, k_fb, 1.7, 0, a, b)
            assert np.allclose(po, np.real(pr), rtol=1e-12) and np.allclose(qo, np.real(qr), rtol=1e-12)
        um_r, us_r = utils.compute_remainder_overapproximations(q, k_fb, l_mu, l_sig)
        um_o, us_o = reach_oracle.compute_remainder_overapproximations(q, k_fb, l_mu, l_sig)
        assert np.iscomplexobj(um_r)      # the reference's complex128 leak (utils.py:133-141)
        assert np.allclose(um_o, np.real(um_r), rtol=1e-13) and np.allclose(us_o, np.real(us_r), rtol=1e-13)
        hor = 4
        kfb = 0.5 * rng.randn(hor - 1, n_u, n_s)
        kff = 0.2 * rng.randn(hor, n_u)
        _, _, pa_r, qa_r = reach.multistep_reachability(p, gp, kfb, kff, l_mu, l_sig, None, 2.0, 0, a, b, None)
        pb, qb, _ = reach_oracle.multistep_batch(p, gp, kfb, kff[None], l_mu, l_sig, None, 2.0, a, b)
        assert np.allclose(pb[0], np.real(pa_r), rtol=1e-9) and np.allclose(qb[0], np.real(qa_r), rtol=1e-9)


# ---------------------------------------------------------------------------------- scoring oracle (SafeMPC assembly)
def _random_candidates(rng, bsz, hor, n_s, n_u):
    p_all = 0.3 * rng.standard_normal((bsz, hor, n_s))
    m = 0.2 * rng.standard_normal((bsz, hor, n_s, n_s))
    q_all = m @ np.swapaxes(m, -1, -2) + 1e-3 * np.eye(n_s)
    var_all = rng.uniform(1e-4, 1e-2, size=(bsz, hor, n_s))
    k_ff = 0.5 * rng.standard_normal((bsz, hor, n_u))
    k_fb = rng.standard_normal((hor - 1, n_u, n_s))
    return p_all, q_all, var_all, k_ff, k_fb


def test_score_oracle_known_answer_and_reference_distance():
    """Hand-computable case + the assembly evaluated with the reference's own lin_ellipsoid_safety_distance."""
    from oracle import ref_loader, score_oracle
    # one candidate, H=2, n_s=2, n_u=1: Q = diag(0.04, 0.09), K = [1, 0] -> sqrt(K Q K^T) = 0.2
    p_all = np.array([[0.1, -0.2], [0.3, 0.4]])
    q_all = np.array([np.diag([0.04, 0.09]), np.diag([0.01, 0.16])])
    k_ff = np.array([[0.5], [-0.25]])
    k_fb = np.array([[[1.0, 0.0]]])
    cb = np.array([[-1.0, 1.0]])
    h_mat = np.vstack((np.eye(2), -np.eye(2)))
    h_vec = np.ones((4, 1))
    g = score_oracle.constraints_one(p_all, q_all, k_ff, k_fb, cb, h_mat, h_vec, h_mat, h_vec)
    want = np.array([0.5 - 1.0, -1.0 - 0.5,                      # u_0 bounds
                     -0.25 + 0.2 - 1.0, 0.25 + 0.2 - 1.0,        # step-1 control through the ellipsoid
                     0.1 + 0.2 - 1, -0.2 + 0.3 - 1, -0.1 + 0.2 - 1, 0.2 + 0.3 - 1,     # obstacle polytope, ellipsoid 0
                     0.3 + 0.1 - 1, 0.4 + 0.4 - 1, -0.3 + 0.1 - 1, -0.4 + 0.4 - 1])    # terminal polytope, ellipsoid 1
    assert np.allclose(g, want, rtol=0, atol=1e-15)
    assert np.isclose(score_oracle.exploration_cost_one(np.array([[0.04, 0.05], [0.16, 0.09]])), -(0.3 + 0.5))
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    gp_reach, _, _ = ref_loader.load()
    rng = np.random.default_rng(3)
    p, q, v, kff, kfb = _random_candidates(rng, 5, 4, 3, 2)
    cb = np.array([[-1.0, 1.0], [-0.5, 0.7]])
    h_obs = rng.standard_normal((5, 3))
    c_own, f_own, v_own, g_own = score_oracle.score_batch(p, q, v, kff, kfb, cb, h_obs, np.ones((5, 1)), h_obs[:2],
                                                          np.ones((2, 1)))
    c_ref, f_ref, v_ref, g_ref = score_oracle.score_batch(p, q, v, kff, kfb, cb, h_obs, np.ones((5, 1)), h_obs[:2],
                                                          np.ones((2, 1)),
                                                          dist_fn=gp_reach.lin_ellipsoid_safety_distance)
    assert g_own.shape == (5, 2 * 2 * 4 + 3 * 5 + 2)
    assert np.allclose(g_own, np.real(g_ref), rtol=1e-14, atol=1e-15)
    assert np.array_equal(f_own, f_ref)


# ---------------------------------------------------------------------------------- Gaussian uncertainty propagation
def _uprop_gp(g, reduced=False):
    from oracle.gp_oracle import GPOracle
    kern = [str(k) for k in g["kern_types"]]
    if reduced:
        return GPOracle(g["x_train"][:, 1:], g["y_train"], kern, g["lengthscale"][:, 1:], g["variance"], g["noise"])
    return GPOracle(g["x_train"], g["y_train"], kern, g["lengthscale"], g["variance"], g["noise"])


def test_golden_uncertainty_propagation_batch_oracle(golden_dir):
    """Closed-form batch oracle == outputs of the reference's multi_step_taylor_symbolic / mean_equivalent_multistep."""
    from oracle import uprop_oracle
    g = np.load(os.path.join(golden_dir, "uncertainty_propagation.npz"))
    for tag, taylor in (("taylor", True), ("meaneq", False)):
        for pr, (a, b, red, tm) in (("lin", (g["a"], g["b"], False, None)), ("nolin", (None, None, False, None)),
                                    ("trafo", (g["a"], g["b"], True, g["t_mat"]))):
            mu, sig, _ = uprop_oracle.multistep_batch(g["mu0"], _uprop_gp(g, red), g["k_ff"], g["k_fb"], a, b, tm,
                                                      taylor)
            assert np.allclose(mu, g["mu_%s_%s" % (tag, pr)], rtol=1e-10, atol=1e-12)
            assert np.allclose(sig, g["sigma_%s_%s" % (tag, pr)], rtol=1e-9, atol=1e-13)


def test_uncertainty_propagation_oracle_equals_reference_live():
    from oracle import ref_loader, uprop_oracle
    from oracle.gp_oracle import GPOracle
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    up = ref_loader.load_uncertainty_propagation()
    rng = np.random.default_rng(21)
    n_s, n_u, n, hor = 3, 2, 40, 4
    x = rng.uniform(-1, 1, (n, n_s + n_u))
    y = np.sin(x @ rng.standard_normal((n_s + n_u, n_s)))
    gp = GPOracle(x, y, ["rbf", "mat52", "rbf"], rng.uniform(0.8, 2, (n_s, n_s + n_u)), rng.uniform(.5, 1.5, n_s),
                  np.full(n_s, 0.01))
    a = np.eye(n_s) + 0.1 * rng.standard_normal((n_s, n_s))
    b = 0.3 * rng.standard_normal((n_s, n_u))
    k_ff = 0.1 * rng.standard_normal((hor, n_u))
    k_fb = 0.2 * rng.standard_normal((hor - 1, n_u, n_s))
    mu0 = 0.1 * rng.standard_normal((n_s, 1))
    for fn, taylor in ((up.multi_step_taylor_symbolic, True), (up.mean_equivalent_multistep, False)):
        m_ref, s_ref, _ = fn(mu0, gp, k_ff, k_fb, None, a, b)
        m, s, _ = uprop_oracle.multistep_batch(mu0[:, 0], gp, k_ff[None], k_fb, a, b, None, taylor)
        assert np.allclose(m[0], m_ref, rtol=1e-12, atol=1e-14)
        assert np.allclose(s[0].reshape(hor, -1), s_ref, rtol=1e-10, atol=1e-14)
