"""GPU tests of the scoring kernels and the sampling SafeMPC driver, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def se():
    import safe_exploration_b200 as pkg
    pkg._lib.load()
    return pkg


def _model_and_rollout(se, name="C2", n_train=300, horizon=6, batch=257):
    from safe_exploration_b200 import workloads
    w = workloads.make(name, batch=batch, n_train=n_train, horizon=horizon)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    res = se.rollout(gp, w.p0, w.k_ff, w.k_fb, w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    return w, gp, res


@pytest.mark.parametrize("name,per_candidate_gain", [("C2", False), ("C3", True)])
def test_score_matches_oracle(se, name, per_candidate_gain):
    from oracle import score_oracle
    w, gp, res = _model_and_rollout(se, name)
    rng = np.random.default_rng(4)
    k_fb = w.k_fb
    if per_candidate_gain:
        k_fb = w.k_fb[None] + 0.01 * rng.standard_normal((w.k_ff.shape[0],) + w.k_fb.shape)
        res = se.rollout(gp, w.p0, w.k_ff, k_fb, w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
    h_obs = rng.standard_normal((3, w.n_s))
    # thresholds from the data (constraint values with all bounds at zero), so that each family of constraints
    # rejects a share of the candidates and the batch mixes feasible and infeasible ones
    zero_cb = np.zeros((w.n_u, 2))
    _, _, _, g0 = score_oracle.score_batch(res.p_all, res.q_all, res.var_all, w.k_ff, k_fb, zero_cb, h_obs,
                                           np.zeros((3, 1)), w.h_mat, np.zeros((2 * w.n_s, 1)))
    n_c = 2 * w.n_u * w.k_ff.shape[1]
    n_o = 3 * (w.k_ff.shape[1] - 1)
    u_abs = float(np.quantile(g0[:, :n_c].max(axis=1), 0.8))
    cb = np.stack((-u_abs * np.ones(w.n_u), u_abs * np.ones(w.n_u)), axis=1)
    h_obs_v = float(np.quantile(g0[:, n_c:n_c + n_o].max(axis=1), 0.8)) * np.ones((3, 1))
    h_safe_v = float(np.quantile(g0[:, n_c + n_o:].max(axis=1), 0.8)) * np.ones((2 * w.n_s, 1))
    for cost, kw in (("exploration", {}), ("quadratic", {"wx": np.diag(np.arange(1, w.n_s + 1.0)),
                                                         "wu": 0.5 * np.eye(w.n_u), "x_ref": 0.1 * np.ones(w.n_s)})):
        sc = se.score_rollouts(res, w.k_ff, k_fb, w.h_mat, h_safe_v, cb, h_obs, h_obs_v, cost=cost, want_g=True, **kw)
        c_o, f_o, v_o, g_o = score_oracle.score_batch(res.p_all, res.q_all, res.var_all, w.k_ff, k_fb, cb, h_obs,
                                                      h_obs_v, w.h_mat, h_safe_v, cost=cost, **kw)
        assert sc.g.shape == g_o.shape
        assert np.allclose(sc.g, g_o, rtol=1e-12, atol=1e-13)
        assert np.allclose(sc.cost, c_o, rtol=1e-12, atol=1e-13)
        assert np.allclose(sc.violation, v_o, rtol=1e-12, atol=1e-13)
        assert np.array_equal(sc.feasible.astype(bool), f_o)
        assert 0 < f_o.sum() < f_o.size, "test case should mix feasible and infeasible candidates"
        idx, c, v, f = se.best_candidate(sc)
        want = int(np.flatnonzero(f_o)[np.argmin(c_o[f_o])])
        assert f and idx == want and np.isclose(c, c_o[want]) and np.isclose(v, v_o[want])
    # without control bounds / obstacles only the terminal rows remain
    sc = se.score_rollouts(res, w.k_ff, k_fb, w.h_mat, h_safe_v, want_g=True)
    assert sc.g.shape[1] == 2 * w.n_s
    gp.close()


def test_best_candidate_without_feasible_and_with_bad_status(se):
    w, gp, res = _model_and_rollout(se, batch=64)
    tight = 1e-4 * np.ones((2 * w.n_s, 1))          # nobody fits
    sc = se.score_rollouts(res, w.k_ff, w.k_fb, w.h_mat, tight)
    assert sc.feasible.sum() == 0
    idx, c, v, f = se.best_candidate(sc)
    assert not f and idx == int(np.argmin(sc.violation)) and np.isclose(v, sc.violation.min())
    # a candidate whose rollout reported a failure is never feasible, however good its numbers look
    loose = 10.0 * np.ones((2 * w.n_s, 1))
    status = res.status.copy()
    sc0 = se.score_rollouts(res, w.k_ff, w.k_fb, w.h_mat, loose)
    best0 = se.best_candidate(sc0)[0]
    status[best0] = 1
    sc1 = se.score_rollouts(res._replace(status=status), w.k_ff, w.k_fb, w.h_mat, loose)
    assert sc0.feasible.all() and sc1.feasible[best0] == 0 and sc1.feasible.sum() == 63
    assert se.best_candidate(sc1)[0] != best0
    gp.close()


def test_score_device_tensors_match_host_arrays(se):
    import torch
    w, gp, res = _model_and_rollout(se, batch=100)
    dev = gp.device
    kff_d = torch.as_tensor(w.k_ff, device=dev)
    res_d = se.rollout(gp, torch.as_tensor(w.p0, device=dev), kff_d, torch.as_tensor(w.k_fb, device=dev), w.l_mu,
                       w.l_sigma, None, None, w.c_safety, w.a, w.b)
    safe_v = 0.3 * np.ones((2 * w.n_s, 1))
    sc_d = se.score_rollouts(res_d, kff_d, torch.as_tensor(w.k_fb, device=dev), w.h_mat, safe_v)
    sc_h = se.score_rollouts(res, w.k_ff, w.k_fb, w.h_mat, safe_v)
    assert torch.is_tensor(sc_d.cost) and sc_d.cost.is_cuda
    assert np.array_equal(sc_d.cost.cpu().numpy(), sc_h.cost)
    assert np.array_equal(sc_d.feasible.cpu().numpy(), sc_h.feasible)
    assert se.best_candidate(sc_d) == se.best_candidate(sc_h)
    gp.close()


def _pendulum_mpc(se, n_safe=4, h_safe_bound=0.5, **kw):
    from safe_exploration_b200 import workloads
    w = workloads.make("C2", batch=8, n_train=200, horizon=n_safe)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    opt_env = {"l_mu": w.l_mu, "l_sigma": w.l_sigma, "h_mat_safe": w.h_mat,
               "h_safe": h_safe_bound * np.ones((2 * w.n_s, 1)), "lin_model": (w.a, w.b),
               "ctrl_bounds": np.array([[-1.0, 1.0]])}
    kw.setdefault("opt_perf_trajectory", {"n_perf": 1})       # no performance trajectory unless a test asks for one
    mpc = se.SamplingSafeMPC(n_safe, gp, opt_env, np.eye(w.n_s), np.eye(w.n_u), beta_safety=2.0, n_samples=512,
                             n_iter=2, n_elite=32, seed=1, **kw)
    return w, gp, mpc


def test_sampling_mpc_returns_a_feasible_action_and_its_certificate(se):
    from oracle import reach_oracle, score_oracle
    from oracle.gp_oracle import GPOracle
    w, gp, mpc = _pendulum_mpc(se)
    x0 = np.array([0.02, -0.03])
    u, feasible, success, k_fb, k_ff_all, p_safe, q_safe = mpc.get_action(x0, sol_verbose=True)
    assert feasible and success and u.shape == (w.n_u,) and -1.0 <= u[0] <= 1.0
    assert k_ff_all.shape == (3, w.n_u) and p_safe.shape == (4, w.n_s) and q_safe.shape == (4, w.n_s, w.n_s)
    assert mpc.n_fail == 0
    # certificate: the returned plan, re-evaluated by the CPU oracle, satisfies every constraint
    ora = GPOracle(w.x_train, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in w.hyp]),
                   [h["variance"] for h in w.hyp], gp.total_noise())
    seq = np.vstack((u[None], k_ff_all))
    p_o, q_o, v_o = reach_oracle.multistep_batch(x0, ora, k_fb.reshape(3, w.n_u, w.n_s), seq[None], w.l_mu, w.l_sigma,
                                                 None, 2.0, w.a, w.b)
    assert np.allclose(p_o[0], p_safe, rtol=1e-6, atol=1e-9) and np.allclose(q_o[0], q_safe, rtol=1e-6, atol=1e-9)
    g = score_oracle.constraints_one(p_o[0], q_o[0], seq, k_fb.reshape(3, w.n_u, w.n_s), mpc.ctrl_bounds, None, None,
                                     mpc.h_mat_safe, mpc.h_safe)
    assert np.all(g < 1e-5)
    feas2, g_term = mpc.eval_safety_constraints(p_safe, q_safe)
    assert feas2 and g_term.shape == (2 * w.n_s, 1)
    # receding horizon: the next call starts from the shifted plan and stays feasible
    u2, success2 = mpc.get_action(p_safe[0])
    assert success2 and mpc.n_fail == 0 and u2.shape == (w.n_u,)
    # lqr_only short-cut of the reference
    u_lqr, fail = mpc.get_action(x0, lqr_only=True)
    assert np.allclose(u_lqr, mpc.get_lqr_feedback().reshape(w.n_u, w.n_s) @ x0) and fail is False
    gp.close()


def test_sampling_mpc_falls_back_like_the_reference(se):
    """Infeasible problem: first the shifted old solution (n_fail < n_safe), then the safe policy."""
    w, gp, mpc = _pendulum_mpc(se)
    x0 = np.array([0.02, -0.03])
    u, feasible, *_ = mpc.get_action(x0, sol_verbose=True)
    assert feasible
    old = (mpc.k_ff_safe.copy(), mpc.k_fb_safe_all.copy(), mpc.p_safe.copy())
    mpc.h_safe = 1e-6 * np.ones_like(mpc.h_safe)        # now nothing is feasible
    x1 = old[2][0] + 1e-3
    u1, feasible1, *_ = mpc.get_action(x1, sol_verbose=True)
    assert not feasible1 and mpc.n_fail == 1
    want = old[0][0] + old[1][0].reshape(w.n_u, w.n_s) @ (x1 - old[2][0])      # feedback_ctrl on the old plan
    assert np.allclose(u1, want)
    for _ in range(3):
        u_k, feasible_k, *_ = mpc.get_action(x1, sol_verbose=True)
    assert mpc.n_fail == 4 and np.allclose(u_k, mpc.safe_policy(x1))
    gp.close()


@pytest.mark.parametrize("perf", ["mean_equivalent", "taylor"])
def test_sampling_mpc_performance_trajectory(se, perf):
    """n_perf > 1 (the reference's default, safempc_simple.py:19-20): every candidate carries performance controls;
    the cost is the reference's default with a performance trajectory (:292-303), recomputed here from the returned
    plan with the oracle's Gaussian propagation."""
    from oracle import uprop_oracle
    from oracle.gp_oracle import GPOracle
    w, gp, mpc = _pendulum_mpc(se, opt_perf_trajectory={"n_perf": 5, "type_perf_traj": perf})
    assert mpc.n_perf == 5 and mpc.r == 1 and mpc.perf_has_fb
    x0 = np.array([0.02, -0.03])
    u, feasible, success, k_fb, k_ff_all, p_safe, q_safe = mpc.get_action(x0, sol_verbose=True)
    assert feasible and success and mpc.k_ff_perf.shape == (4, w.n_u)
    assert np.all(np.abs(mpc.k_ff_perf) <= 1.0 + 1e-12)
    # the cost of the returned plan, by the oracle
    ora = GPOracle(w.x_train, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in w.hyp]),
                   [h["variance"] for h in w.hyp], gp.total_noise())
    seq_perf = np.vstack((u[None], mpc.k_ff_perf))
    k_fb_perf = mpc.get_lqr_feedback().reshape(w.n_u, w.n_s)
    res_p = mpc._rollout_perf(x0, seq_perf[None], k_fb_perf)
    mu_o, sig_o, var_o = uprop_oracle.multistep_batch(x0, ora, seq_perf[None], np.tile(k_fb_perf[None], (4, 1, 1)), w.a,
                                                      w.b, None, perf == "taylor")
    mu_o, sig_o, var_o = mu_o[0], sig_o[0], var_o[0]
    assert np.allclose(res_p.p_all[0], mu_o, rtol=1e-6, atol=1e-9) and np.allclose(res_p.var_all[0], var_o, rtol=1e-5)
    assert np.allclose(res_p.q_all[0], sig_o, rtol=1e-5, atol=1e-12)
    d = mu_o[1:4] - p_safe[1:4]
    want = float(np.einsum("ti,ij,tj->", d, 0.1 * mpc.wx_cost, d) - np.sum(np.sqrt(np.sum(var_o, axis=1))))
    res_s = mpc._rollout(x0, np.vstack((u[None], k_ff_all))[None], k_fb.reshape(3, w.n_u, w.n_s))
    got = float(mpc._perf_cost(res_s, res_p)[0])
    assert abs(got - want) <= 1e-6 * abs(want)
    # receding horizon keeps the shifted performance controls as the next mean
    u2, success2 = mpc.get_action(p_safe[0])
    assert success2 and mpc.n_fail == 0
    with pytest.raises(NotImplementedError):
        se.SamplingSafeMPC(4, gp, {"l_mu": w.l_mu, "l_sigma": w.l_sigma, "h_mat_safe": w.h_mat,
                                   "h_safe": np.ones((4, 1)), "lin_model": (w.a, w.b)}, np.eye(2), np.eye(1),
                           opt_perf_trajectory={"r": 2})
    gp.close()


def test_sampling_mpc_init_uncertainty_custom_cost_and_opt_x0(se):
    from oracle import reach_oracle, score_oracle
    from oracle.gp_oracle import GPOracle
    w, gp, mpc = _pendulum_mpc(se)
    x0 = np.array([0.02, -0.03])
    q0 = np.diag([2e-4, 1e-4])
    with pytest.raises(ValueError):
        mpc.solve(x0, q_0=q0)                       # not initialised for it (safempc_simple.py:181-199)
    mpc.init_solver(init_uncertainty=True)
    _, u, feasible, success, k_fb, k_ff_all, p_safe, q_safe = mpc.solve(x0, q_0=q0, sol_verbose=True)
    assert feasible and success
    # certificate by the oracle, from the ellipsoid (x0, q0) with the LQR gain on step 0
    ora = GPOracle(w.x_train, w.y_train, w.kern_types, np.stack([h["lengthscale"] for h in w.hyp]),
                   [h["variance"] for h in w.hyp], gp.total_noise())
    k0 = mpc.get_lqr_feedback().reshape(w.n_u, w.n_s)
    seq = np.vstack((u[None], k_ff_all))
    p_o, q_o, _ = reach_oracle.multistep_batch(x0, ora, k_fb.reshape(3, w.n_u, w.n_s), seq[None], w.l_mu, w.l_sigma, q0,
                                               2.0, w.a, w.b, k0)
    assert np.allclose(p_o[0], p_safe, rtol=1e-6, atol=1e-9) and np.allclose(q_o[0], q_safe, rtol=1e-6, atol=1e-9)
    # the bound on u_0 carries the feedback term of (q0, k0): value by value against the reference's formula
    res = mpc._rollout(x0, seq[None], k_fb.reshape(3, w.n_u, w.n_s), q0, k0)
    sc = se.score_rollouts(res, seq[None], k_fb.reshape(3, w.n_u, w.n_s), mpc.h_mat_safe, mpc.h_safe, mpc.ctrl_bounds,
                           want_g=True, q_0=q0, k_fb_0=k0)
    sd0 = np.sqrt(np.diag(k0 @ q0 @ k0.T))
    assert np.allclose(sc.g[0, :2], [u[0] + sd0[0] - 1.0, -1.0 - u[0] + sd0[0]], rtol=1e-12, atol=1e-14)
    g_rest = score_oracle.constraints_one(p_o[0], q_o[0], seq, k_fb.reshape(3, w.n_u, w.n_s), mpc.ctrl_bounds, None, None,
                                          mpc.h_mat_safe, mpc.h_safe)
    assert np.allclose(sc.g[0, 2:], g_rest[2:], rtol=1e-6, atol=1e-9)
    # a custom cost with the reference's argument order picks the plan it asks for
    mpc.init_solver()
    _, u_d, feas_d, *_ = mpc.solve(x0, sol_verbose=True)
    seen = {}

    def cost(p_0, u_0, p_all, q_all, k_ff, k_fb_, sig):
        seen["shapes"] = (p_0.shape, u_0.shape, p_all.shape, q_all.shape, k_ff.shape, np.shape(k_fb_), sig.shape)
        return np.sum((u_0 - 0.25) ** 2, axis=1)
    mpc.init_solver(cost_func=cost)
    _, u_c, feas_c, *_ = mpc.solve(x0, sol_verbose=True)
    assert feas_d and feas_c and seen["shapes"] == ((512, 2), (512, 1), (512, 4, 2), (512, 4, 2, 2), (512, 3, 1), (3, 1, 2),
                                                    (512, 4, 2))
    assert abs(u_c[0] - 0.25) <= abs(u_d[0] - 0.25) + 1e-9 and u_c[0] != u_d[0]
    # opt_x0: the initial state is a decision variable; a cost that rewards x_0[0] pulls it there
    mpc.init_solver(cost_func=lambda p_0, u_0, *rest: -p_0[:, 0], opt_x0=True)
    mpc.n_iter = 3
    x_opt, u_o, ok = mpc.solve(x0)
    assert ok and x_opt.shape == (w.n_s, 1) and x_opt[0, 0] > x0[0] + 0.02
    gp.close()


def test_sampling_mpc_without_control_bounds_does_not_clip_u0(se):
    from safe_exploration_b200 import workloads
    w = workloads.make("C2", batch=8, n_train=200, horizon=2)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    opt_env = {"l_mu": w.l_mu, "l_sigma": w.l_sigma, "h_mat_safe": w.h_mat, "h_safe": 1e3 * np.ones((2 * w.n_s, 1)),
               "lin_model": (w.a, w.b)}
    mpc = se.SamplingSafeMPC(2, gp, opt_env, np.eye(w.n_s), np.eye(w.n_u), beta_safety=2.0, n_samples=256, n_iter=1,
                             n_elite=16, seed=3, opt_perf_trajectory={"n_perf": 1}, sigma0=1e-3)
    mpc.init_solver(cost_func=lambda p_0, u_0, *rest: np.sum((u_0 - 1.3) ** 2, axis=1))
    _, u, feasible, *_ = mpc.solve(np.array([0.0, 0.0]), u_0=np.array([1.3]), sol_verbose=True)
    assert feasible and abs(u[0] - 1.3) < 0.05        # a warm start outside [-1, 1] survives: there are no bounds
    gp.close()


def test_sampling_mpc_update_model_subtracts_the_prior(se):
    w, gp, mpc = _pendulum_mpc(se)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, size=(50, w.n_s + w.n_u))
    y = x[:, :w.n_s] @ w.a.T + x[:, w.n_s:] @ w.b.T + 0.01
    mpc.update_model(x, y, replace_old=True)
    assert gp.x_train.shape == (50, w.n_s + w.n_u) and np.allclose(gp.y_train, 0.01)
    gp.close()


# ---------------------------------------------------------------------------------- Cautious MPC
def test_cautious_constraint_layout_matches_oracle(se):
    """CautiousMPC.generate_safety_constraints (cautious_mpc.py:337-442) on propagated Gaussian covariances: the
    device assembly (layout="cautious") against the per-candidate oracle, value by value and in the same order."""
    from oracle import score_oracle
    from safe_exploration_b200 import workloads
    w = workloads.make("C3", batch=130, n_train=250, horizon=5)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    zeros = np.zeros(w.n_s)
    k_fb1 = w.k_fb[0]
    k_fb_all = np.tile(k_fb1[None], (w.k_ff.shape[1] - 1, 1, 1))
    res = se.rollout(gp, w.p0, w.k_ff, k_fb_all, zeros, zeros, None, None, 1.0, w.a, w.b, None, True, 2)
    rng = np.random.default_rng(2)
    h_obs = rng.standard_normal((3, w.n_s))
    beta = 2.2
    g0 = np.array([score_oracle.constraints_cautious_one(res.p_all[i], res.q_all[i], w.k_ff[i], k_fb_all,
                                                          np.zeros((w.n_u, 2)), h_obs, np.zeros((3, 1)), beta)
                   for i in range(w.k_ff.shape[0])])
    n_c = 2 * w.n_u * w.k_ff.shape[1]
    u_abs = float(np.quantile(g0[:, :n_c].max(axis=1), 0.8))
    cb = np.stack((-u_abs * np.ones(w.n_u), u_abs * np.ones(w.n_u)), axis=1)
    h_obs_v = float(np.quantile(g0[:, n_c:].max(axis=1), 0.8)) * np.ones((3, 1))
    sc = se.score_rollouts(res, w.k_ff, k_fb_all, None, None, cb, h_obs, h_obs_v, cost="quadratic", wx=np.eye(w.n_s),
                           wu=np.eye(w.n_u), eps_constraints=1e-6, c_safety=beta, want_g=True, layout="cautious")
    g_o = np.array([score_oracle.constraints_cautious_one(res.p_all[i], res.q_all[i], w.k_ff[i], k_fb_all, cb, h_obs,
                                                           h_obs_v, beta) for i in range(w.k_ff.shape[0])])
    assert sc.g.shape == g_o.shape == (w.k_ff.shape[0], n_c + 3 * w.k_ff.shape[1])
    assert np.allclose(sc.g, g_o, rtol=1e-12, atol=1e-13)
    f_o = g_o.max(axis=1) < 1e-6
    assert np.array_equal(sc.feasible.astype(bool), f_o) and 0 < f_o.sum() < f_o.size
    gp.close()


def _pendulum_cautious(se, obs_bound=0.6, **kw):
    from safe_exploration_b200 import workloads
    w = workloads.make("C2", batch=8, n_train=200, horizon=5)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    env = {"l_mu": w.l_mu, "l_sigma": w.l_sigma, "h_mat_safe": w.h_mat, "h_safe": 0.5 * np.ones((2 * w.n_s, 1)),
           "lin_model": (w.a, w.b), "ctrl_bounds": np.array([[-1.0, 1.0]]), "h_mat_obs": w.h_mat,
           "h_obs": obs_bound * np.ones((2 * w.n_s, 1))}
    mpc = se.SamplingCautiousMPC(5, gp, env, 2.0, perf_trajectory=kw.pop("perf_trajectory", "mean_equivalent"),
                                 k_fb=w.k_fb[0], n_samples=512, n_iter=2, n_elite=32, wx_cost=np.eye(w.n_s),
                                 wu_cost=0.1 * np.eye(w.n_u), **kw)
    return w, gp, mpc


@pytest.mark.parametrize("perf", ["mean_equivalent", "taylor"])
def test_sampling_cautious_mpc_feasible_action_and_fallback(se, perf):
    from oracle import score_oracle
    w, gp, mpc = _pendulum_cautious(se, perf_trajectory=perf)
    with pytest.raises(AssertionError):
        mpc.get_action(np.zeros(w.n_s))                       # solver not initialised (cautious_mpc.py:223)
    mpc.init_solver()
    x0 = np.array([0.03, -0.02])
    u, code, mu_all, sigma_all, k_ff, k_fb = mpc.get_action(x0, verbose=True)
    assert code == 0 and u.shape == (w.n_u,) and mpc.n_fail == 0
    assert mu_all.shape == (6, w.n_s) and np.allclose(mu_all[0], x0) and sigma_all.shape == (5, w.n_s, w.n_s)
    # the returned plan really satisfies the reference's constraints (oracle, value by value)
    g = score_oracle.constraints_cautious_one(mu_all[1:], sigma_all, k_ff, np.tile(k_fb[None], (4, 1, 1)),
                                              mpc.ctrl_bounds, mpc.h_mat_obs, mpc.h_obs, mpc.beta_safety)
    assert np.all(g < 1e-6)
    m2, s2, v2 = mpc.f_multistep_eval(x0, k_ff, k_fb)
    assert np.allclose(m2, mu_all[1:]) and np.allclose(s2, sigma_all)
    # a custom cost with the reference's argument list picks a different plan
    mpc.init_solver(lambda mu_0, u_0, mu, sig, kff, kfb, sg: np.sum((u_0 - 0.3) ** 2, axis=1))
    u_c, code_c = mpc.get_action(x0)
    assert code_c == 0 and abs(u_c[0] - 0.3) < 0.1
    # infeasible problem: shifted old plan (exit code 1) T-1 times, then the feedback law (exit code 3)
    old = mpc.k_ff_old.copy()
    mpc.h_obs = -1.0 * np.ones_like(mpc.h_obs)
    codes = []
    for k in range(6):
        u_k, code_k = mpc.get_action(x0)
        codes.append(code_k)
        if code_k == 1:
            assert np.allclose(u_k, old[mpc.n_fail])
    assert codes == [1, 1, 1, 1, 3, 3] and np.allclose(u_k, mpc.k_fb @ x0)
    gp.close()
