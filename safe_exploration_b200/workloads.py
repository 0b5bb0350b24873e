"""Synthetic workloads C1..C5 of BASELINE.json, frozen by seed.

Pure NumPy/SciPy host code that only *generates inputs* (training data, hyper-parameters, linear prior,
LQR feedback gain, candidate control sequences); shared by bench.py and the tests so the GPU path and the
CPU oracle see identical inputs.  The dynamics are restated from the equations of motion the reference
integrates (safe_exploration/environments.py:363-388 inverted pendulum, :875-896 cart-pole) as one explicit
Euler step; the GP learns the residual to the linearisation at the origin, as the reference's SafeMPC does
(safempc_simple.py:150-160, lin_model).
"""
import collections

import numpy as np
import scipy.linalg as sla

Workload = collections.namedtuple("Workload", [
    "name", "n_s", "n_u", "n_train", "horizon", "batch", "x_train", "y_train", "kern_types", "hyp", "a", "b",
    "k_fb", "p0", "k_ff", "l_mu", "l_sigma", "c_safety", "h_mat", "h_vec"])

# (name, system, n_s, n_u, N, H, B_total, gpus, kernel)
CONFIGS = collections.OrderedDict([
    ("C1", ("pendulum", 2, 1, 50, 5, 1, 1, "rbf")),
    ("C2", ("pendulum", 2, 1, 500, 10, 4096, 1, "rbf")),
    ("C3", ("cartpole", 4, 1, 2000, 15, 16384, 1, "mat52")),
    ("C4", ("cartpole", 4, 1, 5000, 20, 65536, 8, "rbf")),
    ("C5", ("synthetic10", 10, 3, 10000, 30, 131072, 8, "rbf")),
])


def pendulum_step(x, u, dt=0.05, length=0.5, mass=0.15, grav=9.82, fric=0.0):
    """One Euler step of the inverted pendulum, state [d_theta, theta] (environments.py:363-388)."""
    inertia = mass * length ** 2
    dz0 = grav / length * np.sin(x[:, 1]) + u[:, 0] / inertia - fric / inertia * x[:, 0]
    dz1 = x[:, 0]
    return x + dt * np.stack((dz0, dz1), axis=1)


def pendulum_linear(dt=0.05, length=0.5, mass=0.15, grav=9.82, fric=0.0):
    inertia = mass * length ** 2
    a = np.eye(2) + dt * np.array([[-fric / inertia, grav / length], [1.0, 0.0]])
    b = dt * np.array([[1.0 / inertia], [0.0]])
    return a, b


def cartpole_step(x, u, dt=0.02, length=0.5, m=0.5, big_m=0.5, fric=0.1, grav=9.82):
    """One Euler step of the cart-pole, state [pos, vel, theta, omega] (environments.py:875-896)."""
    v, th, om = x[:, 1], x[:, 2], x[:, 3]
    f = u[:, 0]
    det = length * (big_m + m * np.sin(th) ** 2)
    dz1 = (f - m * length * om ** 2 * np.sin(th) - fric * om * np.cos(th)
           + 0.5 * m * grav * length * np.sin(2 * th)) * length / det
    dz3 = (f * np.cos(th) - 0.5 * m * length * om ** 2 * np.sin(2 * th)
           - fric * (m + big_m) * om / (m * length) + (m + big_m) * grav * np.sin(th)) / det
    return x + dt * np.stack((v, dz1, om, dz3), axis=1)


def cartpole_linear(dt=0.02, length=0.5, m=0.5, big_m=0.5, fric=0.1, grav=9.82):
    """Linearisation of cartpole_step at the origin (finite differences would do; this is analytic)."""
    det = length * big_m
    jac = np.zeros((4, 4))
    jac[0, 1] = 1.0
    jac[1, 2] = (m * grav * length) * length / det
    jac[1, 3] = -fric * length / det
    jac[2, 3] = 1.0
    jac[3, 2] = (m + big_m) * grav / det
    jac[3, 3] = -fric * (m + big_m) / (m * length) / det
    bu = np.array([[0.0], [length / det], [0.0], [1.0 / det]])
    return np.eye(4) + dt * jac, dt * bu


def lqr_gain(a, b, q=None, r=None):
    """Infinite-horizon discrete LQR; returns k_fb with u = k_fb x (i.e. minus the gain of the reference's
    dlqr, utils.py:20-35)."""
    n, m = b.shape
    q = np.eye(n) if q is None else q
    r = np.eye(m) if r is None else r
    x = sla.solve_discrete_are(a, b, q, r)
    k = np.linalg.solve(b.T @ x @ b + r, b.T @ x @ a)
    return -k


def make(name, batch=None, n_train=None, horizon=None, seed_offset=0, kern=None):
    """Build workload `name` (C1..C5).  `batch` overrides the number of candidate sequences (per-GPU shards,
    small parity cases); `n_train` / `horizon` shrink the model for CPU-sized parity tests; `kern` swaps the kernel
    ("lin_rbf" / "lin_mat52": the composite kernels of the reference's journal configs,
    experiments/journal_experiment_configs/defaultconfig_episode.py:39, with hyper-parameters in the reference's own
    key layout)."""
    system, n_s, n_u, n_cfg, h_cfg, b_cfg, _, kern_cfg = CONFIGS[name]
    kern = kern or kern_cfg
    n = int(n_train or n_cfg)
    hor = int(horizon or h_cfg)
    bsz = int(batch or b_cfg)
    dim = n_s + n_u
    rng0 = np.random.default_rng(0 + seed_offset)
    rng1 = np.random.default_rng(1 + seed_offset)
    rng2 = np.random.default_rng(2 + seed_offset)
    x_train = rng0.uniform(-1.0, 1.0, size=(n, dim))
    xs, us = x_train[:, :n_s], x_train[:, n_s:]
    if system == "pendulum":
        a, b = pendulum_linear()
        nxt = pendulum_step(xs, us)
        plant_noise = np.array([0.01, 0.01]) ** 2
        u_lim = 1.0
        lqr_r = 1.0
        lipschitz = 1e-3
    elif system == "cartpole":
        a, b = cartpole_linear()
        nxt = cartpole_step(xs, 4.0 * us)            # actions scaled to the +-4 N range of the reference
        b = 4.0 * b
        plant_noise = np.array([0.02, 0.05, 0.02, 0.05]) ** 2
        u_lim = 1.0
        lqr_r = 10.0
        lipschitz = 1e-4
    else:
        rng3 = np.random.default_rng(3 + seed_offset)
        a = 0.8 * np.eye(n_s) + 0.02 * rng3.standard_normal((n_s, n_s))
        b = 0.3 * rng3.standard_normal((n_s, n_u))
        w1 = rng3.standard_normal((dim, 16)) / np.sqrt(dim)
        w2 = 0.05 * rng3.standard_normal((16, n_s))
        nxt = xs @ a.T + us @ b.T + np.tanh(x_train @ w1) @ w2
        plant_noise = np.full(n_s, 1e-2)
        u_lim = 1.0
        lqr_r = 1.0
        lipschitz = 1e-4
    y_train = nxt - (xs @ a.T + us @ b.T) + np.sqrt(plant_noise)[None, :] * rng0.standard_normal((n, n_s))
    if name == "C1":
        # GPy defaults, what gp.train(..., opt_hyp=False) yields (reference test/test_safempc.py:56-69)
        hyp = [{"lengthscale": np.ones(dim), "variance": 1.0, "noise": 1.0} for _ in range(n_s)]
        l_mu = np.array([0.05, 0.02])
        l_sigma = np.array([0.05, 0.02])
    else:
        hyp = [{"lengthscale": rng1.uniform(0.8, 2.0, size=dim), "variance": float(rng1.uniform(0.5, 1.5)),
                "noise": 1e-2} for _ in range(n_s)]
        # Screened with the float64 oracle so every Q_t stays finite over the horizon: the
        # remainder term grows like (l_mu lambda_max(Q (I + K^T K)))^2, so the cart-pole / 10-D cases (|K| ~ 10,
        # H = 15..30) use 1e-4 and a 20 ms step where the pendulum uses the reference test's 1e-3
        # (test/test_gp_reachability_casadi.py:55-56).
        l_mu = np.full(n_s, lipschitz)
        l_sigma = np.full(n_s, lipschitz)
    if kern in ("lin_rbf", "lin_mat52"):
        st = kern[4:]
        hyp = [{"prod.%s.lengthscale" % st: np.array([float(np.mean(h["lengthscale"]))]),
                "prod.%s.variance" % st: float(h["variance"]), "prod.linear.variances": np.array([1.0]),
                "linear.variances": np.full(dim, 0.05), "noise": h["noise"]} for h in hyp]
    k_gain = lqr_gain(a, b, r=lqr_r * np.eye(n_u))
    k_fb = np.tile(k_gain[None], (max(hor - 1, 1), 1, 1))[:max(hor - 1, 0)]
    p0 = 0.05 * rng2.standard_normal(n_s)
    # Candidate control sequences: the ellipsoid centre follows u = k_ff exactly (gp_reachability.py:93-94), so
    # open-loop noise on an unstable plant would leave the data region within a few steps.  Candidates are
    # therefore LQR tracking controls of the linear prior plus exploration noise, k_ff[t] = K p_nom[t] + delta[t],
    # p_nom[t+1] = (A + B K) p_nom[t] + B delta[t]  (what a sampling MPC would propose around its nominal plan).
    delta = 0.05 * rng2.standard_normal((bsz, hor, n_u))
    k_ff = np.empty((bsz, hor, n_u))
    p_nom = np.tile(p0[None], (bsz, 1))
    for t in range(hor):
        k_ff[:, t] = np.clip(p_nom @ k_gain.T + delta[:, t], -u_lim, u_lim)
        p_nom = p_nom @ a.T + k_ff[:, t] @ b.T
    # a box polytope |x_i| <= 1 for scoring (safety distance) demonstrations
    h_mat = np.vstack((np.eye(n_s), -np.eye(n_s)))
    h_vec = np.ones((2 * n_s, 1))
    return Workload(name, n_s, n_u, n, hor, bsz, x_train, y_train, [kern] * n_s, hyp, a, b, k_fb, p0, k_ff, l_mu,
                    l_sigma, 2.0, h_mat, h_vec)


def flop_per_step(n_s, n_u, n_train):
    """Algorithmic flop of one (trajectory, step): SURVEY.md section 8d, F_step = n_s [N^2 + N (5 D + 9)]."""
    dim = n_s + n_u
    return n_s * (n_train ** 2 + n_train * (5 * dim + 9))
