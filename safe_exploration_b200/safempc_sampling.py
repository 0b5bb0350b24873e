"""Sampling SafeMPC: B candidate control sequences rolled out, scored and ranked on the GPU.

The reference solves, at every control step, one nonlinear program with CasADi/IPOPT whose objective and
constraints are symbolic functions of ONE ellipsoid trajectory (reference safe_exploration/safempc_simple.py:
163-284 ``init_solver``, :672-742 ``solve``).  On this path the same quantities are evaluated for B candidate
feed-forward sequences at once (``rollout`` -> ``score_rollouts`` -> ``best_candidate``), which is what
BASELINE.json's north star asks for ("the sequential CasADi/IPOPT path is replaced on this path by massively
parallel sampling rollouts").

* ``score_rollouts``    constraint values, feasibility and cost per candidate: generate_safety_constraints
                        (safempc_simple.py:317-392), _generate_control_constraint (:488-532),
                        eval_safety_constraints (:911-942), default generate_cost_function (:286-315)
* ``best_candidate``    arg-best on the device, across ranks if torch.distributed is initialised
* ``SamplingSafeMPC``   the SimpleSafeMPC surface (constructor arguments, ``get_action``, ``solve``,
                        ``update_model``, ``get_old_solution``, ``eval_safety_constraints``, n_fail fallback
                        logic :874-904, solution shifting :1027-1108) over a cross-entropy style sampler.

The decision variables are the reference's: u_0 and the feed-forward controls k_ff (n_safe-1, n_u); the
feedback gains k_fb are parameters (LQR of the linear prior, get_lqr_feedback :569-597), as in ``solve``.
"""
import collections
import ctypes
import warnings

import numpy as np

from . import _lib
from .gp_reachability import RolloutResult, rollout
from .ssm import BatchedGPSSM

__all__ = ["score_rollouts", "best_candidate", "SamplingSafeMPC", "ScoreResult"]

ScoreResult = collections.namedtuple("ScoreResult", ["cost", "feasible", "violation", "g"])

ATTR_NAMES_PERF = ['type_perf_traj', 'n_perf', 'r', 'perf_has_fb']                # safempc_simple.py:18-20
DEFAULT_OPT_PERF = {'type_perf_traj': 'mean_equivalent', 'n_perf': 5, 'r': 1, 'perf_has_fb': True}
ATTR_NAMES_ENV = ['l_mu', 'l_sigma', 'h_mat_safe', 'h_safe', 'lin_model', 'ctrl_bounds', 'safe_policy',
                  'h_mat_obs', 'h_obs']                                           # safempc_simple.py:22-23
DEFAULT_OPT_ENV = {'ctrl_bounds': None, 'safe_policy': None, 'lin_model': None, 'h_mat_obs': None,
                   'h_obs': None}                                                 # safempc_simple.py:24-25


def _score_params(n_s, n_u, ctrl_bounds, h_mat_obs, h_obs, h_mat_safe, h_safe, cost, wx, wu, x_ref,
                  eps_constraints, eps_noise, c_safety, layout="safempc", q_0=None, k_fb_0=None):
    keep = []

    def hp(x, shape):
        if x is None:
            return None
        arr = _lib.host_f64(x, shape)
        keep.append(arr)
        return _lib.dbl_ptr(arr)

    u_min = u_max = None
    if ctrl_bounds is not None:
        cb = _lib.host_f64(ctrl_bounds, (n_u, 2))
        u_min, u_max = hp(cb[:, 0].copy(), (n_u,)), hp(cb[:, 1].copy(), (n_u,))
    m_obs = 0 if h_mat_obs is None else int(np.shape(h_mat_obs)[0])
    if layout not in ("safempc", "cautious"):
        raise ValueError("layout must be 'safempc' or 'cautious'")
    m_safe = int(np.shape(h_mat_safe)[0]) if (layout == "safempc" and h_mat_safe is not None) else 0
    if cost not in ("exploration", "quadratic"):
        raise ValueError("cost must be 'exploration' or 'quadratic'")
    prm = _lib.ScoreParams(
        u_min, u_max, m_obs, hp(h_mat_obs, (m_obs, n_s)) if m_obs else None, hp(h_obs, (m_obs,)) if m_obs else None,
        m_safe, hp(h_mat_safe, (m_safe, n_s)) if m_safe else None, hp(h_safe, (m_safe,)) if m_safe else None,
        float(c_safety), float(eps_constraints),
        _lib.COST_EXPLORATION if cost == "exploration" else _lib.COST_QUADRATIC, float(eps_noise),
        hp(wx, (n_s, n_s)) if cost == "quadratic" else None, hp(wu, (n_u, n_u)) if cost == "quadratic" else None,
        hp(x_ref, (n_s,)) if (cost == "quadratic" and x_ref is not None) else None,
        _lib.SCORE_CAUTIOUS if layout == "cautious" else _lib.SCORE_SAFEMPC,
        hp(q_0, (n_s, n_s)) if q_0 is not None else None, hp(k_fb_0, (n_u, n_s)) if q_0 is not None else None)
    return prm, keep


def score_rollouts(res, k_ff, k_fb, h_mat_safe, h_safe, ctrl_bounds=None, h_mat_obs=None, h_obs=None,
                   cost="exploration", wx=None, wu=None, x_ref=None, eps_constraints=1e-5, eps_noise=0.0,
                   c_safety=1.0, want_g=False, layout="safempc", q_0=None, k_fb_0=None):
    """Score the candidates of a RolloutResult (device tensors or NumPy arrays; the result has the same kind).

    res        RolloutResult of ``rollout`` for k_ff (B,H,n_u) and k_fb ((H-1),n_u,n_s) or (B,H-1,n_u,n_s)
    Returns ScoreResult(cost (B,), feasible (B,) int32, violation (B,) = max constraint value, g (B,n_g) | None);
    a candidate is feasible iff every constraint value is < eps_constraints and its rollout status is 0.
    Constraint order (safempc_simple.py:317-392): [u_0 - u_max, u_min - u_0], then per step i = 0..H-2 the 2 n_u
    control distances, then per step i = 0..H-2 the m_obs obstacle distances, then the m_safe terminal distances.
    layout="cautious" (CautiousMPC.generate_safety_constraints, cautious_mpc.py:337-442): the same control constraints
    with c_safety on their support term, then the obstacle distances of ALL H states; no terminal set
    (h_mat_safe / h_safe may be None).  q_all then holds the propagated covariances.
    q_0 (n_s,n_s) with k_fb_0 (n_u,n_s): the rollout started from an ellipsoid (init_uncertainty,
    safempc_simple.py:181-199): the bound on u_0 then carries the support term of K_fb_0 Q_0 K_fb_0^T (:350).
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    on_device = torch.is_tensor(res.p_all)
    dev = res.p_all.device if on_device else torch.device("cuda", torch.cuda.current_device())

    def prep(x):
        if x is None:
            return None
        return torch.as_tensor(np.ascontiguousarray(x) if not torch.is_tensor(x) else x, dtype=torch.float64,
                               device=dev).contiguous()

    p_all, q_all, var_all, kff_d = prep(res.p_all), prep(res.q_all), prep(res.var_all), prep(k_ff)
    bsz, hor, n_s = (int(v) for v in p_all.shape)
    n_u = int(kff_d.shape[2])
    kfb_d = prep(k_fb) if hor > 1 else None
    per = (hor - 1) * n_u * n_s
    kfb_stride = 0 if (kfb_d is None or kfb_d.numel() == per) else per
    status = res.status
    if status is not None:
        status = torch.as_tensor(status, device=dev).to(torch.int32).contiguous()
    if cost == "exploration" and var_all is None:
        raise ValueError("the exploration cost needs the predictive variances (rollout(..., want_var=True))")
    if (q_0 is None) != (k_fb_0 is None):
        raise ValueError("q_0 and k_fb_0 go together")
    prm, keep = _score_params(n_s, n_u, ctrl_bounds, h_mat_obs, h_obs, h_mat_safe, h_safe, cost, wx, wu, x_ref,
                              eps_constraints, eps_noise, c_safety, layout, q_0, k_fb_0)
    n_g = lib.segp_score_num_constraints(hor, n_u, ctypes.byref(prm))
    cost_d = torch.empty((bsz,), dtype=torch.float64, device=dev)
    feas_d = torch.empty((bsz,), dtype=torch.int32, device=dev)
    viol_d = torch.empty((bsz,), dtype=torch.float64, device=dev)
    g_d = torch.empty((bsz, n_g), dtype=torch.float64, device=dev) if want_g else None
    _lib.check(lib.segp_score_rollouts(dev.index, bsz, hor, n_s, n_u, _lib.dev_ptr(p_all), _lib.dev_ptr(q_all),
                                       _lib.dev_ptr(var_all), _lib.dev_ptr(kff_d), _lib.dev_ptr(kfb_d), kfb_stride,
                                       _lib.dev_ptr(status), ctypes.byref(prm), _lib.dev_ptr(cost_d),
                                       _lib.dev_ptr(feas_d), _lib.dev_ptr(viol_d), _lib.dev_ptr(g_d),
                                       _lib.current_stream(dev)))
    if on_device:
        return ScoreResult(cost_d, feas_d, viol_d, g_d)
    return ScoreResult(cost_d.cpu().numpy(), feas_d.cpu().numpy(), viol_d.cpu().numpy(),
                       None if g_d is None else g_d.cpu().numpy())


def best_candidate(score, index_offset=0, group=None):
    """(index, cost, violation, feasible) of the best candidate: lowest cost among the feasible ones, else the
    least-violating one (feasible False).  With torch.distributed initialised the result is the best over all
    ranks (``index_offset`` = first global index of this rank's shard) and every rank returns the same tuple."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = score.cost.device if torch.is_tensor(score.cost) else torch.device("cuda", torch.cuda.current_device())
    cost_d = torch.as_tensor(score.cost, dtype=torch.float64, device=dev).contiguous()
    feas_d = torch.as_tensor(score.feasible, device=dev).to(torch.int32).contiguous()
    viol_d = torch.as_tensor(score.violation, dtype=torch.float64, device=dev).contiguous()
    idx, c, v, f = ctypes.c_long(-1), ctypes.c_double(np.inf), ctypes.c_double(np.inf), ctypes.c_int(0)
    _lib.check(lib.segp_argbest(dev.index, int(cost_d.numel()), _lib.dev_ptr(cost_d), _lib.dev_ptr(feas_d),
                                _lib.dev_ptr(viol_d), ctypes.byref(idx), ctypes.byref(c), ctypes.byref(v),
                                ctypes.byref(f), _lib.current_stream(dev)))
    local = (idx.value + index_offset if idx.value >= 0 else -1, c.value, v.value, bool(f.value))
    return _best_across_ranks(local, group)


def _best_across_ranks(local, group=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    devc = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([float(local[0]), local[1], local[2], float(local[3])], dtype=torch.float64, device=devc)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    rows = [tuple(float(x) for x in v) for v in allv]
    rows = [r for r in rows if r[0] >= 0]
    if not rows:
        return (-1, np.inf, np.inf, False)
    # feasible first, then lowest cost (feasible) / lowest violation (infeasible), then lowest index
    rows.sort(key=lambda r: (-r[3], r[1] if r[3] else r[2], r[0]))
    r = rows[0]
    return (int(r[0]), r[1], r[2], bool(r[3]))


def _dlqr(a, b, q, r):
    """Discrete LQR gain, u = -k x (reference utils.dlqr, utils.py:20-35)."""
    import scipy.linalg as sla
    x = sla.solve_discrete_are(a, b, q, r)
    return np.linalg.solve(b.T @ x @ b + r, b.T @ x @ a)


class SamplingSafeMPC(object):
    """SimpleSafeMPC with the IPOPT solve replaced by GPU sampling (safempc_simple.py:28-161 constructor).

    Parameters as the reference: ``n_safe, ssm, opt_env, wx_cost, wu_cost, beta_safety=2.5, rhc=True,
    safe_policy=None, opt_perf_trajectory={}, lin_trafo_gp_input=None, verbosity=0``; ``ssm`` must be a BatchedGPSSM.
    ``opt_perf_trajectory`` as the reference (:18-20, 153-155): ``type_perf_traj`` "mean_equivalent" | "taylor",
    ``n_perf`` (default 5; <= 1 switches the performance trajectory off), ``r`` (1: the trajectories share u_0
    only, the reference's tested case), ``perf_has_fb``.  With n_perf > 1 every candidate carries its own
    performance controls k_ff_perf (n_perf - r, n_u); both trajectories are rolled out on the GPU (the performance
    one as a Gaussian propagation, SEGP_PROP_*), the cost is the reference's default (:295-303): deviation of the
    two mean trajectories over their common steps plus the exploration term on the PERFORMANCE trajectory.
    Sampler options:
    ``n_samples`` candidates per iteration, ``n_iter`` refinement iterations (the sampling distribution is refit to
    the ``n_elite`` best feasible candidates), ``sigma0`` initial standard deviation of the control noise in units
    of the control range, ``seed``.  ``cost`` is "exploration" (the reference's default cost) or "quadratic".
    """

    def __init__(self, n_safe, ssm, opt_env, wx_cost, wu_cost, beta_safety=2.5, rhc=True, safe_policy=None,
                 opt_perf_trajectory={}, lin_trafo_gp_input=None, opts_solver=None, verbosity=0, n_samples=4096,
                 n_iter=2, n_elite=64, sigma0=0.25, cost="exploration", x_ref=None, seed=0, sigma_x0=0.05):
        if not isinstance(ssm, BatchedGPSSM):
            raise TypeError("SamplingSafeMPC needs a BatchedGPSSM")
        self.rhc = rhc
        self.ssm = ssm
        self.n_safe = int(n_safe)
        self.n_fail = self.n_safe          # no backup strategy yet (safempc_simple.py:75)
        self.n_s = ssm.num_states
        self.n_u = ssm.num_actions
        self.safe_policy = safe_policy
        for name in ATTR_NAMES_ENV:        # _set_attributes_from_dict (safempc_simple.py:1003-1016)
            if name in opt_env:
                setattr(self, name, opt_env[name])
            elif name in DEFAULT_OPT_ENV:
                setattr(self, name, DEFAULT_OPT_ENV[name])
            else:
                raise ValueError("Mandatory attribute {} missing in opt_env".format(name))
        if safe_policy is not None:
            self.safe_policy = safe_policy
        for name in ATTR_NAMES_PERF:       # _set_attributes_from_dict(ATTR_NAMES_PERF, ...), safempc_simple.py:153-154
            setattr(self, name, (opt_perf_trajectory or {}).get(name, DEFAULT_OPT_PERF[name]))
        if self.type_perf_traj not in ("mean_equivalent", "taylor"):      # _set_perf_trajectory, :1018-1025
            raise NotImplementedError("Unknown uncertainty propagation method")
        self.n_perf, self.r = int(self.n_perf), int(self.r)
        if self.n_perf > 1 and self.r != 1:
            raise NotImplementedError("coupling the performance and safety trajectories for more than one step "
                                      "(r > 1) is untested in the reference (safempc_simple.py:424-426) and not on "
                                      "the sampling path")
        self.lin_trafo_gp_input = lin_trafo_gp_input
        self.m_obs = 0 if self.h_mat_obs is None else np.shape(self.h_mat_obs)[0]
        if self.h_mat_obs is not None:
            assert np.shape(self.h_mat_obs)[1] == self.n_s, " Wrong shape of obstacle matrix"
            assert np.shape(self.h_obs) == (self.m_obs, 1), \
                " Shapes of obstacle linear inequality matrix/vector must match "
        self.m_safe, n_s_safe = np.shape(self.h_mat_safe)
        assert n_s_safe == self.n_s, " Wrong shape of safety matrix"
        assert np.shape(self.h_safe) == (self.m_safe, 1), \
            " Shapes of safety linear inequality matrix/vector must match "
        self.has_ctrl_bounds = self.ctrl_bounds is not None
        if self.has_ctrl_bounds:
            assert np.shape(self.ctrl_bounds) == (self.n_u, 2), "control bounds need to be of shape n_u x 2"
        self.wx_cost = np.asarray(wx_cost, dtype=np.float64)
        self.wu_cost = np.asarray(wu_cost, dtype=np.float64)
        self.wx_feedback = self.wx_cost
        self.wu_feedback = 1 * self.wu_cost
        self.beta_safety = beta_safety
        self.verbosity = verbosity
        self.lin_prior = False
        self.a = np.eye(self.n_s)
        self.b = np.zeros((self.n_s, self.n_u))
        if self.lin_model is not None:
            self.a, self.b = (np.asarray(m, dtype=np.float64) for m in self.lin_model)
            self.lin_prior = True
            if self.safe_policy is None:
                k = self.get_lqr_feedback().reshape(self.n_u, self.n_s)
                self.safe_policy = lambda x: np.dot(k, x)
        if self.safe_policy is None:
            warnings.warn("No SafePolicy!")
        self.n_samples, self.n_iter, self.n_elite = int(n_samples), int(n_iter), int(n_elite)
        self.sigma0 = float(sigma0)
        self.sigma_x0 = float(sigma_x0)    # opt_x0: standard deviation of the sampled initial states
        self.cost = cost
        self.x_ref = x_ref
        self._rng = np.random.default_rng(seed)
        self.cost_func = None
        self.opt_x0 = False
        self.init_uncertainty = False
        self.k_ff_perf = None              # (n_perf - r, n_u) of the last feasible solution
        self.solver_initialized = True     # nothing to build: kept for callers that check it
        self.k_ff_safe = None              # (n_safe-1, n_u) of the last feasible solution
        self.k_fb_safe_all = None          # (n_safe-1, n_u*n_s)
        self.p_safe = None                 # (n_safe, n_s)
        self.u_apply = None

    # ------------------------------------------------------------------ reference helpers
    def init_solver(self, cost_func=None, opt_x0=False, init_uncertainty=False):
        """safempc_simple.py:163-284 builds the NLP here; on this path there is nothing to compile, the three options
        are remembered:

        cost_func         replaces the default cost.  Called ONCE per iteration on the whole candidate batch with the
                          reference's argument order (:306-313) and a leading candidate axis B on every array:
                          ``cost_func(p_0 (B,n_s), u_0 (B,n_u), p_all (B,n_safe,n_s), q_all (B,n_safe,n_s,n_s),
                          k_ff_safe (B,n_safe-1,n_u), k_fb_safe (n_safe-1,n_u,n_s), sigma_safe (B,n_safe,n_s)
                          [, mu_perf (B,n_perf,n_s), sigma_perf (B,n_perf,n_s,n_s), gp_pred_sigma_perf (B,n_perf,n_s),
                          k_fb_perf, k_ff_perf (B,n_perf-1,n_u)])`` -> (B,) costs (NumPy).
        opt_x0            the initial state is a decision variable (:247-249): candidates sample it around the p_0
                          handed to ``solve`` (standard deviation ``sigma_x0``) and the best candidate's is returned.
        init_uncertainty  ``solve`` takes ``q_0`` / ``k_fb_0`` (:181-199): the rollout starts from the ellipsoid
                          (p_0, q_0) and the bound on u_0 carries its feedback term (:350)."""
        self.cost_func = cost_func
        self.opt_x0 = bool(opt_x0)
        self.init_uncertainty = bool(init_uncertainty)
        self.solver_initialized = True

    def get_lqr_feedback(self, x_0=None, u_0=None):
        """safempc_simple.py:569-597: k_fb = -dlqr(a, b, wx_feedback, wu_feedback), as a (1, n_s*n_u) row."""
        if not self.lin_prior:
            raise NotImplementedError("Cannot compute feed-back matrices without prior model")
        # the linear prior and the weights are fixed after construction: solve the Riccati equation once per value
        key = (self.a.tobytes(), self.b.tobytes(), np.asarray(self.wx_feedback).tobytes(),
               np.asarray(self.wu_feedback).tobytes())
        if getattr(self, "_lqr_key", None) != key:
            self._lqr_key = key
            self._lqr_gain = (-_dlqr(self.a, self.b, self.wx_feedback, self.wu_feedback)).reshape(
                (1, self.n_s * self.n_u))
        return self._lqr_gain.copy()

    def eval_prior(self, state, action):
        """safempc_simple.py:552-566"""
        return np.dot(state, self.a.T) + np.dot(action, self.b.T)

    def _rollout(self, p_0, k_ff, k_fb, q_0=None, k_fb_0=None):
        return rollout(self.ssm, p_0, k_ff, k_fb, self.l_mu, self.l_sigma, q_0, k_fb_0, self.beta_safety, self.a,
                       self.b, self.lin_trafo_gp_input)

    def _rollout_device(self, p_0, k_ff, k_fb, q_0=None, k_fb_0=None):
        """The same rollout with everything resident on the device: the candidates are copied into persistent CUDA
        tensors (same addresses every call, so segp_multistep replays its launches as a CUDA graph) and the result
        stays there for the scoring kernel.  Returns (RolloutResult of CUDA tensors, candidate tensor)."""
        torch = self.ssm._torch
        dev = self.ssm.device
        bsz, hor = int(k_ff.shape[0]), int(k_ff.shape[1])
        p_0 = np.asarray(p_0, dtype=np.float64)
        key = (bsz, hor, p_0.size, k_fb.shape)
        d = getattr(self, "_dev", None)
        if d is None or d["key"] != key:
            f64 = dict(dtype=torch.float64, device=dev)
            d = {"key": key, "k_ff": torch.empty((bsz, hor, self.n_u), **f64), "p_0": torch.empty((p_0.size,), **f64),
                 "k_fb": torch.empty(tuple(k_fb.shape), **f64),
                 "out": RolloutResult(torch.empty((bsz, hor, self.n_s), **f64),
                                      torch.empty((bsz, hor, self.n_s, self.n_s), **f64),
                                      torch.empty((bsz, hor, self.n_s), **f64),
                                      torch.empty((bsz,), dtype=torch.int32, device=dev))}
            self._dev = d
        d["k_ff"].copy_(torch.from_numpy(np.ascontiguousarray(k_ff)))
        d["p_0"].copy_(torch.from_numpy(np.ascontiguousarray(p_0.reshape(-1))))
        d["k_fb"].copy_(torch.from_numpy(np.ascontiguousarray(k_fb)))
        p_arg = d["p_0"] if p_0.size == self.n_s else d["p_0"].reshape(bsz, self.n_s)
        res = rollout(self.ssm, p_arg, d["k_ff"], d["k_fb"], self.l_mu, self.l_sigma, q_0, k_fb_0, self.beta_safety,
                      self.a, self.b, self.lin_trafo_gp_input, out=d["out"])
        return res, d["k_ff"]

    def _rollout_perf(self, p_0, k_ff_perf_traj, k_fb_perf):
        """The performance trajectory of every candidate: mean_equivalent_multistep / multi_step_taylor_symbolic
        (uncertainty_propagation_casadi.py:90-207) from the point p_0, controls [u_0; k_ff_perf], one feedback gain
        k_fb_perf for all steps (safempc_simple.py:430-441).  Returns the Gaussian-propagation RolloutResult:
        p_all = mu_perf (B,n_perf,n_s), q_all = sigma_perf, var_all = gp_pred_sigma_perf."""
        prop = _lib.PROP_MEAN_EQUIVALENT if self.type_perf_traj == "mean_equivalent" else _lib.PROP_TAYLOR
        zeros = np.zeros(self.n_s)
        k_fb = np.tile(np.reshape(k_fb_perf, (1, self.n_u, self.n_s)), (self.n_perf - 1, 1, 1))
        return rollout(self.ssm, p_0, k_ff_perf_traj, k_fb, zeros, zeros, None, None, 1.0, self.a, self.b,
                       self.lin_trafo_gp_input, propagation=prop)

    def _perf_cost(self, res, res_perf, eps_noise=0.0):
        """Default cost with a performance trajectory (generate_cost_function, safempc_simple.py:292-303)."""
        n_dev = min(self.n_perf, self.n_safe)
        d = res_perf.p_all[:, 1:n_dev] - res.p_all[:, 1:n_dev]
        cost = np.einsum("bti,ij,btj->b", d, 0.1 * self.wx_cost, d)
        return cost - np.sum(np.sqrt(np.sum(res_perf.var_all + eps_noise, axis=2)), axis=1)

    def _score(self, res, k_ff, k_fb, want_g=False, q_0=None, k_fb_0=None):
        return score_rollouts(res, k_ff, k_fb, self.h_mat_safe, self.h_safe, self.ctrl_bounds, self.h_mat_obs,
                              self.h_obs, cost=self.cost, wx=self.wx_cost, wu=self.wu_cost, x_ref=self.x_ref,
                              want_g=want_g, q_0=q_0, k_fb_0=k_fb_0)

    def get_safety_trajectory_openloop(self, x_0, u_0, k_fb=None, k_ff=None, q_0=None, k_fb_0=None):
        """safempc_simple.py:599-637 for one control sequence -> (p_all, q_all, var_all)."""
        k_fb = self.k_fb_safe_all if k_fb is None else k_fb
        k_ff = self.k_ff_safe if k_ff is None else k_ff
        if k_fb is None or k_ff is None:
            return None, None, None
        if q_0 is not None and k_fb_0 is None:
            k_fb_0 = self.get_lqr_feedback()
        seq = np.vstack((np.reshape(u_0, (1, self.n_u)), np.reshape(k_ff, (self.n_safe - 1, self.n_u))))
        res = self._rollout(np.reshape(x_0, (self.n_s,)), seq[None], np.reshape(k_fb, (self.n_safe - 1, self.n_u, self.n_s)),
                            None if q_0 is None else np.reshape(q_0, (self.n_s, self.n_s)),
                            None if q_0 is None else np.reshape(k_fb_0, (self.n_u, self.n_s)))
        return res.p_all[0], res.q_all[0], res.var_all[0]

    def eval_safety_constraints(self, p_all, q_all, ubg_term=0., lbg_term=-np.inf, ubg_interm=0.,
                                lbg_interm=-np.inf, terminal_only=False, eps_constraints=1e-5):
        """safempc_simple.py:911-942 (terminal constraint; like the reference, the "intermediate" values are
        evaluated on the LAST ellipsoid with the TERMINAL polytope, :101-103, 932)."""
        from .gp_reachability import lin_ellipsoid_safety_distance
        q_last = np.reshape(q_all[-1], (self.n_s, self.n_s))
        g_term = lin_ellipsoid_safety_distance(np.reshape(p_all[-1], (self.n_s, 1)), q_last, self.h_mat_safe,
                                               self.h_safe)
        feasible = bool(np.all(lbg_term - eps_constraints < g_term) and np.all(g_term < ubg_term + eps_constraints))
        if terminal_only or self.h_mat_obs is None:
            return feasible, g_term
        g_interm = g_term
        feasible_interm = bool(np.all(lbg_interm - eps_constraints < g_interm)
                               and np.all(g_interm < ubg_interm + eps_constraints))
        return feasible and feasible_interm, np.vstack((g_term, g_interm))

    # ------------------------------------------------------------------ the sampler
    def _init_controls(self):
        """Mean of the sampling distribution: the shifted previous solution if there is a fresh one
        (_get_init_controls, safempc_simple.py:1027-1108), else zeros."""
        k_fb_lqr = self.get_lqr_feedback()
        if self.n_fail == 0 and self.k_ff_safe is not None and self.n_safe > 1:
            k_ff_old = np.vstack((np.reshape(self.u_apply, (1, self.n_u)), self.k_ff_safe))   # (n_safe, n_u)
            mean = np.vstack((k_ff_old[1:], k_ff_old[-1:]))                                      # shift, repeat last
            k_fb = np.vstack((self.k_fb_safe_all[1:], self.k_fb_safe_all[-1:])) if self.n_safe > 2 \
                else np.copy(self.k_fb_safe_all)
        else:
            mean = np.zeros((self.n_safe, self.n_u))
            k_fb = np.tile(k_fb_lqr, (max(self.n_safe - 1, 1), 1))[:self.n_safe - 1]
        return mean, k_fb

    def solve(self, p_0, u_0=None, k_ff_all_0=None, k_fb_safe=None, u_perf_0=None, k_fb_perf_0=None,
              sol_verbose=False, q_0=None, k_fb_0=None):
        """One MPC step (safempc_simple.py:672-742, 742-909): sample, roll out, score, rank, fall back.

        Returns ``(x_0, u_apply, success)`` or, with sol_verbose, ``(x_0, u_apply, feasible, success,
        k_fb_safe, k_ff_all, p_safe, q_safe)`` like the reference (without the CasADi solution object).
        ``u_perf_0`` (n_perf - r, n_u) / ``k_fb_perf_0`` (n_u, n_s): initial performance controls and the feedback
        gain of the performance trajectory (default: zeros / the LQR gain, or zeros without ``perf_has_fb``);
        ``q_0`` / ``k_fb_0``: initial ellipsoid and its gain (needs ``init_solver(init_uncertainty=True)``).
        """
        p_0 = np.reshape(np.asarray(p_0, dtype=np.float64), (self.n_s,))
        if q_0 is not None and not self.init_uncertainty:
            raise ValueError("q_0 given but the solver was not initialised with init_uncertainty=True")
        if q_0 is not None:
            q_0 = np.reshape(np.asarray(q_0, dtype=np.float64), (self.n_s, self.n_s))
            k_fb_0 = np.reshape(self.get_lqr_feedback() if k_fb_0 is None else k_fb_0, (self.n_u, self.n_s))
        mean, k_fb_init = self._init_controls()
        if u_0 is not None:
            mean[0] = np.reshape(u_0, (self.n_u,))
        if k_ff_all_0 is not None and self.n_safe > 1:
            mean[1:] = np.reshape(k_ff_all_0, (self.n_safe - 1, self.n_u))
        k_fb = k_fb_init if k_fb_safe is None else np.reshape(k_fb_safe, (self.n_safe - 1, self.n_u * self.n_s))
        k_fb3 = np.reshape(k_fb, (self.n_safe - 1, self.n_u, self.n_s))
        if self.has_ctrl_bounds:
            lo, hi = self.ctrl_bounds[:, 0], self.ctrl_bounds[:, 1]
        else:
            lo, hi = -np.ones(self.n_u), np.ones(self.n_u)      # only the scale of the sampling noise
        std = np.tile(self.sigma0 * (hi - lo), (self.n_safe, 1))
        has_perf = self.n_perf > 1
        n_pf = self.n_perf - self.r if has_perf else 0
        if has_perf:
            mean_perf = np.zeros((n_pf, self.n_u))
            if u_perf_0 is not None:
                mean_perf = np.reshape(np.asarray(u_perf_0, dtype=np.float64), (n_pf, self.n_u)).copy()
            elif self.n_fail == 0 and self.k_ff_perf is not None and n_pf > 0:
                mean_perf = np.vstack((self.k_ff_perf[1:], self.k_ff_perf[-1:]))      # shifted previous solution
            std_perf = np.tile(self.sigma0 * (hi - lo), (n_pf, 1))
            if k_fb_perf_0 is None:
                k_fb_perf_0 = self.get_lqr_feedback() if (self.perf_has_fb and self.lin_prior) else \
                    np.zeros((self.n_u, self.n_s))
            k_fb_perf = np.reshape(k_fb_perf_0, (self.n_u, self.n_s))
        mean_x, std_x = p_0.copy(), np.full(self.n_s, self.sigma_x0)
        best = None
        for _ in range(max(self.n_iter, 1)):
            cand = mean[None] + std[None] * self._rng.standard_normal((self.n_samples, self.n_safe, self.n_u))
            cand[0] = mean                                      # the current mean is always a candidate
            # u_0 is applied at a point (no feedback term), so clipping it to the bounds loses nothing; the later
            # feed-forward terms need head-room for the feedback part and are left to the constraint check.  Without
            # control bounds nothing is clipped (the reference leaves u_0 free then).
            if self.has_ctrl_bounds and q_0 is None:
                cand[:, 0] = np.clip(cand[:, 0], lo, hi)
            if self.opt_x0:
                x_cand = mean_x[None] + std_x[None] * self._rng.standard_normal((self.n_samples, self.n_s))
                x_cand[0] = mean_x
            else:
                x_cand = p_0
            # Default problem (no performance trajectory, no Python cost function): candidates, rollout results and
            # scores stay on the device -- persistent buffers, so the rollout replays as one CUDA graph -- and only
            # the three (B,) score vectors and the best candidate's trajectory come back to the host.
            on_dev = not has_perf and self.cost_func is None
            if on_dev:
                res, cand_d = self._rollout_device(x_cand, cand, k_fb3, q_0, k_fb_0)
                sc_d = self._score(res, cand_d, self._dev["k_fb"], q_0=q_0, k_fb_0=k_fb_0)
                sc = ScoreResult(sc_d.cost.cpu().numpy(), sc_d.feasible.cpu().numpy(), sc_d.violation.cpu().numpy(), None)
            else:
                sc_d = None
                res = self._rollout(x_cand, cand, k_fb3, q_0, k_fb_0)
                sc = self._score(res, cand, k_fb3, q_0=q_0, k_fb_0=k_fb_0)
            cost_v = sc.cost
            res_perf = cand_perf = None
            if has_perf:
                cand_perf = mean_perf[None] + std_perf[None] * self._rng.standard_normal((self.n_samples, n_pf, self.n_u))
                cand_perf[0] = mean_perf
                if self.has_ctrl_bounds:      # plain bounds, no feedback term (safempc_simple.py:476-483)
                    cand_perf = np.clip(cand_perf, lo, hi)
                seq_perf = np.concatenate((cand[:, :1], cand_perf), axis=1)          # [u_0; k_ff_perf], r = 1
                res_perf = self._rollout_perf(x_cand, seq_perf, k_fb_perf)
                cost_v = self._perf_cost(res, res_perf)
                ok = np.all(np.isfinite(res_perf.p_all.reshape(self.n_samples, -1)), axis=1) & (res_perf.status == 0)
                sc = ScoreResult(cost_v, sc.feasible * ok.astype(sc.feasible.dtype), sc.violation, sc.g)
            if self.cost_func is not None:
                x_b = x_cand if self.opt_x0 else np.tile(p_0[None], (self.n_samples, 1))
                args = [x_b, cand[:, 0], res.p_all, res.q_all, cand[:, 1:], k_fb3, res.var_all]
                if has_perf:
                    args += [res_perf.p_all, res_perf.q_all, res_perf.var_all, k_fb_perf, seq_perf[:, 1:]]
                cost_v = np.asarray(self.cost_func(*args), dtype=np.float64).reshape(self.n_samples)
                sc = ScoreResult(cost_v, sc.feasible * np.isfinite(cost_v).astype(sc.feasible.dtype), sc.violation, sc.g)
            idx, cost, viol, feas = best_candidate(sc_d if on_dev else sc)   # arg-best on the device-resident scores
            if idx >= 0 and (best is None or (feas, -cost if feas else -viol) > (best[3], -best[1] if best[3] else -best[2])):
                p_best = res.p_all[idx].cpu().numpy() if on_dev else res.p_all[idx].copy()
                q_best = res.q_all[idx].cpu().numpy() if on_dev else res.q_all[idx].copy()
                best = (cand[idx].copy(), cost, viol, feas, p_best, q_best,
                        None if cand_perf is None else cand_perf[idx].copy(),
                        x_cand[idx].copy() if self.opt_x0 else p_0)
            order = np.lexsort((np.where(sc.feasible > 0, sc.cost, sc.violation), -sc.feasible))
            top = order[:min(self.n_elite, self.n_samples)]
            elite = cand[top]
            mean = elite.mean(axis=0)
            std = np.maximum(elite.std(axis=0), 1e-3 * (hi - lo))
            if has_perf and n_pf > 0:
                mean_perf = cand_perf[top].mean(axis=0)
                std_perf = np.maximum(cand_perf[top].std(axis=0), 1e-3 * (hi - lo))
            if self.opt_x0:
                mean_x = x_cand[top].mean(axis=0)
                std_x = np.maximum(x_cand[top].std(axis=0), 1e-3 * self.sigma_x0)
        feasible = bool(best is not None and best[3])
        success = True
        k_ff_all = p_safe = q_safe = k_fb_out = None
        x_0 = p_0
        if feasible:
            seq, _, _, _, p_safe, q_safe, k_ff_perf, x_0 = best
            u_apply = seq[0].copy()
            k_ff_all = seq[1:].copy()
            k_fb_out = np.copy(k_fb)
            self.n_fail = 0
            if self.rhc:
                self.k_ff_safe = k_ff_all
                self.p_safe = p_safe
                self.k_fb_safe_all = k_fb_out
                self.u_apply = u_apply
                self.k_ff_perf = k_ff_perf
        else:
            self.n_fail += 1
            if self.n_fail >= self.n_safe:
                # too many infeasible steps: safe controller (safempc_simple.py:887-893)
                u_apply = np.reshape(self.safe_policy(p_0), (self.n_u,))
                k_ff_all = u_apply
            else:
                u_apply = np.reshape(self.get_old_solution(p_0), (self.n_u,))
                k_ff_all = u_apply
        if sol_verbose:
            return np.reshape(x_0, (self.n_s, 1)), u_apply, feasible, success, k_fb_out, k_ff_all, p_safe, q_safe
        return np.reshape(x_0, (self.n_s, 1)), u_apply, success

    def get_action(self, x0_mu, lqr_only=False, sol_verbose=False):
        """safempc_simple.py:639-670"""
        if lqr_only:
            return self.safe_policy(x0_mu), False
        x0 = np.asarray(x0_mu, dtype=np.float64).reshape(self.n_s)
        if sol_verbose:
            _, u_apply, feasible, success, k_fb_apply, k_ff_all, p_all, q_all = self.solve(x0, sol_verbose=True)
            return u_apply.reshape(self.n_u, ), feasible, success, k_fb_apply, k_ff_all, p_all, q_all
        _, u_apply, success = self.solve(x0)
        return u_apply.reshape(self.n_u, ), success

    def get_old_solution(self, x, k=None, get_ctrl_traj=False):
        """Shift the last feasible solution (safempc_simple.py:944-1001): u = k_ff[k-1] + K_fb[k-1] (x - p[k-1])."""
        if self.n_fail > self.n_safe or self.k_ff_safe is None:
            warnings.warn("There are no previous solution to be applied. Returning None")
            return None
        k = self.n_fail if k is None else k
        if k < 1:
            warnings.warn("Have to shift at least one timestep back")
            return None
        k_fb_old = np.reshape(self.k_fb_safe_all[k - 1], (self.n_u, self.n_s))
        k_ff = self.k_ff_safe[k - 1]
        p_safe = self.p_safe[k - 1]
        u_apply = k_ff + k_fb_old @ (np.reshape(x, (self.n_s,)) - p_safe)       # utils.feedback_ctrl
        if get_ctrl_traj:
            if k < self.n_safe:
                return (u_apply, self.k_fb_safe_all[k:], np.vstack((u_apply, self.k_ff_safe[k + 1:])), self.p_safe[k:])
            return u_apply, None, u_apply, None
        return u_apply

    def update_model(self, x, y, opt_hyp=False, replace_old=True, reinitialize_solver=True):
        """safempc_simple.py:1111-1140: the GP learns the residual to the linear prior."""
        x = np.asarray(x, dtype=np.float64)
        n_train = x.shape[0]
        x_s = x[:, :self.n_s].reshape((n_train, self.n_s))
        x_u = x[:, self.n_s:].reshape((n_train, self.n_u))
        y_prior = self.eval_prior(x_s, x_u)
        x_trafo = x_s if self.lin_trafo_gp_input is None else x_s @ np.asarray(self.lin_trafo_gp_input).T
        self.ssm.update_model(np.hstack((x_trafo, x_u)), np.asarray(y, dtype=np.float64) - y_prior, opt_hyp,
                              replace_old)
