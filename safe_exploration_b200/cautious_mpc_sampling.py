"""Sampling Cautious MPC: the reference's CautiousMPC with the CasADi/IPOPT solve replaced by GPU sampling.

Mirror of reference safe_exploration/cautious_mpc.py:23-479 for the hot path: the Gaussian
state distribution is propagated T steps under ``u_t = k_ff[t] + K_fb (x_t - mu_t)`` by the Taylor or the
mean-equivalent scheme (``perf_trajectory``, cautious_mpc.py:127-134, 160-162), chance constraints are evaluated on the
propagated covariances (``generate_safety_constraints`` :337-395, ``_generate_control_constraint`` :397-442) and a cost
is minimised.  Here all of it is evaluated for B candidate feed-forward sequences in one batch:

    rollout(..., propagation=TAYLOR | MEAN_EQUIVALENT)  ->  score_rollouts(..., layout="cautious")  ->  best_candidate

inside a cross-entropy style sampler (the same one SamplingSafeMPC uses).  Kept from the reference: the constructor
arguments, ``init_solver(cost_func)``, ``get_action(x0_mu, verbose)`` with its exit codes (0 feasible, 1 shifted old
solution, 3 feedback law only; 2 = "solver crashed" cannot happen here), the RHC warm start (``_get_init_controls``
:320-335), the fall-back logic (``_get_old_solution`` :290-318) and ``update_model`` (:444-479).

``cost_func`` is a Python callable with the reference's argument list
``cost_func(mu_0, u_0, mu_all, sigma_all, k_ff, k_fb, sigma_g)`` (cautious_mpc.py:172), evaluated on NumPy arrays that
carry a leading candidate axis: ``mu_0 (n_s,)``, ``u_0 (B,n_u)``, ``mu_all (B,T,n_s)``, ``sigma_all (B,T,n_s,n_s)``,
``k_ff (B,T-1,n_u)``, ``k_fb (n_u,n_s)``, ``sigma_g (B,T,n_s)`` -> ``(B,)``.  Without one, the quadratic cost
``sum_t (mu_t - x_ref)^T Wx (mu_t - x_ref) + u_t^T Wu u_t`` is evaluated on the device.
Not on this path: ``opt_x0`` (the initial state as a decision variable).
"""
import warnings

import numpy as np

from .gp_reachability import RolloutResult, rollout
from .safempc_sampling import ScoreResult, best_candidate, score_rollouts
from .ssm import BatchedGPSSM

__all__ = ["SamplingCautiousMPC"]

ATTR_NAMES_ENV = ['l_mu', 'l_sigma', 'h_mat_safe', 'h_safe', 'lin_model', 'ctrl_bounds', 'safe_policy',
                  'h_mat_obs', 'h_obs']                                           # cautious_mpc.py:17-18
DEFAULT_OPT_ENV = {'ctrl_bounds': None, 'safe_policy': None, 'lin_model': None, 'h_mat_obs': None,
                   'h_obs': None}                                                 # cautious_mpc.py:19-20
_PROPAGATION = {"taylor": 1, "mean_equivalent": 2}


class SamplingCautiousMPC(object):
    """CautiousMPC(T, gp, env_options, beta_safety, lin_trafo_gp_input=None, perf_trajectory="mean_equivalent",
    k_fb=None) -- cautious_mpc.py:57-109 -- plus the sampler options ``n_samples, n_iter, n_elite, sigma0, seed`` and
    the default-cost weights ``wx_cost, wu_cost, x_ref``."""

    def __init__(self, T, gp, env_options, beta_safety, lin_trafo_gp_input=None, perf_trajectory="mean_equivalent",
                 k_fb=None, n_samples=4096, n_iter=2, n_elite=64, sigma0=0.25, seed=0, wx_cost=None, wu_cost=None,
                 x_ref=None):
        if not isinstance(gp, BatchedGPSSM):
            raise TypeError("SamplingCautiousMPC needs a BatchedGPSSM")
        self.T = int(T)
        self.gp = gp
        self.n_s = gp.n_s_out
        self.n_u = gp.n_u
        self.n_fail = self.T - 1                                   # cautious_mpc.py:65
        self.beta_safety = beta_safety
        self._set_perf_trajectory(perf_trajectory)
        self.opt_x0 = False
        self.cost_func = None
        self.lin_prior = False
        for name in ATTR_NAMES_ENV:                                # _set_attributes_from_dict, cautious_mpc.py:111-125
            if name in env_options:
                setattr(self, name, env_options[name])
            elif name in DEFAULT_OPT_ENV:
                setattr(self, name, DEFAULT_OPT_ENV[name])
            else:
                raise ValueError("Mandatory attribute {} missing in env_options".format(name))
        self.lin_trafo_gp_input = lin_trafo_gp_input
        if self.h_mat_obs is None:
            self.m_obs = 0
        else:
            self.m_obs, n_s_obs = np.shape(self.h_mat_obs)
            assert n_s_obs == self.n_s, " Wrong shape of obstacle matrix"
            assert np.shape(self.h_obs) == (self.m_obs, 1), \
                " Shapes of obstacle linear inequality matrix/vector must match "
        self.has_ctrl_bounds = self.ctrl_bounds is not None
        if self.has_ctrl_bounds:
            assert np.shape(self.ctrl_bounds) == (self.n_u, 2), "control bounds need to be of shape n_u x 2"
        self.m_safe, n_s_safe = np.shape(self.h_mat_safe)
        assert n_s_safe == self.n_s, " Wrong shape of safety matrix"
        assert np.shape(self.h_safe) == (self.m_safe, 1), \
            " Shapes of safety linear inequality matrix/vector must match."
        self.a = np.eye(self.n_s)
        self.b = np.zeros((self.n_s, self.n_u))
        if self.lin_model is not None:
            self.a, self.b = (np.asarray(m, dtype=np.float64) for m in self.lin_model)
            self.lin_prior = True
        self.eval_prior = lambda x, u: np.dot(x, self.a.T) + np.dot(u, self.b.T)
        self.k_fb = None if k_fb is None else np.asarray(k_fb, dtype=np.float64).reshape(self.n_u, self.n_s)
        self.k_ff_old = None
        self.solver_initialized = False
        self.n_samples, self.n_iter, self.n_elite = int(n_samples), int(n_iter), int(n_elite)
        self.sigma0 = float(sigma0)
        self.wx_cost = None if wx_cost is None else np.asarray(wx_cost, dtype=np.float64)
        self.wu_cost = None if wu_cost is None else np.asarray(wu_cost, dtype=np.float64)
        self.x_ref = x_ref
        self._rng = np.random.default_rng(seed)

    def _set_perf_trajectory(self, name):
        """cautious_mpc.py:127-134"""
        if name not in _PROPAGATION:
            raise NotImplementedError("Unknown uncertainty propagation method")
        self.perf_trajectory = name
        self._prop_mode = _PROPAGATION[name]

    # ------------------------------------------------------------------ solver surface
    def init_solver(self, cost_func=None, opt_x0=False):
        """cautious_mpc.py:136-200 builds the NLP here; on this path only the cost is recorded."""
        if opt_x0:
            raise NotImplementedError("opt_x0 is not on the sampling path")
        if self.k_fb is None:
            raise ValueError("a feedback gain k_fb (n_u x n_s) is required")
        if cost_func is None:
            cost_func = self.cost_func
        if cost_func is None and (self.wx_cost is None or self.wu_cost is None):
            raise ValueError("either cost_func or the quadratic weights wx_cost / wu_cost are required")
        self.cost_func = cost_func
        self.solver_initialized = True

    def f_multistep_eval(self, mu_0, k_ff, k_fb):
        """The reference's casadi.Function of the same name (cautious_mpc.py:164-165), for one control sequence
        (T,n_u) or a batch (B,T,n_u): (mu_all, sigma_all, sigma_g)."""
        k_ff = np.asarray(k_ff, dtype=np.float64)
        single = k_ff.ndim == 2
        res = self._propagate(np.reshape(mu_0, (self.n_s,)), k_ff[None] if single else k_ff, k_fb)
        if single:
            return res.p_all[0], res.q_all[0], res.var_all[0]
        return res.p_all, res.q_all, res.var_all

    def _propagate(self, mu_0, k_ff, k_fb):
        k_fb_all = np.tile(np.reshape(k_fb, (1, self.n_u, self.n_s)), (max(self.T - 1, 1), 1, 1))[:self.T - 1]
        zeros = np.zeros(self.n_s)
        return rollout(self.gp, mu_0, k_ff, k_fb_all, zeros, zeros, None, None, 1.0, self.a, self.b,
                       self.lin_trafo_gp_input, True, self._prop_mode)

    def _propagate_device(self, mu_0, k_ff, k_fb):
        """_propagate with the candidates copied into persistent CUDA tensors and the result left on the device for the
        scoring kernel (same addresses every call: the launches replay as a CUDA graph).  Returns (result, k_ff tensor)."""
        torch = self.gp._torch
        dev = self.gp.device
        bsz = int(k_ff.shape[0])
        d = getattr(self, "_dev", None)
        if d is None or d["key"] != bsz:
            f64 = dict(dtype=torch.float64, device=dev)
            d = {"key": bsz, "k_ff": torch.empty((bsz, self.T, self.n_u), **f64), "mu_0": torch.empty((self.n_s,), **f64),
                 "k_fb": torch.empty((max(self.T - 1, 0), self.n_u, self.n_s), **f64),
                 "out": RolloutResult(torch.empty((bsz, self.T, self.n_s), **f64),
                                      torch.empty((bsz, self.T, self.n_s, self.n_s), **f64),
                                      torch.empty((bsz, self.T, self.n_s), **f64),
                                      torch.empty((bsz,), dtype=torch.int32, device=dev))}
            self._dev = d
        k_fb_all = np.tile(np.reshape(k_fb, (1, self.n_u, self.n_s)), (max(self.T - 1, 1), 1, 1))[:self.T - 1]
        d["k_ff"].copy_(torch.from_numpy(np.ascontiguousarray(k_ff)))
        d["mu_0"].copy_(torch.from_numpy(np.ascontiguousarray(np.reshape(mu_0, (self.n_s,)))))
        d["k_fb"].copy_(torch.from_numpy(np.ascontiguousarray(k_fb_all)))
        zeros = np.zeros(self.n_s)
        res = rollout(self.gp, d["mu_0"], d["k_ff"], d["k_fb"], zeros, zeros, None, None, 1.0, self.a, self.b,
                      self.lin_trafo_gp_input, True, self._prop_mode, out=d["out"])
        return res, d["k_ff"]

    def generate_safety_constraints(self, p_all, q_all, u_0, k_fb, k_ff):
        """cautious_mpc.py:337-395 evaluated numerically for a batch: p_all (B,T,n_s), q_all (B,T,n_s,n_s), u_0 (B,n_u),
        k_fb (n_u,n_s), k_ff (B,T-1,n_u) -> ScoreResult with g (B,n_g) in the reference's order and the feasibility
        test of _get_solution (:267-271: every g within its bounds up to feas_tol)."""
        from .gp_reachability import RolloutResult
        seq = np.concatenate((np.reshape(u_0, (-1, 1, self.n_u)), np.reshape(k_ff, (-1, self.T - 1, self.n_u))), axis=1)
        res = RolloutResult(np.asarray(p_all), np.asarray(q_all), None, None)
        return self._score(res, seq, k_fb, want_g=True)

    def _score(self, res, seq, k_fb, want_g=False):
        k_fb_all = np.tile(np.reshape(k_fb, (1, self.n_u, self.n_s)), (max(self.T - 1, 1), 1, 1))[:self.T - 1]
        quad = self.cost_func is None
        wx = self.wx_cost if quad else np.zeros((self.n_s, self.n_s))
        wu = self.wu_cost if quad else np.zeros((self.n_u, self.n_u))
        return score_rollouts(res, seq, k_fb_all, None, None, self.ctrl_bounds, self.h_mat_obs, self.h_obs,
                              cost="quadratic", wx=wx, wu=wu, x_ref=self.x_ref, eps_constraints=1e-6,
                              c_safety=self.beta_safety, want_g=want_g, layout="cautious")

    def _get_init_controls(self):
        """cautious_mpc.py:320-335"""
        if self.n_fail == 0 and self.k_ff_old is not None:
            k_ff_old = np.copy(self.k_ff_old)
            k_ff_0 = np.vstack((k_ff_old[1:, :], k_ff_old[-1, :]))
        else:
            k_ff_0 = np.zeros((self.T, self.n_u))
        return k_ff_0, self.k_fb

    def get_action(self, x0_mu, verbose=False):
        """cautious_mpc.py:202-288.  Returns (u_apply, exit_code) or, verbose and feasible,
        (u_apply, exit_code, vstack(mu_0, mu_all), sigma_all, k_ff, k_fb)."""
        assert self.solver_initialized, "Need to initialize the solver first!"
        mu_0 = np.reshape(np.asarray(x0_mu, dtype=np.float64), (self.n_s,))
        mean, k_fb_0 = self._get_init_controls()
        if self.has_ctrl_bounds:
            lo, hi = np.asarray(self.ctrl_bounds)[:, 0], np.asarray(self.ctrl_bounds)[:, 1]
        else:
            lo, hi = -np.ones(self.n_u), np.ones(self.n_u)
        std = np.tile(self.sigma0 * (hi - lo), (self.T, 1))
        best = None
        for _ in range(max(self.n_iter, 1)):
            cand = mean[None] + std[None] * self._rng.standard_normal((self.n_samples, self.T, self.n_u))
            cand[0] = mean
            if self.has_ctrl_bounds:                      # u_0 is applied at a point: plain bounds (:363-367); without
                cand[:, 0] = np.clip(cand[:, 0], lo, hi)  # control bounds the reference leaves it free: no clipping
            # default cost: candidates, propagated trajectories and scores stay on the device; only the three (B,) score
            # vectors and the best candidate's trajectory come back (a Python cost function needs host arrays)
            on_dev = self.cost_func is None
            if on_dev:
                res, cand_d = self._propagate_device(mu_0, cand, k_fb_0)
                sc_d = self._score(res, cand_d, k_fb_0)
                sc = ScoreResult(sc_d.cost.cpu().numpy(), sc_d.feasible.cpu().numpy(), sc_d.violation.cpu().numpy(), None)
            else:
                sc_d = None
                res = self._propagate(mu_0, cand, k_fb_0)
                sc = self._score(res, cand, k_fb_0)
                cost = np.asarray(self.cost_func(mu_0, cand[:, 0], res.p_all, res.q_all, cand[:, 1:], k_fb_0,
                                                 res.var_all), dtype=np.float64).reshape(-1)
                sc = ScoreResult(cost, sc.feasible, sc.violation, sc.g)
            idx, cost_b, viol_b, feas_b = best_candidate(sc_d if on_dev else sc)
            if idx >= 0 and (best is None or (feas_b, -cost_b if feas_b else -viol_b) >
                             (best[3], -best[1] if best[3] else -best[2])):
                p_best = res.p_all[idx].cpu().numpy() if on_dev else res.p_all[idx].copy()
                q_best = res.q_all[idx].cpu().numpy() if on_dev else res.q_all[idx].copy()
                best = (cand[idx].copy(), cost_b, viol_b, feas_b, p_best, q_best)
            order = np.lexsort((np.where(sc.feasible > 0, sc.cost, sc.violation), -sc.feasible))
            elite = cand[order[:min(self.n_elite, self.n_samples)]]
            mean = elite.mean(axis=0)
            std = np.maximum(elite.std(axis=0), 1e-3 * (hi - lo))
        if best is not None and best[3]:
            k_ff = best[0]
            self.k_ff_old = k_ff
            self.n_fail = 0
            u_apply = np.array(k_ff[0, :]).reshape(self.n_u, )
            if verbose:
                return u_apply, 0, np.vstack((mu_0[None, :], best[4])), best[5], k_ff, k_fb_0
            return u_apply, 0
        self.n_fail += 1
        u_apply, exit_code = self._get_old_solution(mu_0)
        if verbose:
            return u_apply, exit_code, None, None, None
        return u_apply, exit_code

    def _get_old_solution(self, x0_mu):
        """cautious_mpc.py:290-318: 1 = shifted old solution, 3 = feedback law only."""
        if self.n_fail < self.T and self.k_ff_old is not None:
            u_apply = self.k_ff_old[self.n_fail, :]
            exit_code = 1
        else:
            u_apply = np.dot(self.k_fb, np.reshape(x0_mu, (self.n_s,)))
            exit_code = 3
        return np.reshape(u_apply, (self.n_u, )), exit_code

    def update_model(self, x, y, opt_hyp=False, replace_old=True, reinitialize_solver=True):
        """cautious_mpc.py:444-479: the GP learns the residual to the linear prior on transformed inputs."""
        x = np.asarray(x, dtype=np.float64)
        n_train = np.shape(x)[0]
        x_s = x[:, :self.n_s].reshape((n_train, self.n_s))
        x_u = x[:, self.n_s:].reshape((n_train, self.n_u))
        y_prior = self.eval_prior(x_s, x_u)
        x_trafo = x_s if self.lin_trafo_gp_input is None else np.dot(x_s, np.asarray(self.lin_trafo_gp_input).T)
        self.gp.update_model(np.hstack((x_trafo, x_u)), np.asarray(y, dtype=np.float64) - y_prior, opt_hyp, replace_old)
        if reinitialize_solver:
            self.init_solver(self.cost_func)
        else:
            warnings.warn("Updating gp without reinitializing the solver!")
