"""Wall time of SamplingSafeMPC.get_action (the caller that replaces the IPOPT solve, safempc_simple.py:874-1001):
pendulum GP, n_safe = 10, 4096 candidates x 2 refinement iterations; cProfile of the hot loop.

    python scripts/profile_mpc.py [n_train] [n_samples]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import safe_exploration_b200 as se  # noqa: E402
from safe_exploration_b200 import workloads  # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n_safe = 10
w = workloads.make("C2", batch=8, n_train=n_train, horizon=n_safe)
gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
opt_env = {"l_mu": w.l_mu, "l_sigma": w.l_sigma, "h_mat_safe": w.h_mat, "h_safe": 0.5 * np.ones((2 * w.n_s, 1)),
           "lin_model": (w.a, w.b), "ctrl_bounds": np.array([[-1.0, 1.0]])}
mpc = se.SamplingSafeMPC(n_safe, gp, opt_env, np.eye(w.n_s), np.eye(w.n_u), beta_safety=2.0, n_samples=n_samples, n_iter=2,
                         n_elite=64, seed=1, opt_perf_trajectory={"n_perf": 1})
x0 = np.array([0.02, -0.03])
for _ in range(3):
    mpc.get_action(x0)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    u, ok = mpc.get_action(x0)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print("get_action: %.3f ms per call (N=%d, %d candidates x 2 iterations, H=%d) -> %.0f rollouts/s through the solver surface"
      % (1e3 * dt, n_train, n_samples, n_safe, 2 * n_samples / dt))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    mpc.get_action(x0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
