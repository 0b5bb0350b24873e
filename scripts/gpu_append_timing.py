"""Time segp_append (incremental) against a from-scratch factorisation at the C4 model size (1 GPU)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import safe_exploration_b200 as se          # noqa: E402
from safe_exploration_b200 import workloads  # noqa: E402

for name, n0 in (("C4", 4900), ("C3", 1930)):
    w = workloads.make(name, batch=8)
    x, y = w.x_train, w.y_train
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, x[:n0], y[:n0], kern_types=w.kern_types, hyp=w.hyp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gp.train(x[:n0], y[:n0])
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    gp.append_data(x[n0:n0 + 1], y[n0:n0 + 1])          # first call: switches dense-W keeping on (full path)
    n = n0 + 1
    out = []
    for add in (1, 1, 1, 8, 32):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gp.append_data(x[n:n + add], y[n:n + add])
        torch.cuda.synchronize()
        out.append((add, time.perf_counter() - t0, bool(gp.get_option("append_incremental"))))
        n += add
    print("%s: N=%d n_s=%d: full factorisation %.1f ms; appends (points, ms, incremental): %s" % (
        name, n0, w.n_s, 1e3 * t_full, ", ".join("(%d, %.1f, %s)" % (a, 1e3 * t, i) for a, t, i in out)))
    gp.close()
