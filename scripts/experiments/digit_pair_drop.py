"""Numerical experiment (CPU, NumPy): which of the 15 digit-plane products (a + c < 5; a = W digit, c = K* digit) of the
tcgen05 int8 contraction could be dropped?  Emulates the product path exactly (balanced base-254 digits, exact integer
products) and reports the variance error against float64 on the benchmark model for several drop sets.
    python scripts/experiments/digit_pair_drop.py [C4] [n_train]"""
import sys
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, ".")
from safe_exploration_b200 import workloads
from oracle import gp_oracle

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
n_train = int(sys.argv[2]) if len(sys.argv) > 2 else None
nb = 192
w = workloads.make(name, batch=nb, n_train=n_train)
rng = np.random.default_rng(7)
z = np.concatenate([w.p0[None] + 0.1 * rng.standard_normal((nb, w.n_s)), w.k_ff[:, 0]], axis=1)
S = 5


def digits(r):
    """r in [-1, 1] -> S balanced base-254 digits (first scale 127), as float arrays."""
    out = []
    x = r * 127.0
    for a in range(S):
        q = np.rint(x)
        out.append(q)
        x = (x - q) * 254.0
    return out


drop_sets = {"none (15 products)": [], "drop (0,4)": [(0, 4)], "drop (4,0)": [(4, 0)], "drop (1,3)": [(1, 3)],
             "drop (0,4),(0,3)": [(0, 4), (0, 3)], "drop all a+c=4 (10 products)": [(a, 4 - a) for a in range(5)]}
for d in range(w.n_s):
    hyp = w.hyp[d]
    kxx = gp_oracle.kernel(w.kern_types[d], w.x_train, w.x_train, hyp["variance"], hyp["lengthscale"])
    kxx[np.diag_indices_from(kxx)] += hyp["noise"] + 1e-5 + 1e-8
    L = np.linalg.cholesky(kxx)
    W = sla.solve_triangular(L, np.eye(L.shape[0]), lower=True)
    ks = gp_oracle.kernel(w.kern_types[d], w.x_train, z, hyp["variance"], hyp["lengthscale"])
    v = W @ ks
    var = hyp["variance"] - np.sum(v * v, axis=0)
    rowmax = np.abs(W).max(axis=1)
    wd = digits(W / rowmax[:, None])
    kd = digits(ks / hyp["variance"])
    print("dim %d: var/k** median %.2e; mean |W digit 0| %.2f (of 127), mean |K* digit 0| %.1f" % (
        d, np.median(var / hyp["variance"]), np.abs(wd[0]).mean(), np.abs(kd[0]).mean()))
    prods = {(a, c): wd[a] @ kd[c] for a in range(S) for c in range(S - a)}
    for label, drop in drop_sets.items():
        acc = np.zeros_like(v)
        for (a, c), p in prods.items():
            if (a, c) not in drop:
                acc += p / (127.0 * 127.0 * 254.0 ** (a + c))
        vv = acc * rowmax[:, None] * hyp["variance"]
        rel = np.abs((hyp["variance"] - np.sum(vv * vv, axis=0)) - var) / np.abs(var)
        print("   %-32s var rel err max %.2e median %.2e" % (label, rel.max(), np.median(rel)))
