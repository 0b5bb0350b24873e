"""Numerical experiment (CPU, NumPy): split W = diag(W) + W_off, scale each row of W_off by ITS max-abs (the
diagonal entry 1/L_ii dominates every row of W = L^-1 on the benchmark models and wastes the leading digit plane),
add the diagonal term D_ii k*_i in float64 in the epilogue.  How many digit-plane products are then needed?
    python scripts/experiments/diag_split.py [C4] [n_train]"""
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


def digits(r, s):
    out = []
    x = r * 127.0
    for a in range(s):
        q = np.rint(x)
        out.append(q)
        x = (x - q) * 254.0
    return out


for d in range(w.n_s):
    hyp = w.hyp[d]
    kxx = gp_oracle.kernel(w.kern_types[d], w.x_train, w.x_train, hyp["variance"], hyp["lengthscale"])
    kxx[np.diag_indices_from(kxx)] += hyp["noise"] + 1e-5 + 1e-8
    L = np.linalg.cholesky(kxx)
    W = sla.solve_triangular(L, np.eye(L.shape[0]), lower=True)
    ks = gp_oracle.kernel(w.kern_types[d], w.x_train, z, hyp["variance"], hyp["lengthscale"])
    v = W @ ks
    var = hyp["variance"] - np.sum(v * v, axis=0)
    dg = np.diag(W).copy()
    woff = W - np.diag(dg)
    rmax = np.abs(woff).max(axis=1)
    rmax[rmax == 0] = 1.0
    print("dim %d: var/k** median %.2e; rowmax(off)/diag: median %.3e max %.3e" % (
        d, np.median(var / hyp["variance"]), np.median(rmax / dg), np.max(rmax / dg)))
    kd = digits(ks / hyp["variance"], 5)
    khat = sum(kd[c] / (127.0 * 254.0 ** c) for c in range(5)) * hyp["variance"]
    for sw, g in ((5, 5), (4, 4), (4, 5), (3, 3), (3, 4)):
        wd = digits(woff / rmax[:, None], sw)
        acc = np.zeros_like(v)
        npr = 0
        for a in range(sw):
            for c in range(5):
                if a + c < g:
                    acc += (wd[a] @ kd[c]) / (127.0 * 127.0 * 254.0 ** (a + c))
                    npr += 1
        vv = acc * rmax[:, None] * hyp["variance"] + dg[:, None] * khat
        rel = np.abs((hyp["variance"] - np.sum(vv * vv, axis=0)) - var) / np.abs(var)
        print("   W_off digits %d, pairs a+c<%d (%2d products): var rel err max %.2e median %.2e" % (
            sw, g, npr, rel.max(), np.median(rel)))
