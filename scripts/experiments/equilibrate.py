"""Numerical experiment (CPU, NumPy): does two-sided scaling of W_off = W - diag(W) (rows R_i, columns C_j >= 1 folded
into K* as k_j / C_j) shrink its dynamic range enough to drop a digit plane?  See DESIGN.md section 4.
    python scripts/experiments/equilibrate.py [C4] [n_train]"""
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


for d in range(min(w.n_s, 2)):
    hyp = w.hyp[d]
    kxx = gp_oracle.kernel(w.kern_types[d], w.x_train, w.x_train, hyp["variance"], hyp["lengthscale"])
    kxx[np.diag_indices_from(kxx)] += hyp["noise"] + 1e-5 + 1e-8
    L = np.linalg.cholesky(kxx)
    W = sla.solve_triangular(L, np.eye(L.shape[0]), lower=True)
    ks = gp_oracle.kernel(w.kern_types[d], w.x_train, z, hyp["variance"], hyp["lengthscale"]) / hyp["variance"]
    v = W @ ks * hyp["variance"]
    var = hyp["variance"] - np.sum(v * v, axis=0)
    dg = np.diag(W).copy()
    woff = np.abs(W - np.diag(dg))
    for cap in (1.0, 16.0, 254.0):
        c = np.ones(W.shape[0])
        for _ in range(6):
            r = (woff * c[None, :]).max(axis=1)
            r[r == 0] = 1.0
            cm = (woff / r[:, None]).max(axis=0)
            cm[cm == 0] = 1.0
            c = np.clip(1.0 / cm, 1.0, cap)
        r = (woff * c[None, :]).max(axis=1)
        r[r == 0] = 1.0
        what = (W - np.diag(dg)) * c[None, :] / r[:, None]
        khat = ks / c[:, None]
        print("dim %d cap %5.0f: median C %.1f; mean |What| %.4f (was %.4f)" % (
            d, cap, np.median(c), np.abs(what).mean(), (woff / woff.max(axis=1, keepdims=True).clip(1e-300)).mean()))
        for sw, sk, g in ((5, 5, 5), (4, 5, 4), (4, 4, 4), (4, 5, 5)):
            wd = digits(what, sw)
            kd = digits(khat, sk)
            kq = sum(kd[q] / (127.0 * 254.0 ** q) for q in range(sk))
            acc = np.zeros_like(v)
            npr = 0
            for a in range(sw):
                for q in range(sk):
                    if a + q < g:
                        acc += (wd[a] @ kd[q]) / (127.0 * 127.0 * 254.0 ** (a + q))
                        npr += 1
            vv = (acc * r[:, None] + dg[:, None] * (kq * c[:, None])) * hyp["variance"]
            rel = np.abs((hyp["variance"] - np.sum(vv * vv, axis=0)) - var) / np.abs(var)
            print("     digits W %d K %d, a+c<%d (%2d products): var rel err max %.2e median %.2e" % (
                sw, sk, g, npr, rel.max(), np.median(rel)))
