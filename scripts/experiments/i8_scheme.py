"""Numerical experiment (CPU, NumPy): the digit-plane scheme of round 2, emulated exactly.

  * W = L^-1 rows are scaled by the max-abs of their OFF-diagonal entries (the diagonal 1/L_ii dominates every row and
    would waste the leading digit plane); the diagonal term W_ii k*_i is added in float64 in the epilogue from the
    K* digits;
  * S digits per operand, pairs a + c < S  ->  S (S + 1) / 2 int8 products (S = 4: 10, S = 5: 15);
  * the statistical error model segp_factorize uses for its a-priori figure is printed next to the measured error.

    python scripts/experiments/i8_scheme.py [C4] [n_train] [noise] [dims]"""
import sys
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, ".")
from safe_exploration_b200 import workloads
from oracle import gp_oracle

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
n_train = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
noise = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-" else None
ndims = int(sys.argv[4]) if len(sys.argv) > 4 else 2
nb = 256
w = workloads.make(name, batch=nb, n_train=n_train)
rng = np.random.default_rng(7)
z = np.concatenate([w.p0[None] + 0.1 * rng.standard_normal((nb, w.n_s)), w.k_ff[:, 0]], axis=1)
zu = rng.uniform(-1, 1, size=z.shape)     # probe-style inputs: uniform over the training box


BASE = 256.0


def digits(r, s):
    """first s balanced base-256 digits of V = rn(r 127 2^32) (tri_i8.cu: split_digits)"""
    v = np.rint(r * (127.0 * 2.0 ** 32))
    out = []
    for a in range(4, 0, -1):
        d = np.mod(v + 128.0, 256.0) - 128.0
        out.append(d)
        v = (v - d) / 256.0
    out.append(v)
    return out[::-1][:s]


for d in range(min(w.n_s, ndims)):
    hyp = dict(w.hyp[d])
    if noise is not None:
        hyp["noise"] = noise
    var_f = hyp["variance"]
    kxx = gp_oracle.kernel(w.kern_types[d], w.x_train, w.x_train, var_f, hyp["lengthscale"])
    kxx[np.diag_indices_from(kxx)] += hyp["noise"] + 1e-5 + 1e-8
    L = np.linalg.cholesky(kxx)
    n = L.shape[0]
    W = sla.solve_triangular(L, np.eye(n), lower=True)
    dg = np.diag(W).copy()
    woff = W - np.diag(dg)
    rmax = np.abs(woff).max(axis=1)
    rmax = np.maximum(rmax, dg / 256.0)
    wn = woff / rmax[:, None]
    rows = np.arange(1, n + 1)
    rw2 = (wn ** 2).sum(axis=1) / rows          # mean square of the scaled row entries
    print("dim %d: noise %.1e  rowmax_off: median %.3g max %.3g; diag/rowmax median %.3g max %.3g" % (
        d, hyp["noise"], np.median(rmax), rmax.max(), np.median(dg / rmax), (dg / rmax).max()))
    for tag, zz in (("rollout-like", z), ("uniform", zu)):
        ks = gp_oracle.kernel(w.kern_types[d], w.x_train, zz, var_f, hyp["lengthscale"])
        v = W @ ks
        var = var_f - np.sum(v * v, axis=0)
        ku = ks / var_f
        kd = digits(ku, 5)
        rk2 = (ku ** 2).mean()
        print("  [%s] var/k** min %.2e median %.2e" % (tag, (var / var_f).min(), np.median(var / var_f)))
        wd = digits(wn, 5)
        for s in (4, 5):
            acc = np.zeros_like(v)
            for a in range(s):
                for c in range(s - a):
                    acc += (wd[a] @ kd[c]) / (127.0 * 127.0 * BASE ** (a + c))
            khat = sum(kd[c] / (127.0 * BASE ** c) for c in range(5)) * var_f
            vv = acc * rmax[:, None] * var_f + dg[:, None] * khat
            q = np.sum(vv * vv, axis=0)
            err = np.abs(q - np.sum(v * v, axis=0))
            # statistical model: per product term, variance u^2 [ (rw2 + rk2) / 12 + (s - 1) / 36 ], u = 1/(127 256^(s-1))
            u = 1.0 / (127.0 * BASE ** (s - 1))
            var_i = (rmax * var_f * u) ** 2 * rows * ((rw2 + rk2) / 12.0 + (s - 1) * (BASE / 127.0) ** 2 / 144.0)
            # d(sum v^2) = 2 sum v_i dv_i ; with sum v_i^2 <= var_f spread evenly:  std ~ 2 sqrt(var_f mean(var_i))
            model_rms = 2.0 * np.sqrt(var_f * var_i.mean())
            model_act = 2.0 * np.sqrt((v * v * var_i[:, None]).sum(axis=0))      # knowing v (oracle only)
            print("     S=%d: abs err max %.2e rms %.2e | model 1-sigma %.2e (v-weighted: median %.2e) | rel-to-var "
                  "max %.2e median %.2e" % (s, err.max(), np.sqrt((err ** 2).mean()), model_rms,
                                            np.median(model_act), (err / np.abs(var)).max(),
                                            np.median(err / np.abs(var))))
