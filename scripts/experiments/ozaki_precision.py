"""Numerical experiment (CPU, NumPy): how many int8 slices does the W x K* contraction need so that
sigma^2 = k** - |W k*|^2 keeps rtol 1e-4 on the benchmark models?  Emulates the slicing exactly
(integer slices, exact integer accumulation) and compares with float64."""
import sys
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, ".")
from safe_exploration_b200 import workloads
from oracle import gp_oracle

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
n_train = int(sys.argv[2]) if len(sys.argv) > 2 else None
nb = 256
w = workloads.make(name, batch=nb, n_train=n_train)
rng = np.random.default_rng(7)
# test inputs: around the rollout region (p0 + small spread, LQR-like controls)
z = np.concatenate([w.p0[None] + 0.1 * rng.standard_normal((nb, w.n_s)), w.k_ff[:, 0]], axis=1)


def slice_rows(a, s, bits):
    """a: (R, K) float64.  Per-row power-of-two scale, then s signed slices of `bits` magnitude bits.
    Returns int slices (s, R, K) and scale exponents e (R,) with a ~= 2^e * sum_i slices[i] * 2^(-bits (i+1))."""
    amax = np.abs(a).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax, 1e-300))) + 1, 0)   # |a| < 2^e
    r = a / np.exp2(e)[:, None]                                                   # |r| < 1
    out = []
    for i in range(s):
        r = r * (1 << bits)
        q = np.rint(r) if True else np.trunc(r)     # round to nearest: |q| <= 2^bits, residual in [-.5,.5]
        out.append(q)
        r = r - q
    return np.stack(out), e


for d in range(min(w.n_s, 2)):
    hyp = w.hyp[d]
    ls = hyp["lengthscale"]
    kxx = gp_oracle.kernel(w.kern_types[d], w.x_train, w.x_train, hyp["variance"], ls)
    kxx[np.diag_indices_from(kxx)] += hyp["noise"] + 1e-5 + 1e-8
    L = np.linalg.cholesky(kxx)
    W = sla.solve_triangular(L, np.eye(L.shape[0]), lower=True)
    ks = gp_oracle.kernel(w.kern_types[d], w.x_train, z, hyp["variance"], ls)      # (N, nb)
    v = W @ ks
    var = hyp["variance"] - np.sum(v * v, axis=0)
    print("dim %d: var/k** min %.3e median %.3e ; |W| max %.2f" % (d, (var / hyp["variance"]).min(),
                                                                 np.median(var / hyp["variance"]), np.abs(W).max()))
    for bits in (6, 7):
        for s in (4, 5, 6, 7):
            ws, ew = slice_rows(W, s, bits)                 # rows of W
            bs, eb = slice_rows(ks.T.copy(), s, bits)       # columns of K*
            acc = np.zeros_like(v)
            for i in range(s):
                for j in range(s - i):
                    acc += (ws[i] @ bs[j].T) * 2.0 ** (-bits * (i + j + 2))
            vv = acc * np.exp2(ew)[:, None] * np.exp2(eb)[None, :]
            var_o = hyp["variance"] - np.sum(vv * vv, axis=0)
            rel = np.abs(var_o - var) / np.abs(var)
            print("   bits %d slices %d (products %2d): var rel err max %.2e median %.2e" % (
                bits, s, s * (s + 1) // 2, rel.max(), np.median(rel)))
    # fp32-accumulating pipelines for comparison (what any floating tcgen05 kind delivers at best): float32
    # operands, products and sums
    v32 = (W.astype(np.float32) @ ks.astype(np.float32)).astype(np.float64)
    var32 = hyp["variance"] - np.sum(v32 * v32, axis=0)
    rel = np.abs(var32 - var) / np.abs(var)
    print("   float32 operands + float32 accumulation: var rel err max %.2e median %.2e" % (rel.max(), np.median(rel)))
