import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import safe_exploration_b200 as se
lib = se._lib.load()
def gemm(a, b, c, alpha, beta, trans_b, flags):
    m, k = a.shape; n = b.shape[0] if trans_b else b.shape[1]
    out = np.ascontiguousarray(c, dtype=np.float64).copy()
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    se._lib.check(lib.segp_i8_gemm_selftest(0, m, n, k, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                            out.ctypes.data_as(ctypes.c_void_p), float(alpha), float(beta), int(trans_b), int(flags)))
    return out
np.set_printoptions(linewidth=200, precision=6)
for (m, n, k) in ((128, 128, 64), (128, 128, 128), (256, 128, 64)):
    a = np.ones((m, k)); b = np.ones((n, k))
    g = gemm(a, b, np.zeros((m, n)), 1.0, 0.0, 1, 0)
    print(m, n, k, "ones: unique", np.unique(g)[:10], "expected", k)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((m, k)); b = rng.standard_normal((n, k))
    g = gemm(a, b, np.zeros((m, n)), 1.0, 0.0, 1, 0)
    w = a @ b.T
    err = np.abs(g - w)
    print("  random: max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape), "ratio g/w sample", (g / w)[0, :4], (g / w)[70, 60:64])
    print("  err by 32-col chunk and 32-row quadrant:"); 
    for q in range(m // 32): print("   ", [float("%.2e" % err[q*32:(q+1)*32, c*32:(c+1)*32].max()) for c in range(n // 32)])
