"""Small end-to-end case for compute-sanitizer: every kernel family once (tcgen05 contractions in modes 4 and 5, the
float64 path with composite kernels, incremental append, data selection, scoring, the factorisation with its GEMMs on
tcgen05 and the GEMM self-test)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import safe_exploration_b200 as se          # noqa: E402
from safe_exploration_b200 import workloads  # noqa: E402

w = workloads.make("C3", batch=200, n_train=300, horizon=3)
args = (w.l_mu, w.l_sigma, None, None, w.c_safety, w.a, w.b)
out = {}
for mode in (4, 5, 0):
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp, tri_mode=mode)
    out[mode] = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
    if mode == 4:
        gp.update_model(w.x_train[:1] + 0.1, w.y_train[:1], replace_old=False)     # switches dense-W keeping on
        gp.update_model(w.x_train[:3] + 0.2, w.y_train[:3], replace_old=False)     # incremental
        assert gp.get_option("append_incremental") == 1
        idx, _ = gp.select_maxvar(w.x_train, 40)
        sc = se.score_rollouts(out[4], w.k_ff, w.k_fb, w.h_mat, np.ones((2 * w.n_s, 1)), np.array([[-1.0, 1.0]]))
        assert sc.cost.shape == (200,)
    gp.close()
assert np.array_equal(out[4].q_all, out[5].q_all)
dim = w.n_s + w.n_u
hyp = [{"prod.mat52.lengthscale": np.array([0.9]), "prod.mat52.variance": 1.0, "prod.linear.variances": np.array([0.5]),
        "linear.variances": np.full(dim, 0.2), "noise": 1e-2}] + [dict(h) for h in w.hyp[1:]]
gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=["lin_mat52"] + list(w.kern_types[1:]), hyp=hyp)
res = se.rollout(gp, w.p0, w.k_ff, w.k_fb, *args)
assert np.all(np.isfinite(res.q_all))
gp.close()
# factorisation with the tensor-core GEMMs (fd_split, gemm_i8d, two-level potrf, trtri levels >= 256) + scratch cache
w2 = workloads.make("C3", batch=96, n_train=600, horizon=2)
gp = se.BatchedGPSSM(w2.n_s, w2.n_s, w2.n_u, None, None, kern_types=w2.kern_types, hyp=w2.hyp)
gp.set_option("fact_i8", 1)
gp.train(w2.x_train, w2.y_train)
assert gp.get_option("fact_i8_effective") == 1
b1 = gp.beta.copy()
gp.train(w2.x_train, w2.y_train)          # second factorisation: scratch from the cache
assert gp.get_option("scratch_cached_bytes") > 0 and np.array_equal(b1, gp.beta)
res2 = se.rollout(gp, w2.p0, w2.k_ff, w2.k_fb, w2.l_mu, w2.l_sigma, None, None, w2.c_safety, w2.a, w2.b)
assert np.all(np.isfinite(res2.q_all))
gp.close()
import ctypes  # noqa: E402
lib = se._lib.load()
rng = np.random.default_rng(0)
a_m, b_m, c_m = rng.standard_normal((256, 192)), rng.standard_normal((128, 192)), np.zeros((256, 128))
se._lib.check(lib.segp_i8_gemm_selftest(0, 256, 128, 192, a_m.ctypes.data_as(ctypes.c_void_p),
                                        b_m.ctypes.data_as(ctypes.c_void_p), c_m.ctypes.data_as(ctypes.c_void_p),
                                        1.0, 0.0, 1, 0))
assert np.allclose(c_m, a_m @ b_m.T, rtol=0, atol=1e-12)
print("sanitizer case ok")
