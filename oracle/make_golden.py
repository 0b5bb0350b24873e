"""Regenerate tests/golden/*.npz from the REFERENCE's own functions.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Run in the build container only:

    python -m oracle.make_golden

Every output array in the golden files is produced by the reference's unmodified
``gp_reachability.onestep_reachability`` / ``multistep_reachability`` /
``lin_ellipsoid_safety_distance`` / ``utils.compute_remainder_overapproximations`` /
``utils_ellipsoid.*`` (imported via oracle/ref_loader.py).  The ``ssm`` callable handed to them
is oracle.gp_oracle.GPOracle (GPy and CasADi cannot be installed here), evaluating the GP in the
explicit-inverse form the reference's SimpleGPModel.__call__ uses.  Inputs follow the reference's
own test fixtures (test/test_gp_reachability_casadi.py:30-67, test/test_utils_casadi.py:99-121,
test/test_safempc.py data) with FIXED hyper-parameters instead of optimised ones.
The input arrays are stored next to the outputs so the files are self-contained on the GPU box.
"""
import os

import numpy as np

from . import ref_loader
from .gp_oracle import GPOracle

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_TEST = os.path.join(ref_loader.REFERENCE_ROOT, "safe_exploration", "test")


def _run_multistep(reach, gp, p0, k_fb, k_ff, l_mu, l_sig, q0, c, a, b, k_fb_init):
    """One trajectory at a time through the reference's multistep_reachability."""
    bsz, hor, n_u = k_ff.shape
    n_s = p0.shape[-1]
    p_all = np.empty((bsz, hor, n_s))
    q_all = np.empty((bsz, hor, n_s, n_s))
    for i in range(bsz):
        kfb_i = k_fb if k_fb.ndim == 3 else k_fb[i]
        p0_i = (p0 if p0.ndim == 1 else p0[i]).reshape(n_s, 1)
        _, _, pa, qa = reach.multistep_reachability(p0_i, gp, kfb_i, k_ff[i], l_mu, l_sig, q0, c, 0,
                                                    a, b, k_fb_init)
        p_all[i] = np.real(pa)
        q_all[i] = np.real(qa)
    return p_all, q_all


def golden_invpend_c1(reach):
    """BASELINE config C1: inverted pendulum fixture, N=50, H=5, GPy default hyper-parameters
    (what gp.train(..., opt_hyp=False) yields: l=1, s2=1, noise=1 (+1e-5 +1e-8))."""
    d = np.load(os.path.join(REF_TEST, "invpend_data.npz"))
    x, y = d["X"][:50], d["y"][:50]
    n_s, n_u, hor, bsz = 2, 1, 5, 4
    gp_args = dict(kern_types=["rbf", "rbf"], lengthscale=np.ones((n_s, 3)), variance=np.ones(n_s),
                   noise=np.full(n_s, 1.0 + 1e-5 + 1e-8))
    gp = GPOracle(x, y, **gp_args)
    rng = np.random.RandomState(125)
    p0 = 0.1 * rng.randn(n_s)
    k_ff = 0.1 * rng.randn(bsz, hor, n_u)
    k_fb = np.tile(np.array([[-1.2, -0.4]]), (hor - 1, 1, 1))
    l_mu = np.array([0.05, 0.02])
    l_sig = np.array([0.05, 0.02])
    p_all, q_all = _run_multistep(reach, gp, p0, k_fb, k_ff, l_mu, l_sig, None, 2.0, None, None, None)
    np.savez(os.path.join(OUT, "invpend_c1.npz"), x_train=x, y_train=y, p0=p0, k_ff=k_ff, k_fb=k_fb,
             l_mu=l_mu, l_sigma=l_sig, c_safety=2.0, p_all=p_all, q_all=q_all,
             kern_types=np.array(gp_args["kern_types"]), lengthscale=gp_args["lengthscale"],
             variance=gp_args["variance"], noise=gp_args["noise"])


def golden_invpend_reach_test(reach):
    """The reference's test_gp_reachability_casadi.py:30-67 recipe (seed 125, random a/b, m=50
    random subset, L=1e-3, c_safety=2, q = .2*[[.5,.2],[.2,.65]] or None), fixed hyper-parameters."""
    np.random.seed(125)
    n_s, n_u, c = 2, 1, 2.0
    a = np.random.rand(n_s, n_s)
    b = np.random.rand(n_s, n_u)
    d = np.load(os.path.join(REF_TEST, "invpend_data.npz"))
    idx = np.random.choice(d["X"].shape[0], size=50, replace=False)
    x, y = d["X"][idx], d["y"][idx]
    gp_args = dict(kern_types=["rbf", "mat52"],
                   lengthscale=np.array([[0.9, 1.7, 2.3], [1.4, 0.8, 3.0]]),
                   variance=np.array([0.6, 1.3]), noise=np.array([0.02, 0.05]) + 1e-5 + 1e-8)
    gp = GPOracle(x, y, **gp_args)
    l_mu = np.array([0.001] * n_s)
    l_sig = np.array([0.001] * n_s)
    k_fb = np.random.rand(n_u, n_s)
    k_ff = np.random.rand(n_u, 1)
    p = .1 * np.random.randn(n_s, 1)
    q = .2 * np.array([[.5, .2], [.2, .65]])
    out = {}
    for tag, aa, bb in (("lin", a, b), ("nolin", None, None)):
        p1, q1 = reach.onestep_reachability(p, gp, k_ff, l_mu, l_sig, q, k_fb, c, 0, a=aa, b=bb)
        out["p1_set_" + tag], out["q1_set_" + tag] = np.real(p1), np.real(q1)
        p1, q1 = reach.onestep_reachability(p, gp, k_ff, l_mu, l_sig, None, k_fb, c, 0, a=aa, b=bb)
        out["p1_point_" + tag], out["q1_point_" + tag] = np.real(p1), np.real(q1)
    # multistep, T=3 (test_gp_reachability_casadi.py:97-150)
    t = 3
    u_0 = .2 * np.random.randn(n_u, 1)
    k_fb_0 = np.random.randn(t - 1, n_u, n_s)
    k_ff_m = np.random.randn(t - 1, n_u)
    k_ff_all = np.vstack((u_0.T, k_ff_m))
    k_fb_apply = k_fb_0 + k_fb[None]
    _, _, p_all, q_all = reach.multistep_reachability(p, gp, k_fb_apply, k_ff_all, l_mu, l_sig, None, c, 0,
                                                      a, b, None)
    # safety distance on the last ellipsoid (gp_reachability.py:215-250)
    h_mat = np.array([[1., 0.], [-1., 0.], [0., 1.], [0., -1.], [0.6, 0.8]])
    h_vec = np.array([[1.], [1.], [2.], [2.], [1.5]])
    dist = reach.lin_ellipsoid_safety_distance(np.real(p_all[-1])[:, None], np.real(q_all[-1]), h_mat, h_vec, c)
    np.savez(os.path.join(OUT, "invpend_reach_test.npz"), x_train=x, y_train=y, a=a, b=b, p=p, q=q,
             k_fb=k_fb, k_ff=k_ff, l_mu=l_mu, l_sigma=l_sig, c_safety=c,
             k_fb_multi=k_fb_apply, k_ff_multi=k_ff_all, p_all=np.real(p_all), q_all=np.real(q_all),
             h_mat=h_mat, h_vec=h_vec, dist=np.real(dist),
             kern_types=np.array(gp_args["kern_types"]), lengthscale=gp_args["lengthscale"],
             variance=gp_args["variance"], noise=gp_args["noise"], **out)


def golden_cartpole(reach):
    """Cart-pole fixture (test/data_cartpole.npz: X 173x5, y 173x4, a, b), mixed rbf / mat52 kernels,
    per-trajectory feedback gains, H=6, B=6, both with and without an initial ellipsoid."""
    d = np.load(os.path.join(REF_TEST, "data_cartpole.npz"), allow_pickle=True)
    x, y, a, b = d["X"], d["y"], d["a"], d["b"]
    n_s, n_u, hor, bsz = 4, 1, 6, 6
    rng = np.random.RandomState(12345)
    gp_args = dict(kern_types=["rbf", "mat52", "mat52", "rbf"],
                   lengthscale=rng.uniform(0.8, 2.5, size=(n_s, 5)),
                   variance=rng.uniform(0.5, 1.5, size=n_s),
                   noise=rng.uniform(0.005, 0.02, size=n_s) + 1e-5 + 1e-8)
    gp = GPOracle(x, y, **gp_args)
    p0 = 0.05 * rng.randn(bsz, n_s)
    k_ff = 0.1 * rng.randn(bsz, hor, n_u)
    k_fb = 0.3 * rng.randn(bsz, hor - 1, n_u, n_s)
    k_fb_init = 0.3 * rng.randn(n_u, n_s)
    l_mu = np.array([1e-3, 2e-3, 1e-3, 3e-3])
    l_sig = np.array([2e-3, 1e-3, 2e-3, 1e-3])
    q0 = 0.01 * np.array([[2., .3, 0., .1], [.3, 1., .2, 0.], [0., .2, 1.5, .4], [.1, 0., .4, 1.]])
    p_all, q_all = _run_multistep(reach, gp, p0, k_fb, k_ff, l_mu, l_sig, None, 2.0, a, b, None)
    p_all_q0, q_all_q0 = _run_multistep(reach, gp, p0, k_fb, k_ff, l_mu, l_sig, q0, 1.5, a, b, k_fb_init)
    np.savez(os.path.join(OUT, "cartpole.npz"), x_train=x, y_train=y, a=a, b=b, p0=p0, k_ff=k_ff, k_fb=k_fb,
             k_fb_init=k_fb_init, q0=q0, l_mu=l_mu, l_sigma=l_sig, p_all=p_all, q_all=q_all,
             p_all_q0=p_all_q0, q_all_q0=q_all_q0,
             kern_types=np.array(gp_args["kern_types"]), lengthscale=gp_args["lengthscale"],
             variance=gp_args["variance"], noise=gp_args["noise"])


def golden_ellipsoid_algebra(utils, uell):
    """test_utils_casadi.py:99-121 inputs (seed 0; (n_s,n_u) in (2,1),(3,2),(5,4),(8,3)) through the
    reference's compute_remainder_overapproximations; plus random sum_two_ellipsoids /
    ellipsoid_from_rectangle cases and the literal cases of test_utils_ellipsoid.py:13-27."""
    out = {}
    for tag, (n_s, n_u) in zip(("t_1", "t_2", "t_3", "t_4"), ((2, 1), (3, 2), (5, 4), (8, 3))):
        np.random.seed(0)
        x_0 = np.random.rand(n_s, n_s)
        q = x_0 @ x_0.T + 0.1 * np.eye(n_s)
        k_fb = np.random.randn(n_u, n_s)
        l_mu = np.array([.1] * n_s)
        l_sig = np.array([.1] * n_s)
        u_mu, u_sig = utils.compute_remainder_overapproximations(q, k_fb, l_mu, l_sig)
        out.update({tag + "_q": q, tag + "_k_fb": k_fb, tag + "_l_mu": l_mu, tag + "_l_sigma": l_sig,
                    tag + "_u_mu": np.real(u_mu), tag + "_u_sigma": np.real(u_sig)})
    rng = np.random.RandomState(7)
    for i, n in enumerate((2, 4, 10)):
        m1 = rng.randn(n, n)
        m2 = rng.randn(n, n)
        q1 = m1 @ m1.T + 0.05 * np.eye(n)
        q2 = m2 @ m2.T + 0.05 * np.eye(n)
        p1 = rng.randn(n, 1)
        p2 = rng.randn(n, 1)
        ps, qs = uell.sum_two_ellipsoids(p1, q1, p2, q2)
        ub = rng.uniform(0.01, 2.0, size=n)
        out.update({"s%d_p1" % i: p1, "s%d_q1" % i: q1, "s%d_p2" % i: p2, "s%d_q2" % i: q2,
                    "s%d_p" % i: ps, "s%d_q" % i: qs, "s%d_ub" % i: ub,
                    "s%d_qrect" % i: uell.ellipsoid_from_rectangle(ub)})
    for tag, ub in (("rectangle", [0.1, 0.3, 0.5]), ("cube", [0.1] * 3)):
        out["rect_" + tag + "_ub"] = np.array(ub)
        out["rect_" + tag + "_q"] = uell.ellipsoid_from_rectangle(ub)
    np.savez(os.path.join(OUT, "ellipsoid_algebra.npz"), **out)


def golden_uncertainty_propagation():
    """Outputs of the reference's own multi_step_taylor_symbolic / mean_equivalent_multistep
    (uncertainty_propagation_casadi.py, run numerically through the NumPy-backed CasADi shim) on the cart-pole
    fixture, with and without the linear prior and the GP-input transform."""
    up = ref_loader.load_uncertainty_propagation()
    d = np.load(os.path.join(REF_TEST, "data_cartpole.npz"))
    x, y = d["X"][:80], d["y"][:80]
    n_s, n_u, hor, bsz = y.shape[1], x.shape[1] - y.shape[1], 5, 6
    rng = np.random.RandomState(11)
    ls = rng.uniform(0.8, 2.5, size=(n_s, n_s + n_u))
    var = rng.uniform(0.5, 1.5, size=n_s)
    noise = np.full(n_s, 0.05 + 1e-5 + 1e-8)
    kern = ["rbf", "mat52", "rbf", "mat52"][:n_s]
    gp = GPOracle(x, y, kern, ls, var, noise)
    a = np.eye(n_s) + 0.05 * rng.randn(n_s, n_s)
    b = 0.1 * rng.randn(n_s, n_u)
    mu0 = 0.1 * rng.randn(bsz, n_s)
    k_ff = 0.2 * rng.randn(bsz, hor, n_u)
    k_fb = 0.3 * rng.randn(bsz, hor - 1, n_u, n_s)
    out = dict(x_train=x, y_train=y, kern_types=np.array(kern), lengthscale=ls, variance=var, noise=noise, a=a, b=b,
               mu0=mu0, k_ff=k_ff, k_fb=k_fb)
    # GP on a reduced input (drop the first state, as the reference's cart-pole configs do)
    t_mat = np.eye(n_s)[1:]
    gp_t = GPOracle(x[:, 1:], y, kern, ls[:, 1:], var, noise)
    out["t_mat"] = t_mat
    for tag, fn in (("taylor", up.multi_step_taylor_symbolic), ("meaneq", up.mean_equivalent_multistep)):
        for pr, (aa, bb, model, tm) in (("lin", (a, b, gp, None)), ("nolin", (None, None, gp, None)),
                                        ("trafo", (a, b, gp_t, t_mat))):
            mu_all = np.empty((bsz, hor, n_s))
            sig_all = np.empty((bsz, hor, n_s, n_s))
            for i in range(bsz):
                m, sg, _ = fn(mu0[i].reshape(n_s, 1), model, k_ff[i], k_fb[i], None, aa, bb, tm)
                mu_all[i] = np.asarray(m, dtype=np.float64)
                sig_all[i] = np.asarray(sg, dtype=np.float64).reshape(hor, n_s, n_s)
            out["mu_%s_%s" % (tag, pr)] = mu_all
            out["sigma_%s_%s" % (tag, pr)] = sig_all
    np.savez(os.path.join(OUT, "uncertainty_propagation.npz"), **out)


GP_CASES = (
    # name, fixture, N, kernels per output dimension
    ("pend_rbf_mat52", "invpend_data.npz", 60, ["rbf", "mat52"]),
    ("pend_composite", "invpend_data.npz", 70, ["lin_rbf", "lin_mat52"]),
    ("cart_mixed", "data_cartpole.npz", 130, ["rbf", "lin_mat52", "mat52", "lin_rbf"]),
)


def gp_case_hyp(kern, dim, rng):
    """Fixed hyper-parameters in the reference's own dict layout (ssm_gpy/gaussian_process.py:515-538)."""
    if kern in ("rbf", "mat52"):
        return {"lengthscale": rng.uniform(0.7, 2.5, size=dim), "variance": float(rng.uniform(0.5, 1.5))}
    st = "rbf" if kern == "lin_rbf" else "mat52"
    return {"prod.%s.lengthscale" % st: np.array([rng.uniform(0.6, 1.8)]),
            "prod.%s.variance" % st: float(rng.uniform(0.5, 1.5)),
            "prod.linear.variances": np.array([rng.uniform(0.3, 1.2)]),
            "linear.variances": rng.uniform(0.05, 0.6, size=dim)}


def golden_gp_pred():
    """Kernel rows, prior variances, predictive means and variances from the reference's OWN functions
    (ssm_gpy/gp_models_utils_casadi.py: _k_rbf, _k_mat52, _k_lin_rbf, _k_lin_mat52 through _get_kernel_function, and
    gp_pred :177-197 -- the arithmetic SimpleGPModel.__call__ executes), evaluated numerically through the NumPy-backed
    CasADi shim.  The posterior state handed to gp_pred is what SimpleGPModel.train stores
    (ssm_gpy/gaussian_process.py:258-263): inv_K = (K + noise I)^-1 and beta = inv_K y, with K from the same reference
    kernel function (diagonal from its diag_only branch; the off-diagonal branch takes sqrt of a rounded-negative
    r^2 there) and the inverse by LAPACK."""
    g = ref_loader.load_gp_utils()
    out = {}
    for name, fixture, n, kerns in GP_CASES:
        d = np.load(os.path.join(REF_TEST, fixture), allow_pickle=True)
        x, y = np.asarray(d["X"][:n], dtype=np.float64), np.asarray(d["y"][:n], dtype=np.float64)
        n_s, dim = y.shape[1], x.shape[1]
        rng = np.random.RandomState(len(name) * 7 + n)
        z = np.vstack((x[:4] + 0.05 * rng.randn(4, dim), rng.uniform(-1.0, 1.0, size=(12, dim)) * np.abs(x).max(axis=0)))
        noise = rng.uniform(0.01, 0.05, size=n_s) + 1e-5 + 1e-8
        out[name + "/x_train"], out[name + "/y_train"], out[name + "/z"] = x, y, z
        out[name + "/noise"], out[name + "/kern_types"] = noise, np.array(kerns)
        mu = np.empty((z.shape[0], n_s))
        var = np.empty((z.shape[0], n_s))
        prior = np.empty((z.shape[0], n_s))
        kst = np.empty((n_s, z.shape[0], n))
        for i, kern in enumerate(kerns):
            hyp = gp_case_hyp(kern, dim, rng)
            for k, v in hyp.items():
                out["%s/hyp%d/%s" % (name, i, k)] = np.asarray(v)
            kfun = g._get_kernel_function(kern, hyp)
            with np.errstate(invalid="ignore"):
                kmat = np.asarray(kfun(x, y=x), dtype=np.float64)
            kmat = 0.5 * (kmat + kmat.T)
            kmat[np.diag_indices(n)] = np.asarray(kfun(x, diag_only=True), dtype=np.float64).reshape(-1) + noise[i]
            inv_k = np.linalg.inv(kmat)
            inv_k = 0.5 * (inv_k + inv_k.T)
            beta = inv_k @ y[:, i:i + 1]
            m, v = g.gp_pred(z, kfun, beta, x, inv_k)
            mu[:, i], var[:, i] = np.asarray(m).reshape(-1), np.asarray(v).reshape(-1)
            prior[:, i] = np.asarray(kfun(z, diag_only=True), dtype=np.float64).reshape(-1)
            kst[i] = np.asarray(kfun(z, y=x), dtype=np.float64)
        out[name + "/mu"], out[name + "/var"], out[name + "/prior"], out[name + "/kstar"] = mu, var, prior, kst
    np.savez(os.path.join(OUT, "gp_pred_reference.npz"), **out)


def main():
    reach, utils, uell = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    golden_gp_pred()
    golden_uncertainty_propagation()
    golden_invpend_c1(reach)
    golden_invpend_reach_test(reach)
    golden_cartpole(reach)
    golden_ellipsoid_algebra(utils, uell)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
