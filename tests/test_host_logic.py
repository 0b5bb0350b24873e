"""CPU tests of host-side logic that needs no GPU: the product's hyper-parameter mapping of the composite kernels against
the oracle's (which is pinned by the reference's own kernel functions), bench.py's bookkeeping helpers."""
import numpy as np
import pytest

from oracle import gp_oracle


@pytest.mark.parametrize("kern", ["lin_rbf", "lin_mat52"])
@pytest.mark.parametrize("semantics", ["casadi", "gpy"])
def test_composite_hyper_parameter_mapping_matches_the_oracle(kern, semantics):
    from safe_exploration_b200.ssm import _composite_vectors
    st = kern[4:]
    rng = np.random.default_rng(0)
    dim = 4
    hyp = {"prod.%s.lengthscale" % st: np.array([rng.uniform(0.5, 2)]), "prod.%s.variance" % st: rng.uniform(0.5, 2),
           "prod.linear.variances": np.array([rng.uniform(0.2, 1)]), "linear.variances": rng.uniform(0.1, 1, dim)}
    got = _composite_vectors(kern, hyp, dim, semantics)
    want = gp_oracle.composite_vectors(kern, hyp, dim, semantics)
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # a scalar linear variance is broadcast (GPy's non-ARD Linear), a wrong length is rejected
    hyp1 = dict(hyp, **{"linear.variances": np.array([0.3])})
    assert np.array_equal(_composite_vectors(kern, hyp1, dim, semantics)[2], np.full(dim, 0.3))
    with pytest.raises(ValueError):
        _composite_vectors(kern, dict(hyp, **{"linear.variances": np.ones(dim + 1)}), dim, semantics)
    with pytest.raises(ValueError):
        _composite_vectors(kern, hyp, dim, "other")


def test_casadi_semantics_uses_input_column_one_only():
    """gp_models_utils_casadi.py:82-95: the product term of the composite kernels sees x[:, 1]."""
    from safe_exploration_b200.ssm import _composite_vectors
    hyp = {"prod.rbf.lengthscale": np.array([0.7]), "prod.rbf.variance": 1.3, "prod.linear.variances": np.array([0.4]),
           "linear.variances": np.array([0.1, 0.2, 0.3])}
    s, a, v, var = _composite_vectors("lin_rbf", hyp, 3, "casadi")
    assert np.array_equal(s, [0.0, 1.0 / 0.7, 0.0]) and np.array_equal(a, [0.0, 0.4, 0.0])
    assert np.array_equal(v, [0.1, 0.2, 0.3]) and var == 1.3
    s, a, _, _ = _composite_vectors("lin_rbf", hyp, 3, "gpy")
    assert np.allclose(s, 1.0 / 0.7) and np.allclose(a, 0.4)


def test_bench_l2_policy_and_workload_sizes():
    import bench
    from safe_exploration_b200 import workloads
    assert "no flush" in bench._l2_policy(workloads.make("C2", batch=8), 4096)
    for name, b in (("C3", 16384), ("C4", 8192)):
        assert bench._l2_policy(workloads.make(name, batch=8, n_train=None), b).startswith("inputs larger than L2")
    # algorithmic work of SURVEY section 8d: F_step = n_s [N^2 + N (5 D + 9)]
    assert workloads.flop_per_step(4, 1, 5000) == 4 * (5000 ** 2 + 5000 * (5 * 5 + 9))
