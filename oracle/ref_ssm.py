"""The reference's CPU path, assembled from the reference's own code where that code can run here.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Used by ``bench.py --impl reference``.

``ReferenceGP`` is what ``SimpleGPModel.__call__`` evaluates (ssm_gpy/gaussian_process.py:135-175 ->
gp_models_utils_casadi.py:234-288): per output dimension the reference's OWN ``gp_pred`` (:177-197, explicit-inverse
quadratic form, 2 N^2 flop) on the reference's OWN kernel functions (``_get_kernel_function``, :200-231), run
numerically through the NumPy-backed CasADi stand-in.  Two things are restated because their providers are not
installable: the posterior state ``inv_K, beta`` (GPy: ``posterior.woodbury_inv/vector``, ssm_gpy/
gaussian_process.py:258-261; here LAPACK through NumPy) and the mean Jacobian (CasADi algorithmic differentiation,
:275-280; here the closed form of oracle/gp_oracle.py, pinned by finite differences).  The reference rebuilds an SX
expression graph on every call (:240-260); that construction cost has no numeric counterpart and is NOT charged: the
timed arm is the reference's arithmetic, not its graph building.
"""
import numpy as np

from . import ref_loader


class ReferenceGP(object):
    def __init__(self, ora, hyp, kern_types):
        self.ora = ora
        ora._ensure_inv()
        self.ref = ref_loader.load_gp_utils()
        self.kfun = [self.ref._get_kernel_function(k, {"lengthscale": np.asarray(h["lengthscale"]), "variance": h["variance"]})
                     for k, h in zip(kern_types, hyp)]
        self.n_s = len(kern_types)

    def __call__(self, states, actions):
        n, _ = np.shape(states)
        if n > 1:     # as ssm_gpy/gaussian_process.py:142-143
            raise NotImplementedError("Currently do not support multiple state-action pairs to evaluate on.")
        z = np.hstack((np.asarray(states, dtype=np.float64), np.asarray(actions, dtype=np.float64)))
        mu = np.empty((self.n_s, 1))
        var = np.empty((self.n_s, 1))
        for d in range(self.n_s):
            m, v = self.ref.gp_pred(z, self.kfun[d], self.ora.beta[:, d:d + 1], self.ora.x_train, self.ora.inv_K[d])
            mu[d, 0] = np.asarray(m).reshape(-1)[0]
            var[d, 0] = np.asarray(v).reshape(-1)[0]
        return mu, var, self.ora.jacobian(z)[0]
