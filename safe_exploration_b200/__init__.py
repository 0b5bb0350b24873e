"""safe_exploration_b200: the batched GP-posterior + ellipsoid-reachability hot path of
befelix/safe-exploration, as hand-written CUDA for B200 (sm_100a) behind the reference's own plugin API.

    from safe_exploration_b200 import BatchedGPSSM
    from safe_exploration_b200.gp_reachability import onestep_reachability, multistep_reachability

Importing the package does not need a GPU; constructing a model or calling any reachability function
does, and raises otherwise (there is no CPU fallback; the CPU oracle under oracle/ is test-only).
"""
from . import _lib  # noqa: F401
from .ssm import BatchedGPSSM  # noqa: F401
from . import gp_reachability, safempc_sampling, uncertainty_propagation, utils, utils_ellipsoid  # noqa: F401
from .safempc_sampling import SamplingSafeMPC, best_candidate, score_rollouts  # noqa: F401
from .cautious_mpc_sampling import SamplingCautiousMPC  # noqa: F401
from .gp_reachability import (RolloutResult, lin_ellipsoid_safety_distance, multistep_reachability,  # noqa: F401
                              onestep_reachability, pinned_result, rollout)

__version__ = "0.1.0"
