"""CPU oracle for the GP one-step posterior + ellipsoid reachability path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product
(``safe_exploration_b200``) never imports this package and raises if its CUDA
library is missing.

Contents
--------
gp_oracle.py     float64 NumPy restatement of the GP half of the path
                 (reference ssm_gpy/gp_models_utils_casadi.py:17-70,160-197 and
                 ssm_gpy/gaussian_process.py:238-263).  GPy and CasADi are not
                 installable here, so this half is a restatement pinned by
                 identities (explicit-inverse form == Cholesky form, analytic
                 Jacobian == finite differences) -- "GP posterior values: parity
                 unpinned by literals", see DESIGN.md.
reach_oracle.py  float64 NumPy restatement of the ellipsoid half (reference
                 gp_reachability.py:19-250, utils.py:108-144,
                 utils_ellipsoid.py:63-94,197-233), single trajectory and
                 vectorised over B.  PINNED: checked against the reference's own
                 unmodified functions (imported through ref_loader.py) and the
                 reference's known-answer tests; golden vectors produced by the
                 reference code are committed under tests/golden/.
score_oracle.py  restatement of the SafeMPC constraint / cost assembly (safempc_simple.py:286-392,
                 488-532, 911-942) around the pinned safety distance.
uprop_oracle.py  closed-form batch restatement of uncertainty_propagation_casadi.py:11-283.
                 PINNED against the reference's own functions (live through the NumPy-backed
                 CasADi shim, and tests/golden/uncertainty_propagation.npz).
ref_loader.py    imports the reference's own gp_reachability / utils /
                 utils_ellipsoid from /root/reference (this container only)
                 through a one-function ``casadi.reshape`` stand-in.
make_golden.py   regenerates tests/golden/*.npz from the reference functions.
"""
