"""Factorise one workload's model (segp_set_model + segp_factorize) a few times; run under
`ncu --metrics gpu__time_duration.sum` for the launch list of the setup path, or stand-alone for wall times.

    python scripts/profile_setup.py C4 [repeats]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import safe_exploration_b200 as se  # noqa: E402
from safe_exploration_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = workloads.make(name, batch=96)
torch.cuda.set_device(0)
gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, None, None, kern_types=w.kern_types, hyp=w.hyp, device=0)
if os.environ.get("SEGP_FACT_I8"):
    gp.set_option("fact_i8", int(os.environ["SEGP_FACT_I8"]))
gp.train(w.x_train, w.y_train)
torch.cuda.synchronize()
for i in range(reps):
    t0 = time.perf_counter()
    gp.train(w.x_train, w.y_train)
    torch.cuda.synchronize()
    print("%s factorise #%d: %.1f ms (fact_i8_effective %d)" % (name, i, 1e3 * (time.perf_counter() - t0), gp.get_option("fact_i8_effective")))
gp.close()
