set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import safe_exploration_b200 as se
from safe_exploration_b200 import workloads
for name in ("C4", "C3", "C5"):
    w = workloads.make(name, batch=8)
    gp = se.BatchedGPSSM(w.n_s, w.n_s, w.n_u, w.x_train, w.y_train, kern_types=w.kern_types, hyp=w.hyp)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gp._factorize()
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("%s: N=%d n_s=%d segp_factorize %s ms" % (name, w.n_train, w.n_s, ", ".join("%.1f" % (1e3 * t) for t in ts)))
    gp.close()
PY
