#!/bin/bash
# Ad-hoc experiment pass (1 x B200): short bench lines for the default library and for the variants under
# build_variants/ named on the command line (a variant is copied over safe_exploration_b200/libsegp.so on the box).
mkdir -p gpurun_out
B="python bench.py"
run() {  # tag
  $B --scaling weak --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_c4w_$1.json 2> gpurun_out/exp_c4w_$1.err
  $B --config C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_c3_$1.json 2> gpurun_out/exp_c3_$1.err
  $B --config C2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/exp_c2_$1.json 2> gpurun_out/exp_c2_$1.err
  python - <<PY
import json
for c in ("c4w","c3","c2"):
    try:
        d=json.loads(open("gpurun_out/exp_%s_$1.json"%c).read().strip().splitlines()[-1])
        print("$1", c, "%.0f"%d["value"], "ms/step %.4f"%d["ms_per_step"], "tri_ms", d["roofline"].get("avg_launch_ms"), "share", d["roofline"].get("share_of_step"), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as e: print("$1", c, "ERR", e)
PY
}
run base
cp safe_exploration_b200/libsegp.so /tmp/libsegp_base.so
for v in "$@"; do cp build_variants/libsegp_$v.so safe_exploration_b200/libsegp.so; run $v; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2; done
cp /tmp/libsegp_base.so safe_exploration_b200/libsegp.so
