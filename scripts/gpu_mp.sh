# persistent folded contraction (tri_mode 5): tcgen05 tests, then C4 / C3 / C5 / C2 bench lines against tri_mode 4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q --timeout 180 > gpurun_out/pytest_mp.log 2>&1; tail -5 gpurun_out/pytest_mp.log
for c in C4 C3 C2; do
  for m in 5 4; do
    timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode $m > gpurun_out/mp_${c}_$m.json 2> gpurun_out/mp_${c}_$m.err
  done
done
timeout 300 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 5 > gpurun_out/mp_C5_5.json 2> gpurun_out/mp_C5_5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/mp_C*.json")):
    try:
        j=json.load(open(f)); r=j["roofline"]
        print(f, "value %.0f ms/step %.2f tri_avg %.3f share %.3f frac %.4f finite %s clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["share_of_step"],r["frac"],j["all_finite"],j["clocks"]))
    except Exception as e:
        print(f,"failed",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
