set -x
mkdir -p gpurun_out
for c in C3 C4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --overlap > gpurun_out/co_${c}_overlap.json 2> gpurun_out/co_${c}_overlap.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/co_C*_overlap.json")):
    try:
        j=json.load(open(f)); r=j["roofline"]
        print(f, "value %.0f ms/step %.2f tri_avg %.3f n %d share %.3f kernel %s finite %s clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["launches_timed"],r["share_of_step"],r["kernel"],j["all_finite"],j["clocks"]))
    except Exception as e:
        print(f,"failed",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
