# panel-group (L2 residency of the K* planes) sweep for the tcgen05 contraction at C5 and C4
set -x
mkdir -p gpurun_out
for g in 24 12 8; do
  timeout 400 python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 4 --i8-panel-group $g > gpurun_out/pg_C5_4_$g.json 2> gpurun_out/pg_C5_4_$g.err
done
timeout 400 python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 5 --i8-panel-group 12 > gpurun_out/pg_C5_5_12.json 2> gpurun_out/pg_C5_5_12.err
for g in 24 16 12; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 4 --i8-panel-group $g > gpurun_out/pg_C4_4_$g.json 2> gpurun_out/pg_C4_4_$g.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/pg_C*.json")):
    try:
        j=json.load(open(f)); r=j["roofline"]
        print(f, "value %.0f ms/step %.2f tri_avg %.3f share %.3f frac %.4f clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["share_of_step"],r["frac"],j["clocks"]))
    except Exception as e:
        print(f,"failed",e); print(open(f.replace(".json",".err")).read()[-600:])
PY
