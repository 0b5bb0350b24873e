set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "pipelined" > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
bash scripts/gpu_timeline.sh C4 | tail -10
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --overlap > gpurun_out/co_C4_overlap.json 2> gpurun_out/co_C4_overlap.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/co_C4_serial.json 2> gpurun_out/co_C4_serial.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/co_C4_*.json")):
    try:
        j=json.load(open(f)); r=j["roofline"]
        print(f, "value %.0f ms/step %.2f tri_avg %.3f n %d share %.3f kernel %s finite %s clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["launches_timed"],r["share_of_step"],r["kernel"],j["all_finite"],j["clocks"]))
    except Exception as e:
        print(f,"failed",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
