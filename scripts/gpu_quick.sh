# quick iteration: tcgen05 + parity tests, then C4 bench (default schedule and --no-overlap)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $QUICK_ARGS > gpurun_out/q_C4_a.json 2> gpurun_out/q_C4_a.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --overlap $QUICK_ARGS > gpurun_out/q_C4_b.json 2> gpurun_out/q_C4_b.err
python - <<'PY'
import json
for k in ("a","b"):
    try:
        j=json.load(open("gpurun_out/q_C4_%s.json"%k)); r=j["roofline"]
        print(k,j.get("schedule","")[:10],"value %.0f e2e %.0f ms/step %.2f tri_avg %.3f n %d share %.3f achieved %.1f frac %.4f clocks %s"%(j["value"],j["e2e"]["value"],j["ms_per_step"],r["avg_launch_ms"],r["launches_timed"],r["share_of_step"],r["achieved"],r["frac"],j["clocks"]))
    except Exception as e:
        print(k,"failed",e); print(open("gpurun_out/q_C4_%s.err"%k).read()[-1500:])
PY
