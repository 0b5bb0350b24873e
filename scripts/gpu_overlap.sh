# Pipelined (two-stream half-chunk) schedule against the serial one: tests, then C4/C3/C5 bench lines for both
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for c in C4 C3 C5; do
  st=5; [ $c = C5 ] && st=3
  timeout 600 python bench.py --config $c --steps $st --warmup 3 --no-cpu-baseline > gpurun_out/ov_${c}_serial.json 2> gpurun_out/ov_${c}_serial.err
  timeout 600 python bench.py --config $c --steps $st --warmup 3 --no-cpu-baseline --overlap > gpurun_out/ov_${c}_overlap.json 2> gpurun_out/ov_${c}_overlap.err
done
python - <<'PY'
import json
for c in ("C4","C3","C5"):
    for k in ("overlap","serial"):
        try:
            j=json.load(open("gpurun_out/ov_%s_%s.json"%(c,k))); r=j["roofline"]
            print(c,k,"value %.0f e2e %.0f ms/step %.2f tri_avg %.3f n %d share %.3f achieved %.1f frac %.4f clocks %s"%(j["value"],j["e2e"]["value"],j["ms_per_step"],r["avg_launch_ms"],r["launches_timed"],r["share_of_step"],r["achieved"],r["frac"],j["clocks"]))
        except Exception as e:
            print(c,k,"failed",e)
PY
