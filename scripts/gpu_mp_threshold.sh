# where does the persistent contraction stop paying? C4's system at N = 3000 / 4000, tri_mode 4 vs 5
mkdir -p gpurun_out
for n in 3000 4000; do
  for m in 4 5; do
    timeout 300 python bench.py --n-train $n --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode $m > gpurun_out/th_${n}_$m.json 2> gpurun_out/th_${n}_$m.err
    python - <<PY
import json
j=json.load(open("gpurun_out/th_${n}_$m.json")); r=j["roofline"]
print("N=$n mode $m: value %.0f tri_avg %.3f clocks %s"%(j["value"],r["avg_launch_ms"],j["clocks"]))
PY
  done
done
