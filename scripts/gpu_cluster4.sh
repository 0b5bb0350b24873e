set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q --timeout 300 -k "c4_model_size" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for cl in 4 2 4 2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 4 --i8-cluster $cl > gpurun_out/cl_C4_$cl.json 2> gpurun_out/cl_C4_$cl.err
  python - <<PY
import json
j=json.load(open("gpurun_out/cl_C4_$cl.json")); r=j["roofline"]
print("cluster $cl: value %.0f ms/step %.2f tri_avg %.3f frac %.4f clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["frac"],j["clocks"]))
PY
done
timeout 300 python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --tri-mode 4 --i8-cluster 4 > gpurun_out/cl_C5_4.json 2> gpurun_out/cl_C5_4.err
python - <<PY
import json
j=json.load(open("gpurun_out/cl_C5_4.json")); r=j["roofline"]
print("C5 cluster 4: value %.0f ms/step %.2f tri_avg %.3f frac %.4f clocks %s"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["frac"],j["clocks"]))
PY
