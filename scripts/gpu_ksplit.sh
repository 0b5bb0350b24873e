# K* kernel: split count sweep (grid granularity / tail effect), C4 and C3
set -x
mkdir -p gpurun_out
for ks in 0 10 14 20 40; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --ksplit $ks > gpurun_out/ks_C4_$ks.json 2> gpurun_out/ks_C4_$ks.err
done
for ks in 0 8 16; do
  timeout 600 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --ksplit $ks > gpurun_out/ks_C3_$ks.json 2> gpurun_out/ks_C3_$ks.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ks_C*.json")):
    try:
        j=json.load(open(f)); r=j["roofline"]
        print(f, "value %.0f ms/step %.2f tri_avg %.3f share %.3f non-tri ms/step %.2f"%(j["value"],j["ms_per_step"],r["avg_launch_ms"],r["share_of_step"],j["ms_per_step"]*(1-r["share_of_step"])))
    except Exception as e:
        print(f,"failed",e)
PY
