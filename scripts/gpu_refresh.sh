set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
python bench.py --config C2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
python bench.py --config C5 --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
python - <<'PY'
import json
for c in ("c4","c3","c2","c5"):
    j=json.load(open("gpurun_out/bench_%s_n1.json"%c)); r=j["roofline"]
    print(c, "value %.0f e2e %.0f setup %.3f kernel %s tri %.3f frac %.4f parity %s"%(j["value"],j["e2e"]["value"],j["setup_s"],r["kernel"],r["avg_launch_ms"],r["frac"],j["parity"]["max_rel_err"] if j.get("parity") else None))
PY
