mkdir -p gpurun_out
M=${1:-3}
for ab in 0 1 2 4 5 6; do SEGP_I8_ABLATE=$ab timeout 300 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --tri-mode $M --no-cpu-baseline > gpurun_out/ab$ab.json 2> gpurun_out/ab$ab.err; python -c "
import json; d=json.load(open('gpurun_out/ab$ab.json')); r=d['roofline']; print('mode $M ablate $ab tri ms', r['avg_launch_ms'], 'step ms', d['ms_per_step'])"; tail -2 gpurun_out/ab$ab.err; done
