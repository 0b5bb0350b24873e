set -x
mkdir -p gpurun_out
for m in 2 1; do timeout 300 python bench.py --steps 5 --warmup 3 --tri-mode $m --no-cpu-baseline > gpurun_out/bench_c4_mode$m.json 2> gpurun_out/bench_c4_mode$m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_mode$m.json')); r=d['roofline']; print('mode $m', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['pipe_executed_tops'], r['share_of_step'], d['clocks'])"; tail -3 gpurun_out/bench_c4_mode$m.err; done
ncu --set full --clock-control none --import-source on -k regex:tri_i8x2 -s 2 -c 1 -o gpurun_out/prof_tri_i8x2_c4 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --tri-mode 2 > gpurun_out/ncu_full_i8x2.log 2>&1
