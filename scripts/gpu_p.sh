mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -x -q > gpurun_out/i8_tests.log 2>&1; tail -5 gpurun_out/i8_tests.log
for m in 4 2; do timeout 300 python bench.py --steps 5 --warmup 3 --tri-mode $m --no-cpu-baseline > gpurun_out/bench_c4_mode$m.json 2> gpurun_out/bench_c4_mode$m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_mode$m.json')); r=d['roofline']; print('mode $m', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['pipe_executed_tops'], r['share_of_step'], d['clocks'])"; tail -3 gpurun_out/bench_c4_mode$m.err; done
