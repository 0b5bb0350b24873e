set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4_i8.json 2> gpurun_out/bench_c4_i8.err
cat gpurun_out/bench_c4_i8.json; tail -5 gpurun_out/bench_c4_i8.err
python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_i8.json 2> gpurun_out/bench_c3_i8.err
python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_i8.json 2> gpurun_out/bench_c2_i8.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step' -c 200 --csv --log-file gpurun_out/launches_c4_i8.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_launches_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tri_i8 -s 2 -c 1 -o gpurun_out/prof_tri_i8_c4 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kstar_i8 -s 2 -c 1 -o gpurun_out/prof_kstar_i8_c4 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_kstar_i8.log 2>&1
ls -la gpurun_out | tail -12
