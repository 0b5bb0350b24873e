# Final round-1 measurement pass (1 GPU): tests, smoke, bench for every config, fp64 arm, reference arm, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -c 600 gpurun_out/bench_c4_n1.json
python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
python bench.py --config C2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
python bench.py --config C5 --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
python bench.py --tri-mode 0 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1_fp64dmma.json 2> gpurun_out/bench_c4_n1_fp64dmma.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step' -c 200 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tri_i8m -s 2 -c 1 -o gpurun_out/prof_tri_i8m_c4 -f python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_tri.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kstar_i8 -s 2 -c 1 -o gpurun_out/prof_kstar_i8_c4 -f python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_kstar.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ellipsoid_step -s 2 -c 1 -o gpurun_out/prof_ellipsoid_c4 -f python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_ell.log 2>&1
ls -la gpurun_out | tail -20
