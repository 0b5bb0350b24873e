set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
tail -c 3000 gpurun_out/bench_c4_n1.json
python bench.py --config C2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
python bench.py --config C3 --steps 3 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_sumsq|kstar|ellipsoid_step' -c 200 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tri_sumsq -s 2 -c 2 -o gpurun_out/prof_tri_c4 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
