set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tri_i8mp -s 2 -c 1 -o gpurun_out/prof_tri_i8mp_c3 -f python bench.py --config C3 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full_mp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tri_i8m_kernel -s 2 -c 1 -o gpurun_out/prof_tri_i8m_c3 -f python bench.py --config C3 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --tri-mode 4 > gpurun_out/ncu_full_m_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step' -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_launches_c3.log 2>&1
ls -la gpurun_out | tail -5
