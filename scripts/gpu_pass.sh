#!/bin/bash
# One GPU measurement pass (1 x B200) under gpurun: which parts run is chosen by the words on the command line.
#   tests    pytest -m gpu (all)          smoke   __graft_entry__.smoke()
#   bench    bench.py for C4 (quoted configuration, B = 65536) + C4 weak shard (B = 8192) + C2 / C3 / C5
#   ncu      launch lists (C4 / C3 / C2 steps, C4 factorisation) + ncu --set full of the rollout kernels and the factorisation GEMM
#   setup    factorisation wall times and host-side phase breakdown, tensor-core GEMMs off / on
# Outputs under gpurun_out/ (scratch); scripts/collect_profiles.py copies the judged summaries to profiles/round2/.
mkdir -p gpurun_out
B="python bench.py"
for what in "$@"; do
case $what in
tests)
  timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log ;;
tests_new)
  timeout 1700 python -m pytest tests/test_gpu_precision.py tests/test_gpu_parity.py tests/test_gpu_sampling_mpc.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log ;;
smoke)
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
bench)
  $B --steps 5 --warmup 3 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -c 400 gpurun_out/bench_c4_n1.json
  $B --scaling weak --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_weak_n1.json 2> gpurun_out/bench_c4_weak_n1.err
  $B --scaling weak --i8-digits 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_weak_n1_15products.json 2> gpurun_out/bench_c4_weak_n1_15products.err
  $B --scaling weak --i8-digits 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_weak_n1_10products.json 2> gpurun_out/bench_c4_weak_n1_10products.err
  $B --scaling weak --i8-digits 4 --i8-cluster 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_weak_n1_10products_cl4.json 2> gpurun_out/bench_c4_weak_n1_10products_cl4.err
  $B --scaling weak --i8-digits 4 --tri-mode 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_weak_n1_10products_persistent.json 2> gpurun_out/bench_c4_weak_n1_10products_persistent.err ;;
bench_all)
  $B --config C3 --steps 5 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err
  $B --config C2 --steps 20 --warmup 5 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
  $B --config C2 --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_c2_n1_nograph.json 2> gpurun_out/bench_c2_n1_nograph.err
  $B --config C5 --scaling weak --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
  $B --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err ;;
ncu)
  NB="$B --scaling weak --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --no-graph"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step|i8_guard' -c 400 --csv --log-file gpurun_out/launches_c4.csv $NB > gpurun_out/ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:tri_i8m -s 4 -c 1 -o gpurun_out/prof_tri_i8m_c4 -f $NB > gpurun_out/ncu_full_tri.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:kstar_i8 -s 4 -c 1 -o gpurun_out/prof_kstar_i8_c4 -f $NB > gpurun_out/ncu_full_kstar.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ellipsoid_step -s 4 -c 1 -o gpurun_out/prof_ellipsoid_c4 -f $NB > gpurun_out/ncu_full_ell.log 2>&1
  for c in C2 C3; do
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step|i8_guard' -c 300 --csv --log-file gpurun_out/launches_${c,,}.csv $B --config $c --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches_${c,,}.log 2>&1
  done
  SEGP_FACT_I8=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_setup_c4.csv python scripts/profile_setup.py C4 0 > gpurun_out/ncu_setup.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:gemm_i8d -s 2 -c 1 -o gpurun_out/prof_gemm_i8d_c4 -f python scripts/profile_setup.py C4 0 > gpurun_out/ncu_full_gemm.log 2>&1 ;;
setup)
  for cfg in C3 C4 C5; do for mode in 0 1; do
    SEGP_FACT_TIMING=1 SEGP_FACT_I8=$mode python scripts/profile_setup.py $cfg 3 > gpurun_out/setup_${cfg,,}_fact_i8_$mode.log 2>&1; tail -1 gpurun_out/setup_${cfg,,}_fact_i8_$mode.log
  done; done ;;
esac
done
ls -la gpurun_out | tail -30
