#!/bin/bash
# One GPU measurement pass (1 x B200) under gpurun: which parts run is chosen by the words on the command line.
#   tests    pytest -m gpu (all)          smoke   __graft_entry__.smoke()
#   bench    bench.py for C4 (quoted configuration, B = 65536) + C4 weak shard (B = 8192) + C2 / C3 / C5
#   ncu      launch list of a C4 step + ncu --set full of the three rollout kernels
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
small)
  for ss in 0 2 4; do $B --config C2 --substreams $ss --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_n1_ss${ss}.json 2> gpurun_out/bench_c2_n1_ss${ss}.err; done
  SEGP_ELL_THREADS=32 $B --config C2 --substreams 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_n1_ss4_e32.json 2> gpurun_out/bench_c2_n1_ss4_e32.err
  SEGP_ELL_THREADS=32 $B --config C2 --substreams 0 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_n1_ss0_e32.json 2> gpurun_out/bench_c2_n1_ss0_e32.err
  timeout 600 python -m pytest tests/test_gpu_precision.py -m gpu -q -k "substream or graph" > gpurun_out/pytest_small.log 2>&1; tail -3 gpurun_out/pytest_small.log ;;
ncu)
  NB="$B --scaling weak --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --no-graph"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step|i8_guard' -c 400 --csv --log-file gpurun_out/launches_c4.csv $NB > gpurun_out/ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:tri_i8m -s 4 -c 1 -o gpurun_out/prof_tri_i8m_c4 -f $NB > gpurun_out/ncu_full_tri.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:kstar_i8 -s 4 -c 1 -o gpurun_out/prof_kstar_i8_c4 -f $NB > gpurun_out/ncu_full_kstar.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ellipsoid_step -s 4 -c 1 -o gpurun_out/prof_ellipsoid_c4 -f $NB > gpurun_out/ncu_full_ell.log 2>&1 ;;
esac
done
ls -la gpurun_out | tail -30
