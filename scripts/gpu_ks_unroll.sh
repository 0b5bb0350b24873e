mkdir -p gpurun_out
for U in 8 16; do
  touch safe_exploration_b200/csrc/tri_i8.cu
  SEGP_EXTRA_NVCC_FLAGS="-DSEGP_KS_UNROLL=$U" python -m safe_exploration_b200.build > /dev/null 2>&1
  for c in C4 C3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kstar' -c 30 --csv --log-file gpurun_out/ks_$U.csv python bench.py --config $c --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ks_$U.csv')) if len(r)>10 and r[0].isdigit()]
v=[float(r[-1])/1e6 for r in rows]
print("unroll $U $c kstar avg ms", sum(v)/len(v), len(v))
PY
  done
done
