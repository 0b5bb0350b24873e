mkdir -p gpurun_out
for cfg in "0 3" "1 3" "2 3" "1 1"; do
  set -- $cfg
  echo "== SEGP_CO_PRIO=$1 SEGP_CO_CTAS=$2"
  SEGP_CO_PRIO=$1 SEGP_CO_CTAS=$2 SEGP_TIMELINE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --overlap 2> gpurun_out/tl.err > /dev/null
  grep "segp timeline" gpurun_out/tl.err | tail -18 | grep -E "step 2"
done
