mkdir -p gpurun_out
SEGP_TIMELINE=1 timeout 300 python bench.py --config ${1:-C4} --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --overlap 2> gpurun_out/timeline_${1:-C4}.err > /dev/null
grep "segp timeline" gpurun_out/timeline_${1:-C4}.err | tail -18
