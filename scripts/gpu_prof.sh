mkdir -p gpurun_out
for ab in 0 5; do SEGP_I8_PROF=1 SEGP_I8_ABLATE=$ab timeout 300 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --tri-mode 3 --no-cpu-baseline 2>&1 >/dev/null | tail -8; done
