mkdir -p gpurun_out
timeout 900 python bench.py --config C5 --steps 3 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; tail -c 2500 gpurun_out/bench_c5_n1.json; tail -5 gpurun_out/bench_c5_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
