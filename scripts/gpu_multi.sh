#!/bin/bash
# N-GPU pass under `gpurun --gpus N`: product arm (strong scaling: the quoted configuration's B over N GPUs) and the
# reference arm launched the same way (rank 0 alone computes).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -c 300 gpurun_out/bench_c4_n$N.json
$TR bench.py --gpus $N --steps 3 --warmup 3 --scaling weak --no-cpu-baseline > gpurun_out/bench_c4_weak_n$N.json 2> gpurun_out/bench_c4_weak_n$N.err
$TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; tail -c 300 gpurun_out/bench_ref_n$N.json
grep -i "setup" gpurun_out/bench_c4_n$N.err | tail -3
