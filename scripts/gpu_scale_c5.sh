# C5 on N GPUs of one box (B = 16384 candidates per GPU, N = 10000, H = 30, n_s = 10, n_u = 3)
n=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --config C5 --gpus $n --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/bench_c5_n$n.json 2> gpurun_out/bench_c5_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n$n.json')); print('n=$n value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'setup_s', d['setup_s'], d['clocks'])"; tail -2 gpurun_out/bench_c5_n$n.err
