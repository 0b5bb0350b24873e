# C4 weak scaling on N GPUs of one box: bash scripts/gpu_scale.sh N  (run under gpurun --gpus N)
n=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_c4_n$n.json 2> gpurun_out/bench_c4_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_n$n.json')); print('n=$n value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'setup_s', d['setup_s'], d['clocks'])"; tail -2 gpurun_out/bench_c4_n$n.err
