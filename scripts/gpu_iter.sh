mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -x -q > gpurun_out/i8_tests.log 2>&1; tail -3 gpurun_out/i8_tests.log
for c in C4 C3 C2; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_it.json 2> gpurun_out/bench_${c}_it.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_it.json')); r=d['roofline']; print('$c', 'value', d['value'], 'ms/step', d['ms_per_step'], 'tri ms', r['avg_launch_ms'], 'share', r['share_of_step'], 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/bench_${c}_it.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tri_|kstar|ellipsoid_step' -c 100 --csv --log-file gpurun_out/launches_it.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_it.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split('(')[0]; agg.setdefault(k,[]).append(float(r[-1])/1e6)
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print("%-45s n=%3d avg %7.4f ms share %.4f"%(k,len(v),sum(v)/len(v),sum(v)/tot))
PY
