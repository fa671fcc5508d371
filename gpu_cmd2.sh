timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; tail -c 1500 gpurun_out/r02_bench_v2.err
B200ORB_FAST_SWEEP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_v2_oldsweep.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_v2.json','gpurun_out/r02_bench_v2_oldsweep.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['stages'].items()}, d['stats'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu_bench.log 2>&1; tail -3 gpurun_out/r02_ncu_bench.log | cut -c1-300
