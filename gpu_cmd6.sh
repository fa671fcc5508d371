timeout 300 python -m pytest tests/test_mapping_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; tail -c 800 gpurun_out/r02_bench_v4.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_v4.json',):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['stages'].items()}, d['stats']['mapping_stream_ms_per_step'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
