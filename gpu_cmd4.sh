timeout 300 python -m pytest tests/test_extractor_gpu.py tests/test_pipeline_gpu.py tests/test_mapping_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; tail -c 800 gpurun_out/r02_bench_v3.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_v3.json',):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['stages'].items()}, d['roofline']['stages'].get('mapping'), d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
