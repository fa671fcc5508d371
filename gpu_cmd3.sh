timeout 300 python -m pytest tests/test_extractor_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; tail -c 800 gpurun_out/r02_bench_v2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_v2.json',):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['stages'].items()}, d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu_bench.log 2>&1; tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_quadtree|k_fast_cells|k_orient_desc|k_match_last_fused|k_ocm_scan_keys|k_ocm_apply|k_ocm_bin|k_ocm_centroids|k_resize_g|k_blur7_strip' --launch-skip 30 --launch-count 22 -f -o gpurun_out/r02_full_v2 python tools/ncu_room.py 96 2000 > gpurun_out/r02_ncu_full.log 2>&1; tail -3 gpurun_out/r02_ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
