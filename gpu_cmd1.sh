python -m pytest tests -x -q -m gpu 2>&1 | tee gpurun_out/r02_gpu_tests_b.log | tail -25
B200ORB_FAST_SWEEP=1 python -m pytest tests/test_extractor_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; tail -c 2500 gpurun_out/r02_bench_v1.json; tail -5 gpurun_out/r02_bench_v1.err
B200ORB_FAST_SWEEP=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_v1_oldsweep.json 2>/dev/null; python -c "
import json
for f in ('gpurun_out/r02_bench_v1.json','gpurun_out/r02_bench_v1_oldsweep.json'):
    d=json.load(open(f)); print(f, d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['stages'].items()})
"
