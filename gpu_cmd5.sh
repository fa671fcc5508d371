timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_v1.json 2> gpurun_out/r02_bench_n2_v1.err; tail -c 1200 gpurun_out/r02_bench_n2_v1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_n2_v1.json')); print(round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d.get('merge'), d['stats'])
except Exception as e: print('ERR', e)
PY
