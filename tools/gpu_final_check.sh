#!/bin/bash
# (kept for the record: the last full GPU call of round 2, run as `gpurun -- bash tools/gpu_final_check.sh`)
# last GPU call of the round: parity of the formulations behind B200ORB_EXPERIMENTAL (+ the new dynm_* kernels), launch-shape
# A/B in one process, then the bench line with the configuration that survived.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
date +%s > gpurun_out/f0
# attribution runs (one bit each) beside the full suite; correctness only, so sharing the GPU is fine
( B200ORB_EXPERIMENTAL=1 timeout 170 python -m pytest tests/test_extractor_gpu.py -m gpu -q > gpurun_out/fin_tests_bit0.log 2>&1; echo "rc=$?" >> gpurun_out/fin_tests_bit0.log ) &
( B200ORB_EXPERIMENTAL=2 timeout 170 python -m pytest tests/test_extractor_gpu.py -m gpu -q > gpurun_out/fin_tests_bit1.log 2>&1; echo "rc=$?" >> gpurun_out/fin_tests_bit1.log ) &
( B200ORB_EXPERIMENTAL=3 timeout 190 python -m pytest tests -m gpu -q > gpurun_out/fin_tests_all.log 2>&1; echo "rc=$?" >> gpurun_out/fin_tests_all.log )
wait
date +%s > gpurun_out/f1
timeout 90 python tools/tune_extractor.py 96 2000 > gpurun_out/fin_tune.log 2>&1
date +%s > gpurun_out/f2
eval "$(python - <<'PY'
import json, re
def rc(p):
    try:
        return int(re.findall(r"rc=(\d+)", open(p).read())[-1])
    except Exception:
        return 1
mask = 0
if rc("gpurun_out/fin_tests_all.log") == 0:
    mask = 3
else:
    mask = (1 if rc("gpurun_out/fin_tests_bit0.log") == 0 else 0) | (2 if rc("gpurun_out/fin_tests_bit1.log") == 0 else 0)
wpc, minb = 8, 2
try:
    b = json.load(open("gpurun_out/tune.json"))["best"]
    wpc, minb = int(b["fast_wpc"]), int(b["qt_minb"])
    if b.get("exp_evaluated"):
        mask &= int(b["exp_mask"])
except Exception:
    pass
print("export B200ORB_EXPERIMENTAL=%d B200ORB_FAST_WPC=%d B200ORB_QT_MINB=%d" % (mask, wpc, minb))
PY
)"
echo "chosen: EXPERIMENTAL=$B200ORB_EXPERIMENTAL WPC=$B200ORB_FAST_WPC QT_MINB=$B200ORB_QT_MINB" > gpurun_out/fin_chosen.txt
timeout 120 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
date +%s > gpurun_out/f3
tail -3 gpurun_out/fin_tests_all.log; tail -2 gpurun_out/fin_tests_bit0.log; tail -2 gpurun_out/fin_tests_bit1.log; tail -2 gpurun_out/fin_tune.log; cat gpurun_out/fin_chosen.txt; head -c 300 gpurun_out/fin_bench.json
