#!/bin/bash
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT"
timeout 100 python -m pytest tests/test_zz_tuning_gpu.py tests/test_extractor_gpu.py -m gpu -q > gpurun_out/last_tests.log 2>&1; echo "rc=$?" >> gpurun_out/last_tests.log
tail -4 gpurun_out/last_tests.log
