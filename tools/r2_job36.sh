#!/bin/bash
# round-2 GPU job 36 (last GPU seconds of the round): prior / prefill / full-size self-consistency tests on the LL-merge build
mkdir -p gpurun_out; rm -f gpurun_out/j36_*
timeout 170 python -m pytest tests/test_gpu_prior.py tests/test_gpu_prefill.py tests/test_gpu_fullsize.py -q -x > gpurun_out/j36_tests.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/j36_tests.log
