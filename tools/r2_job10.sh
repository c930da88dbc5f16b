#!/bin/bash
# round-2 GPU job 10: load-pipelining microbenchmark; decode kernel with the statistics poll overlapped with the loads
mkdir -p gpurun_out; rm -f gpurun_out/j10_*
timeout 120 tools/micro/ldbench > gpurun_out/j10_ldbench.txt 2>&1
timeout 200 python tools/step_time.py > gpurun_out/j10_step_time.txt 2> gpurun_out/j10_step_time.err
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j10_phase.txt 2> gpurun_out/j10_phase.err
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py "tests/test_gpu_fullsize_golden.py" -q > gpurun_out/j10_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j10_status.txt
cat gpurun_out/j10_ldbench.txt gpurun_out/j10_step_time.txt; tail -3 gpurun_out/j10_tests.log
