#!/bin/bash
# round-2 GPU job 2: first run of the LL decode kernel - small parity tests, full-size golden parity, timing
mkdir -p gpurun_out; rm -f gpurun_out/j2_*
export JK_VERBOSE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j2_smi.txt
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py -x -q -s > gpurun_out/j2_small.log 2>&1
echo "small rc=$?" >> gpurun_out/j2_status.txt
timeout 900 python -m pytest tests/test_gpu_fullsize_golden.py -q -s > gpurun_out/j2_fullsize_golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/j2_status.txt
timeout 300 python tools/step_time.py > gpurun_out/j2_step_time.txt 2> gpurun_out/j2_step_time.err
echo "step_time rc=$?" >> gpurun_out/j2_status.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j2_phase.txt 2> gpurun_out/j2_phase.err
echo "phase rc=$?" >> gpurun_out/j2_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j2_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j2_status.txt
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/j2_bench_1b.json 2> gpurun_out/j2_bench_1b.err
echo "bench1b rc=$?" >> gpurun_out/j2_status.txt
tail -c 1500 gpurun_out/j2_small.log; tail -c 800 gpurun_out/j2_fullsize_golden.log; cat gpurun_out/j2_step_time.txt; cat gpurun_out/j2_status.txt
