#!/bin/bash
# round-2 GPU job 4: ring fix check (5b / tiny), prefetch on/off, ncu --set full of one decode step
mkdir -p gpurun_out; rm -f gpurun_out/j4_* gpurun_out/parity_r02.jsonl
export JK_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_sampling.py "tests/test_gpu_fullsize_golden.py" -x -q > gpurun_out/j4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j4_status.txt
JK_KV_PREFETCH=0 timeout 300 python tools/step_time.py > gpurun_out/j4_step_time_nopre.txt 2> gpurun_out/j4_step_time.err
JK_KV_PREFETCH=1 timeout 300 python tools/step_time.py > gpurun_out/j4_step_time_pre.txt 2>> gpurun_out/j4_step_time.err
echo "step_time rc=$?" >> gpurun_out/j4_status.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j4_phase.txt 2> gpurun_out/j4_phase.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_decode_step -s 3 -c 1 -f -o gpurun_out/j4_decode python tools/ncu_step.py --steps 5 --pos 4000 > gpurun_out/j4_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/j4_status.txt
ls -la gpurun_out/j4_decode.ncu-rep >> gpurun_out/j4_status.txt
tail -c 800 gpurun_out/j4_tests.log; cat gpurun_out/j4_step_time_nopre.txt gpurun_out/j4_step_time_pre.txt; cat gpurun_out/j4_status.txt
