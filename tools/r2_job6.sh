#!/bin/bash
# round-2 GPU job 6: v2.2 (division-free phases, partial sums published from registers, 16-byte statistics poll)
mkdir -p gpurun_out; rm -f gpurun_out/j6_*
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[full1b_o9]" "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[full5b_o6]" -x -q > gpurun_out/j6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j6_status.txt
timeout 200 python tools/step_time.py > gpurun_out/j6_step_time.txt 2> gpurun_out/j6_step_time.err
JK_NOWAIT=1 timeout 200 python tools/step_time.py > gpurun_out/j6_step_time_nowait.txt 2>> gpurun_out/j6_step_time.err
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j6_phase.txt 2> gpurun_out/j6_phase.err
JK_NOWAIT=1 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j6_phase_nowait.txt 2>> gpurun_out/j6_phase.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jk_decode_step -s 3 -c 1 -f -o gpurun_out/j6_decode python tools/ncu_step.py --steps 5 --pos 4000 > gpurun_out/j6_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/j6_status.txt
tail -3 gpurun_out/j6_tests.log; cat gpurun_out/j6_step_time.txt gpurun_out/j6_step_time_nowait.txt; cat gpurun_out/j6_status.txt
