#!/bin/bash
# round-2 GPU job 35: attention split-part merge over LL words (fixed merger, no ticket): A/B against the ticket version, parity tests at BASELINE geometry
mkdir -p gpurun_out; rm -f gpurun_out/j35_*
echo "== ticket merge (variant tk)" >> gpurun_out/j35_ab.txt
JK_VARIANT=tk timeout 200 python tools/step_time.py >> gpurun_out/j35_ab.txt 2>> gpurun_out/j35_ab.err
echo "== LL merge (default build)" >> gpurun_out/j35_ab.txt
timeout 200 python tools/step_time.py >> gpurun_out/j35_ab.txt 2>> gpurun_out/j35_ab.err
cat gpurun_out/j35_ab.txt
timeout 600 python -m pytest tests/test_gpu_fullsize_golden.py tests/test_gpu_transformer.py -q > gpurun_out/j35_quick.log 2>&1
echo "decode tests rc=$?" >> gpurun_out/j35_status.txt
tail -4 gpurun_out/j35_quick.log; cat gpurun_out/j35_status.txt
