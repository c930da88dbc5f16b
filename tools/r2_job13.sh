#!/bin/bash
# round-2 GPU job 13: fp32 path tests, prefill v2 (K tails, enc-dec, whole-window capacity, get_preds) tests + timings
mkdir -p gpurun_out; rm -f gpurun_out/j13_*
timeout 600 python -m pytest tests/test_gpu_f32_path.py -q > gpurun_out/j13_tests_f32.log 2>&1
echo "f32 tests rc=$?" >> gpurun_out/j13_status.txt
timeout 900 python -m pytest tests/test_gpu_prefill_gemm.py tests/test_gpu_prefill.py -q > gpurun_out/j13_tests_prefill.log 2>&1
echo "prefill tests rc=$?" >> gpurun_out/j13_status.txt
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_fullsize_golden.py tests/test_gpu_fullsize.py tests/test_gpu_sampling.py -q > gpurun_out/j13_tests_rest.log 2>&1
echo "rest tests rc=$?" >> gpurun_out/j13_status.txt
timeout 400 python tools/prefill_time.py > gpurun_out/j13_prefill_1b.txt 2> gpurun_out/j13_prefill.err
JK_WORKLOAD=5b_lyrics JK_N=8 timeout 400 python tools/prefill_time.py > gpurun_out/j13_prefill_5b.txt 2>> gpurun_out/j13_prefill.err
JK_WORKLOAD=small_upsampler timeout 400 python tools/prefill_time.py > gpurun_out/j13_prefill_up.txt 2>> gpurun_out/j13_prefill.err
tail -25 gpurun_out/j13_tests_f32.log; tail -25 gpurun_out/j13_tests_prefill.log; tail -8 gpurun_out/j13_tests_rest.log; cat gpurun_out/j13_status.txt
grep -h "prefill of\|stepping\|capacity" gpurun_out/j13_prefill_*.txt
