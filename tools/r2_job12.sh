#!/bin/bash
# round-2 GPU job 12: A/B on one box - v2.1, HEAD of the record run, v2.4 (double-buffered MMA fragments, shuffle-reduced statistics);
# then the parity suites + the new fp32-path tests on the in-tree build (= v2.4)
mkdir -p gpurun_out; rm -f gpurun_out/j12_*
for v in v21 head v24 head v24; do
  echo "== $v" >> gpurun_out/j12_variants.txt
  JK_VARIANT=$v timeout 200 python tools/step_time.py >> gpurun_out/j12_variants.txt 2>> gpurun_out/j12_variants.err
done
for v in v21 head v24; do
  echo "== 5b_lyrics $v" >> gpurun_out/j12_variants.txt
  JK_VARIANT=$v JK_WORKLOAD=5b_lyrics JK_N=8 timeout 300 python tools/step_time.py >> gpurun_out/j12_variants.txt 2>> gpurun_out/j12_variants.err
  echo "== small_upsampler $v" >> gpurun_out/j12_variants.txt
  JK_VARIANT=$v JK_WORKLOAD=small_upsampler timeout 200 python tools/step_time.py >> gpurun_out/j12_variants.txt 2>> gpurun_out/j12_variants.err
done
timeout 600 python -m pytest tests/test_gpu_f32_path.py -q -x > gpurun_out/j12_tests_f32.log 2>&1
echo "f32 tests rc=$?" >> gpurun_out/j12_status.txt
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_prefill.py tests/test_gpu_fullsize_golden.py tests/test_gpu_fullsize.py -q > gpurun_out/j12_tests_v24.log 2>&1
echo "v24 tests rc=$?" >> gpurun_out/j12_status.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j12_phase_v24.txt 2> gpurun_out/j12_phase.err
cat gpurun_out/j12_variants.txt; tail -30 gpurun_out/j12_tests_f32.log; tail -3 gpurun_out/j12_tests_v24.log; cat gpurun_out/j12_status.txt
