#!/bin/bash
# round-2 GPU job 7: strong vs weak polled loads, partial sums via shared memory vs from registers
mkdir -p gpurun_out; rm -f gpurun_out/j7_*
cp jukebox_b200/libjkb200.so /tmp/default.so
for v in weak strong weakst; do
  cp variants/$v.so jukebox_b200/libjkb200.so
  for x in 0 1 2; do
    echo "== loads $v, JK_XP_DIRECT=$x" >> gpurun_out/j7_variants.txt
    JK_XP_DIRECT=$x timeout 200 python tools/step_time.py >> gpurun_out/j7_variants.txt 2>> gpurun_out/j7_variants.err
  done
done
cp variants/weak.so jukebox_b200/libjkb200.so
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j7_phase.txt 2> gpurun_out/j7_phase.err
JK_NOWAIT=1 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j7_phase_nowait.txt 2>> gpurun_out/j7_phase.err
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py "tests/test_gpu_fullsize_golden.py" tests/test_gpu_vqvae.py -q > gpurun_out/j7_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j7_status.txt
JK_XP_DIRECT=2 timeout 600 python -m pytest tests/test_gpu_transformer.py "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[full1b_o9]" "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[fullup_o2]" -q > gpurun_out/j7_tests_colpar.log 2>&1
echo "colpar tests rc=$?" >> gpurun_out/j7_status.txt
JK_XP_DIRECT=1 timeout 600 python -m pytest "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[full1b_o9]" -q > gpurun_out/j7_tests_direct.log 2>&1
echo "direct tests rc=$?" >> gpurun_out/j7_status.txt
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j7_bench_vqvae.json 2> gpurun_out/j7_bench_vqvae.err
JK_VQVAE_EXACT=1 timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j7_bench_vqvae_exact.json 2>> gpurun_out/j7_bench_vqvae.err
cat gpurun_out/j7_variants.txt; tail -4 gpurun_out/j7_tests.log; cat gpurun_out/j7_status.txt; cut -c1-300 gpurun_out/j7_bench_vqvae.json
