#!/bin/bash
# round-2 GPU job 8: bisect v2.1 vs v2.2 (+ modes); VQ-VAE tensor-core block: test + per-kernel launch list
mkdir -p gpurun_out; rm -f gpurun_out/j8_*
cp jukebox_b200/libjkb200.so /tmp/cur.so
echo "== v21" >> gpurun_out/j8_variants.txt
cp variants/v21.so jukebox_b200/libjkb200.so
timeout 200 python tools/step_time.py >> gpurun_out/j8_variants.txt 2>> gpurun_out/j8_variants.err
cp /tmp/cur.so jukebox_b200/libjkb200.so
for x in 0 1 2; do
  echo "== current, JK_XP_DIRECT=$x" >> gpurun_out/j8_variants.txt
  JK_XP_DIRECT=$x timeout 200 python tools/step_time.py >> gpurun_out/j8_variants.txt 2>> gpurun_out/j8_variants.err
done
JK_XP_DIRECT=2 timeout 600 python -m pytest tests/test_gpu_transformer.py "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[full1b_o9]" "tests/test_gpu_fullsize_golden.py::test_decode_at_baseline_geometry_matches_reference[fullup_o2]" -q > gpurun_out/j8_tests_colpar.log 2>&1
echo "colpar tests rc=$?" >> gpurun_out/j8_status.txt
JK_XP_DIRECT=2 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j8_phase_colpar.txt 2> gpurun_out/j8_phase.err
JK_XP_DIRECT=0 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j8_phase_xp0.txt 2>> gpurun_out/j8_phase.err
timeout 300 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j8_tests_vqvae.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j8_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j8_vqvae_launches.csv python bench.py --workload vqvae_decode --steps 1 --warmup 1 > gpurun_out/j8_vqvae_ncu.log 2>&1
cat gpurun_out/j8_variants.txt; tail -4 gpurun_out/j8_tests_colpar.log; tail -4 gpurun_out/j8_tests_vqvae.log; cat gpurun_out/j8_status.txt
