#!/bin/bash
# round-2 GPU job 22: tcgen05 residual block determinism / accuracy / speed after the L2 prefetch, logits GEMM with 2127 bins
mkdir -p gpurun_out; rm -f gpurun_out/j22_*
timeout 300 python tools/t5_check.py > gpurun_out/j22_t5check.txt 2>&1
cat gpurun_out/j22_t5check.txt | tail -20
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j22_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j22_status.txt
tail -5 gpurun_out/j22_vq.log
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j22_resblock.txt 2>> gpurun_out/j22_resblock.err
done
cat gpurun_out/j22_resblock.txt; tail -3 gpurun_out/j22_resblock.err
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j22_bench_vqvae.json 2> gpurun_out/j22_bench_vqvae.err
cut -c1-200 gpurun_out/j22_bench_vqvae.json; echo
timeout 600 python -m pytest tests/test_gpu_prefill.py -q -k "ragged or oracle" > gpurun_out/j22_ragged.log 2>&1
echo "ragged rc=$?" >> gpurun_out/j22_status.txt
tail -4 gpurun_out/j22_ragged.log
echo "== variant nolg" >> gpurun_out/j22_ab.txt
JK_VARIANT=nolg timeout 300 python tools/step_time.py >> gpurun_out/j22_ab.txt 2>> gpurun_out/j22_ab.err
echo "== default build" >> gpurun_out/j22_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j22_ab.txt 2>> gpurun_out/j22_ab.err
cat gpurun_out/j22_ab.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j22_phase.txt 2>&1
sed -n 5,26p gpurun_out/j22_phase.txt
cat gpurun_out/j22_status.txt
