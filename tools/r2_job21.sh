#!/bin/bash
# round-2 GPU job 21: tcgen05 / TMA residual block (tests + timing vs the mma.sync kernel), decode variants after the epilogue rewrite
mkdir -p gpurun_out; rm -f gpurun_out/j21_*
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q -x > gpurun_out/j21_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j21_status.txt
tail -15 gpurun_out/j21_vq.log
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j21_resblock.txt 2>> gpurun_out/j21_resblock.err
  JK_RESBLOCK_T5=0 JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j21_resblock.txt 2>> gpurun_out/j21_resblock.err
done
JK_DIL=2187 timeout 120 python tools/ncu_resblock.py >> gpurun_out/j21_resblock.txt 2>> gpurun_out/j21_resblock.err
cat gpurun_out/j21_resblock.txt; tail -3 gpurun_out/j21_resblock.err
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j21_bench_vqvae.json 2> gpurun_out/j21_bench_vqvae.err
cut -c1-200 gpurun_out/j21_bench_vqvae.json; echo
for v in nolg noq7 noq7nolg; do
  echo "== variant $v" >> gpurun_out/j21_ab.txt
  JK_VARIANT=$v timeout 300 python tools/step_time.py >> gpurun_out/j21_ab.txt 2>> gpurun_out/j21_ab.err
done
echo "== default build" >> gpurun_out/j21_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j21_ab.txt 2>> gpurun_out/j21_ab.err
cat gpurun_out/j21_ab.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j21_phase.txt 2>&1
sed -n 5,30p gpurun_out/j21_phase.txt
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_fullsize_golden.py -q > gpurun_out/j21_quick.log 2>&1
echo "decode tests rc=$?" >> gpurun_out/j21_status.txt
tail -4 gpurun_out/j21_quick.log; cat gpurun_out/j21_status.txt
