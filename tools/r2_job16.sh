#!/bin/bash
# round-2 GPU job 16: 16-warp tap-at-a-time split-precision kernels + coalesced narrow conv
mkdir -p gpurun_out; rm -f gpurun_out/j16_*
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j16_tests_vqvae.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j16_status.txt
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j16_resblock.txt 2>> gpurun_out/j16_resblock.err
  JK_C=$c JK_DIL=2187 timeout 120 python tools/ncu_resblock.py >> gpurun_out/j16_resblock.txt 2>> gpurun_out/j16_resblock.err
done
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j16_bench_vqvae.json 2> gpurun_out/j16_bench_vqvae.err
timeout 300 python tools/vqvae_profile.py > gpurun_out/j16_vqvae_profile.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_h2 -s 3 -c 1 -f -o gpurun_out/j16_resblock_h2 python tools/ncu_resblock.py > gpurun_out/j16_ncu.log 2>&1
tail -5 gpurun_out/j16_tests_vqvae.log; cat gpurun_out/j16_status.txt gpurun_out/j16_resblock.txt
cut -c1-200 gpurun_out/j16_bench_vqvae.json; echo; grep "kernel\|Kernel" gpurun_out/j16_vqvae_profile.txt | cut -c1-70,150-200 | head -12
