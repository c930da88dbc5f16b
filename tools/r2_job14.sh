#!/bin/bash
# round-2 GPU job 14: fp16-split tensor-core resblock (A/B against the 3xTF32 kernel), VQ-VAE tests + bench, prefill K-tail case
mkdir -p gpurun_out; rm -f gpurun_out/j14_*
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j14_tests_vqvae.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j14_status.txt
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_fullsize.py tests/test_gpu_prior.py -q > gpurun_out/j14_tests_prefill.log 2>&1
echo "prefill tests rc=$?" >> gpurun_out/j14_status.txt
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j14_resblock.txt 2>> gpurun_out/j14_resblock.err
  JK_C=$c JK_RESBLOCK_TF32=1 timeout 120 python tools/ncu_resblock.py | sed 's/jk_resblock_tc/jk_resblock_tc (3xTF32)/' >> gpurun_out/j14_resblock.txt 2>> gpurun_out/j14_resblock.err
  JK_C=$c JK_FN=jk_resblock_tc JK_DIL=2187 timeout 120 python tools/ncu_resblock.py >> gpurun_out/j14_resblock.txt 2>> gpurun_out/j14_resblock.err
done
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j14_bench_vqvae.json 2> gpurun_out/j14_bench_vqvae.err
JK_RESBLOCK_TF32=1 JK_CONV_EXACT=1 timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j14_bench_vqvae_tf32.json 2>> gpurun_out/j14_bench_vqvae.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_h2 -s 3 -c 1 -f -o gpurun_out/j14_resblock_h2 python tools/ncu_resblock.py > gpurun_out/j14_ncu.log 2>&1
JK_RESBLOCK_TF32=1 timeout 300 ncu --set full --clock-control none -k regex:resblock_tc -s 3 -c 1 -f -o gpurun_out/j14_resblock_tf32 python tools/ncu_resblock.py >> gpurun_out/j14_ncu.log 2>&1
timeout 300 python tools/vqvae_profile.py > gpurun_out/j14_vqvae_profile.txt 2>&1
timeout 400 python tools/prefill_time.py > gpurun_out/j14_prefill_1b.txt 2> gpurun_out/j14_prefill.err
JK_PREFILL_SCALAR_ATTN=1 timeout 400 python tools/prefill_time.py > gpurun_out/j14_prefill_1b_scalar.txt 2>> gpurun_out/j14_prefill.err
grep -h "prefill of" gpurun_out/j14_prefill_1b.txt gpurun_out/j14_prefill_1b_scalar.txt
tail -5 gpurun_out/j14_tests_vqvae.log; tail -5 gpurun_out/j14_tests_prefill.log; cat gpurun_out/j14_status.txt gpurun_out/j14_resblock.txt
cut -c1-300 gpurun_out/j14_bench_vqvae.json; echo; cut -c1-300 gpurun_out/j14_bench_vqvae_tf32.json
