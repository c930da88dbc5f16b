#!/bin/bash
# round-2 GPU job 25: tcgen05 residual block with the coalesced output epilogue (variants: 1 / 2 converter groups); logit bias route for 1b_lyrics
mkdir -p gpurun_out; rm -f gpurun_out/j25_*
for v in "" g1 g2; do
  echo "== variant '$v'" >> gpurun_out/j25_t5check.txt
  JK_VARIANT=$v timeout 300 python tools/t5_check.py 2>&1 | grep "^C " | cut -c1-200 >> gpurun_out/j25_t5check.txt
  for c in 64 32; do
    JK_VARIANT=$v JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j25_t5check.txt 2>> gpurun_out/j25_resblock.err
  done
done
cat gpurun_out/j25_t5check.txt
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j25_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j25_status.txt
tail -3 gpurun_out/j25_vq.log
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j25_bench_vqvae.json 2> gpurun_out/j25_bench_vqvae.err
cut -c1-200 gpurun_out/j25_bench_vqvae.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_t5 -s 3 -c 1 -f -o gpurun_out/j25_t5 python tools/ncu_resblock.py > gpurun_out/j25_ncu.log 2>&1
echo "== default build, logit bias" >> gpurun_out/j25_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j25_ab.txt 2>> gpurun_out/j25_ab.err
echo "== default build, no logit bias (fp32 FMA logits)" >> gpurun_out/j25_ab.txt
JK_LOGIT_BIAS=0 timeout 300 python tools/step_time.py >> gpurun_out/j25_ab.txt 2>> gpurun_out/j25_ab.err
cat gpurun_out/j25_ab.txt; tail -3 gpurun_out/j25_ab.err
timeout 900 python -m pytest tests/test_gpu_prior.py tests/test_gpu_prefill.py tests/test_gpu_fullsize.py -q > gpurun_out/j25_prior.log 2>&1
echo "prior tests rc=$?" >> gpurun_out/j25_status.txt
tail -4 gpurun_out/j25_prior.log; cat gpurun_out/j25_status.txt
