#!/bin/bash
# round-2 GPU job 24: tcgen05 residual block: slot release after consumption, two converter groups; variants g1 (one group), g1np (no L2 prefetch)
mkdir -p gpurun_out; rm -f gpurun_out/j24_*
for v in "" g1 g1np; do
  echo "== variant '$v'" >> gpurun_out/j24_t5check.txt
  JK_VARIANT=$v timeout 300 python tools/t5_check.py 2>&1 | grep "^C " | cut -c1-200 >> gpurun_out/j24_t5check.txt
  for c in 64 32; do
    JK_VARIANT=$v JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j24_t5check.txt 2>> gpurun_out/j24_resblock.err
  done
done
cat gpurun_out/j24_t5check.txt
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j24_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j24_status.txt
tail -3 gpurun_out/j24_vq.log
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j24_bench_vqvae.json 2> gpurun_out/j24_bench_vqvae.err
cut -c1-200 gpurun_out/j24_bench_vqvae.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_t5 -s 3 -c 1 -f -o gpurun_out/j24_t5 python tools/ncu_resblock.py > gpurun_out/j24_ncu.log 2>&1
cat gpurun_out/j24_status.txt
