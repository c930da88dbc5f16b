#!/bin/bash
# round-2 GPU job 30: t5 residual prefetch (tests, timing, VQ bench), 5b_lyrics step after the round-robin fix, quick decode tests
mkdir -p gpurun_out; rm -f gpurun_out/j30_*
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j30_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j30_status.txt
tail -2 gpurun_out/j30_vq.log
timeout 200 python tools/t5_check.py 2>&1 | grep "^C " | cut -c1-200 > gpurun_out/j30_t5check.txt
for c in 64 32; do JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j30_t5check.txt 2>> gpurun_out/j30_resblock.err; done
cat gpurun_out/j30_t5check.txt
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j30_bench_vqvae.json 2> gpurun_out/j30_bench_vqvae.err
cut -c1-200 gpurun_out/j30_bench_vqvae.json; echo
echo "== 5b_lyrics, default build (round-robin slots in ring-ordered phases)" >> gpurun_out/j30_ab.txt
JK_WORKLOAD=5b_lyrics JK_N=8 timeout 400 python tools/step_time.py >> gpurun_out/j30_ab.txt 2>> gpurun_out/j30_ab.err
echo "== 5b_lyrics, variant noq3" >> gpurun_out/j30_ab.txt
JK_WORKLOAD=5b_lyrics JK_N=8 JK_VARIANT=noq3 timeout 400 python tools/step_time.py >> gpurun_out/j30_ab.txt 2>> gpurun_out/j30_ab.err
echo "== 1b_lyrics, default build" >> gpurun_out/j30_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j30_ab.txt 2>> gpurun_out/j30_ab.err
cat gpurun_out/j30_ab.txt
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_fullsize_golden.py -q > gpurun_out/j30_quick.log 2>&1
echo "decode tests rc=$?" >> gpurun_out/j30_status.txt
tail -3 gpurun_out/j30_quick.log; cat gpurun_out/j30_status.txt
