#!/bin/bash
# round-2 GPU job 27: narrow conv with cp.async staging (tests, VQ-VAE profile + bench), C3 / C4 benches on the current decode kernel
mkdir -p gpurun_out; rm -f gpurun_out/j27_*
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j27_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j27_status.txt
tail -3 gpurun_out/j27_vq.log
timeout 300 python tools/vqvae_profile.py > gpurun_out/j27_vqvae_profile.txt 2>&1
grep "kernel\|Kernel\|CUDA time total" gpurun_out/j27_vqvae_profile.txt | cut -c1-70,150-200 | head -12
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j27_bench_vqvae.json 2> gpurun_out/j27_bench_vqvae.err
cut -c1-200 gpurun_out/j27_bench_vqvae.json; echo
( time timeout 600 python bench.py --workload small_upsampler --steps 8 --warmup 3 > gpurun_out/j27_bench_c3.json 2> gpurun_out/j27_bench_c3.err ) 2> gpurun_out/j27_c3_time.txt
cut -c1-400 gpurun_out/j27_bench_c3.json; echo; tail -3 gpurun_out/j27_c3_time.txt
( time timeout 600 python bench.py --workload 5b_lyrics --steps 8 --warmup 3 > gpurun_out/j27_bench_c4.json 2> gpurun_out/j27_bench_c4.err ) 2> gpurun_out/j27_c4_time.txt
cut -c1-400 gpurun_out/j27_bench_c4.json; echo; tail -3 gpurun_out/j27_c4_time.txt
cat gpurun_out/j27_status.txt
