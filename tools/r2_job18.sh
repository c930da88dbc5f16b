#!/bin/bash
# round-2 GPU job 18 (after the container was re-created): full GPU suite + smoke + VQ-VAE timings + the driver's bench command
mkdir -p gpurun_out; rm -f gpurun_out/j18_*
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j18_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j18_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j18_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/j18_status.txt
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j18_resblock.txt 2>> gpurun_out/j18_resblock.err
done
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j18_bench_vqvae.json 2> gpurun_out/j18_bench_vqvae.err
timeout 300 python tools/vqvae_profile.py > gpurun_out/j18_vqvae_profile.txt 2>&1
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j18_bench.json 2> gpurun_out/j18_bench.err ) 2> gpurun_out/j18_bench_time.txt
tail -15 gpurun_out/j18_allgpu.log; cat gpurun_out/j18_status.txt; tail -2 gpurun_out/j18_smoke.log; cat gpurun_out/j18_resblock.txt
cut -c1-300 gpurun_out/j18_bench_vqvae.json; echo
grep "kernel\|Kernel" gpurun_out/j18_vqvae_profile.txt | cut -c1-70,150-200 | head -10
cat gpurun_out/j18_bench.json; cat gpurun_out/j18_bench_time.txt
