#!/bin/bash
# round-2 GPU job 17: full GPU suite on the current build + VQ-VAE / prefill timings + smoke
mkdir -p gpurun_out; rm -f gpurun_out/j17_*
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j17_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j17_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j17_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/j17_status.txt
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j17_resblock.txt 2>> gpurun_out/j17_resblock.err
done
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j17_bench_vqvae.json 2> gpurun_out/j17_bench_vqvae.err
timeout 300 python tools/vqvae_profile.py > gpurun_out/j17_vqvae_profile.txt 2>&1
timeout 400 python tools/prefill_time.py > gpurun_out/j17_prefill_1b.txt 2> gpurun_out/j17_prefill.err
JK_WORKLOAD=5b_lyrics JK_N=8 timeout 400 python tools/prefill_time.py > gpurun_out/j17_prefill_5b.txt 2>> gpurun_out/j17_prefill.err
JK_WORKLOAD=small_upsampler timeout 400 python tools/prefill_time.py > gpurun_out/j17_prefill_up.txt 2>> gpurun_out/j17_prefill.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma -s 2 -c 1 -f -o gpurun_out/j17_attn_mma python tools/prefill_time.py > gpurun_out/j17_ncu.log 2>&1
tail -15 gpurun_out/j17_allgpu.log; cat gpurun_out/j17_status.txt; tail -2 gpurun_out/j17_smoke.log; cat gpurun_out/j17_resblock.txt
cut -c1-200 gpurun_out/j17_bench_vqvae.json; echo; grep -h "prefill of\|capacity" gpurun_out/j17_prefill_*.txt
grep "kernel\|Kernel" gpurun_out/j17_vqvae_profile.txt | cut -c1-70,150-200 | head -10
