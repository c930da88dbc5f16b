#!/bin/bash
# round-2 GPU job 20: correctness of the new decode build (logits GEMM, shuffle statistics, slot waits), A/B of its parts, full suite
mkdir -p gpurun_out; rm -f gpurun_out/j20_*
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_fullsize_golden.py tests/test_gpu_prefill.py -q > gpurun_out/j20_quick.log 2>&1
echo "quick rc=$?" >> gpurun_out/j20_status.txt
tail -12 gpurun_out/j20_quick.log
for v in noq5 noq7 nolg noq3 noq3q5; do
  echo "== variant $v" >> gpurun_out/j20_ab.txt
  JK_VARIANT=$v timeout 300 python tools/step_time.py >> gpurun_out/j20_ab.txt 2>> gpurun_out/j20_ab.err
done
echo "== default build" >> gpurun_out/j20_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j20_ab.txt 2>> gpurun_out/j20_ab.err
cat gpurun_out/j20_ab.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j20_phase.txt 2>&1
grep -E "kernel total|logits|per layer" gpurun_out/j20_phase.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j20_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j20_status.txt
tail -12 gpurun_out/j20_allgpu.log; cat gpurun_out/j20_status.txt
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j20_bench_vqvae.json 2> gpurun_out/j20_bench_vqvae.err
cut -c1-200 gpurun_out/j20_bench_vqvae.json
