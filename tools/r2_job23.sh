#!/bin/bash
# round-2 GPU job 23: tcgen05 residual block after the slot-release fix (determinism, tests), one ncu capture of it; logits GEMM A/B without x_cond
mkdir -p gpurun_out; rm -f gpurun_out/j23_*
timeout 300 python tools/t5_check.py > gpurun_out/j23_t5check.txt 2>&1
grep "^C " gpurun_out/j23_t5check.txt
timeout 600 python -m pytest tests/test_gpu_vqvae.py -q > gpurun_out/j23_vq.log 2>&1
echo "vqvae tests rc=$?" >> gpurun_out/j23_status.txt
tail -3 gpurun_out/j23_vq.log
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j23_resblock.txt 2>> gpurun_out/j23_resblock.err
done
cat gpurun_out/j23_resblock.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_t5 -s 3 -c 1 -f -o gpurun_out/j23_t5 python tools/ncu_resblock.py > gpurun_out/j23_ncu.log 2>&1
tail -2 gpurun_out/j23_ncu.log
echo "== variant nolg, no x_cond" >> gpurun_out/j23_ab.txt
JK_XC_NONE=1 JK_VARIANT=nolg timeout 300 python tools/step_time.py >> gpurun_out/j23_ab.txt 2>> gpurun_out/j23_ab.err
echo "== default build, no x_cond" >> gpurun_out/j23_ab.txt
JK_XC_NONE=1 timeout 300 python tools/step_time.py >> gpurun_out/j23_ab.txt 2>> gpurun_out/j23_ab.err
cat gpurun_out/j23_ab.txt; cat gpurun_out/j23_status.txt
