#!/bin/bash
# round-2 GPU job 29: 5b_lyrics (KS = 1) decode step, A/B of the round's switches on one box
mkdir -p gpurun_out; rm -f gpurun_out/j29_*
for v in alloff noq3 noq7 ""; do
  echo "== variant '$v'" >> gpurun_out/j29_ab.txt
  JK_WORKLOAD=5b_lyrics JK_N=8 JK_VARIANT=$v timeout 400 python tools/step_time.py >> gpurun_out/j29_ab.txt 2>> gpurun_out/j29_ab.err
done
cat gpurun_out/j29_ab.txt; tail -3 gpurun_out/j29_ab.err
