#!/bin/bash
# round-2 GPU job 9: are the small per-layer parameters (LN gamma/beta, biases) HBM misses on the critical path?
mkdir -p gpurun_out; rm -f gpurun_out/j9_*
for v in "JK_XP_DIRECT=0" "JK_DEBUG_PARAMS0=1" "JK_DEBUG_PARAMS0=1 JK_NOWAIT=1"; do
  echo "== $v" >> gpurun_out/j9_variants.txt
  env $v timeout 200 python tools/step_time.py >> gpurun_out/j9_variants.txt 2>> gpurun_out/j9_variants.err
done
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j9_phase.txt 2> gpurun_out/j9_phase.err
JK_DEBUG_PARAMS0=1 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j9_phase_params0.txt 2>> gpurun_out/j9_phase.err
cat gpurun_out/j9_variants.txt
