#!/bin/bash
# round-2 GPU job 34: ncu capture of the final tcgen05 residual block (C = 64, level-0 shape) and of the tap-GEMM conv
mkdir -p gpurun_out; rm -f gpurun_out/j34_*
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resblock_t5 -s 3 -c 1 -f -o gpurun_out/j34_t5 python tools/ncu_resblock.py > gpurun_out/j34_ncu.log 2>&1
tail -1 gpurun_out/j34_ncu.log
