#!/bin/bash
# round-2 GPU job 33: smoke() with the tcgen05 residual block added
mkdir -p gpurun_out; rm -f gpurun_out/j33_*
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j33_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/j33_smoke.log
