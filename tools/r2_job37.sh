#!/bin/bash
# round-2 GPU job 37: the whole GPU suite on the final build
mkdir -p gpurun_out; rm -f gpurun_out/j37_*
timeout 165 python -m pytest tests -m gpu -q > gpurun_out/j37_allgpu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/j37_allgpu.log
