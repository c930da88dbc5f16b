#!/bin/bash
# round-2 GPU job 32: ncu launch list of the bench command's timed region (final build)
mkdir -p gpurun_out; rm -f gpurun_out/j32_*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/j32_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/j32_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/j32_launches.csv; tail -2 gpurun_out/j32_ncu_launches.log | cut -c1-200
