#!/bin/bash
# round-2 GPU job 1: full-size golden parity + the new bench on the three prior workloads (current kernel)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt
timeout 900 python -m pytest tests/test_gpu_fullsize_golden.py -x -q -s > gpurun_out/j1_fullsize_golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/j1_status.txt
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/j1_bench_1b.json 2> gpurun_out/j1_bench_1b.err
echo "bench1b rc=$?" >> gpurun_out/j1_status.txt
timeout 400 python bench.py --workload small_upsampler --steps 8 --warmup 3 > gpurun_out/j1_bench_c3.json 2> gpurun_out/j1_bench_c3.err
echo "c3 rc=$?" >> gpurun_out/j1_status.txt
timeout 900 python bench.py --workload 5b_lyrics --steps 8 --warmup 3 > gpurun_out/j1_bench_c4.json 2> gpurun_out/j1_bench_c4.err
echo "c4 rc=$?" >> gpurun_out/j1_status.txt
tail -c 600 gpurun_out/j1_fullsize_golden.log; cat gpurun_out/j1_status.txt
