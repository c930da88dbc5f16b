#!/bin/bash
# round-2 GPU job 3: LL decode kernel v2.1 - all GPU tests, timing, bench, C3/C4 workloads
mkdir -p gpurun_out; rm -f gpurun_out/j3_* gpurun_out/parity_r02.jsonl
export JK_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_prior.py tests/test_gpu_sampling.py tests/test_gpu_prefill.py -x -q -s > gpurun_out/j3_small.log 2>&1
echo "small rc=$?" >> gpurun_out/j3_status.txt
timeout 900 python -m pytest tests/test_gpu_fullsize_golden.py -q -s > gpurun_out/j3_fullsize_golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/j3_status.txt
timeout 300 python tools/step_time.py > gpurun_out/j3_step_time.txt 2> gpurun_out/j3_step_time.err
echo "step_time rc=$?" >> gpurun_out/j3_status.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j3_phase.txt 2> gpurun_out/j3_phase.err
echo "phase rc=$?" >> gpurun_out/j3_status.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j3_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j3_status.txt
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/j3_bench_1b.json 2> gpurun_out/j3_bench_1b.err
echo "bench1b rc=$?" >> gpurun_out/j3_status.txt
timeout 400 python bench.py --workload small_upsampler --steps 8 --warmup 3 > gpurun_out/j3_bench_c3.json 2> gpurun_out/j3_bench_c3.err
echo "c3 rc=$?" >> gpurun_out/j3_status.txt
timeout 900 python bench.py --workload 5b_lyrics --steps 8 --warmup 3 > gpurun_out/j3_bench_c4.json 2> gpurun_out/j3_bench_c4.err
echo "c4 rc=$?" >> gpurun_out/j3_status.txt
tail -c 1200 gpurun_out/j3_small.log; tail -c 600 gpurun_out/j3_fullsize_golden.log; cat gpurun_out/j3_step_time.txt; tail -5 gpurun_out/j3_allgpu.log; cat gpurun_out/j3_status.txt
