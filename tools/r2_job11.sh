#!/bin/bash
# round-2 GPU job 11: the record run - full GPU suite, bench (driver arguments) for C2 / C3 / C4 / VQ-VAE, ncu artefacts
mkdir -p gpurun_out; rm -f gpurun_out/j11_* gpurun_out/parity_r02.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > gpurun_out/j11_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j11_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j11_status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j11_bench_1b.json 2> gpurun_out/j11_bench_1b.err
echo "bench1b rc=$?" >> gpurun_out/j11_status.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/j11_bench_ref.json 2> gpurun_out/j11_bench_ref.err
echo "benchref rc=$?" >> gpurun_out/j11_status.txt
timeout 400 python bench.py --workload small_upsampler --steps 8 --warmup 3 > gpurun_out/j11_bench_c3.json 2> gpurun_out/j11_bench_c3.err
echo "c3 rc=$?" >> gpurun_out/j11_status.txt
timeout 400 python bench.py --workload small_upsampler --n-samples 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j11_bench_c3_n2.json 2> gpurun_out/j11_bench_c3_n2.err
timeout 900 python bench.py --workload 5b_lyrics --steps 8 --warmup 3 > gpurun_out/j11_bench_c4.json 2> gpurun_out/j11_bench_c4.err
echo "c4 rc=$?" >> gpurun_out/j11_status.txt
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j11_bench_vqvae.json 2> gpurun_out/j11_bench_vqvae.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/j11_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/j11_ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/j11_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_decode_step -s 3 -c 1 -f -o gpurun_out/j11_decode python tools/ncu_step.py --steps 5 --pos 4000 > gpurun_out/j11_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/j11_status.txt
tail -4 gpurun_out/j11_allgpu.log; cat gpurun_out/j11_status.txt; cut -c1-400 gpurun_out/j11_bench_1b.json
