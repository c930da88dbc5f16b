#!/bin/bash
# round-2 GPU job 15 (2 GPUs): the driver's multi-GPU launch of bench.py, both arms
mkdir -p gpurun_out; rm -f gpurun_out/j15_*
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/j15_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/j15_bench_n2.json 2> gpurun_out/j15_bench_n2.err
echo "n2 rc=$?" >> gpurun_out/j15_status.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/j15_ref_n2.json 2> gpurun_out/j15_ref_n2.err
echo "ref n2 rc=$?" >> gpurun_out/j15_status.txt
cat gpurun_out/j15_status.txt; cut -c1-600 gpurun_out/j15_bench_n2.json; tail -3 gpurun_out/j15_bench_n2.err; cut -c1-300 gpurun_out/j15_ref_n2.json
