#!/bin/bash
# round-2 GPU job 31 (2 GPUs): the driver's launch line at N = 2 (replica per rank, NCCL scatter of labels / gather of codes)
mkdir -p gpurun_out; rm -f gpurun_out/j31_*
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/j31_bench_n2.json 2> gpurun_out/j31_bench_n2.err ) 2> gpurun_out/j31_time.txt
cut -c1-700 gpurun_out/j31_bench_n2.json; echo; tail -3 gpurun_out/j31_time.txt; tail -3 gpurun_out/j31_bench_n2.err
