#!/bin/bash
# round-2 GPU job 26: decode kernel after the argument-record fix (logit bias A/B, phase profile), full GPU suite, smoke, benches
mkdir -p gpurun_out; rm -f gpurun_out/j26_*
echo "== default build, logit bias (tensor-core logits)" >> gpurun_out/j26_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j26_ab.txt 2>> gpurun_out/j26_ab.err
echo "== default build, no logit bias (fp32 FMA logits)" >> gpurun_out/j26_ab.txt
JK_LOGIT_BIAS=0 timeout 300 python tools/step_time.py >> gpurun_out/j26_ab.txt 2>> gpurun_out/j26_ab.err
cat gpurun_out/j26_ab.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j26_phase.txt 2>&1
sed -n 5,28p gpurun_out/j26_phase.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j26_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j26_status.txt
tail -6 gpurun_out/j26_allgpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j26_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/j26_status.txt
tail -1 gpurun_out/j26_smoke.log
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j26_bench.json 2> gpurun_out/j26_bench.err ) 2> gpurun_out/j26_bench_time.txt
cat gpurun_out/j26_bench.json | cut -c1-1200; cat gpurun_out/j26_bench_time.txt; cat gpurun_out/j26_status.txt
