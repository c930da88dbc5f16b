#!/bin/bash
# round-2 GPU job 19: A/B of the decode-kernel variants on one box, then the full GPU suite / smoke / benches on the default build
mkdir -p gpurun_out; rm -f gpurun_out/j19_*
for v in base q1 q2 q3 q4 base; do
  echo "== variant $v" >> gpurun_out/j19_ab.txt
  JK_VARIANT=$v timeout 300 python tools/step_time.py >> gpurun_out/j19_ab.txt 2>> gpurun_out/j19_ab.err
done
echo "== default build (all on)" >> gpurun_out/j19_ab.txt
timeout 300 python tools/step_time.py >> gpurun_out/j19_ab.txt 2>> gpurun_out/j19_ab.err
cat gpurun_out/j19_ab.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/j19_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j19_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j19_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/j19_status.txt
JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j19_phase.txt 2>&1
for c in 64 32; do
  JK_C=$c timeout 120 python tools/ncu_resblock.py >> gpurun_out/j19_resblock.txt 2>> gpurun_out/j19_resblock.err
done
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j19_bench_vqvae.json 2> gpurun_out/j19_bench_vqvae.err
timeout 300 python tools/vqvae_profile.py > gpurun_out/j19_vqvae_profile.txt 2>&1
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j19_bench.json 2> gpurun_out/j19_bench.err ) 2> gpurun_out/j19_bench_time.txt
tail -15 gpurun_out/j19_allgpu.log; cat gpurun_out/j19_status.txt; tail -2 gpurun_out/j19_smoke.log; cat gpurun_out/j19_resblock.txt
cut -c1-300 gpurun_out/j19_bench_vqvae.json; echo
cat gpurun_out/j19_bench.json; cat gpurun_out/j19_bench_time.txt
