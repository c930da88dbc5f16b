#!/bin/bash
# round-2 GPU job 5: where does the time go - own work vs waiting (JK_NOWAIT), K-split variants; C4 bench; full GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/j5_*
for v in "default" "JK_NOWAIT=1" "JK_KSPLIT=2" "JK_KSPLIT=1" "JK_NOWAIT=1 JK_KSPLIT=1"; do
  echo "== $v" >> gpurun_out/j5_variants.txt
  if [ "$v" = "default" ]; then timeout 200 python tools/step_time.py >> gpurun_out/j5_variants.txt 2>> gpurun_out/j5_variants.err
  else env $v timeout 200 python tools/step_time.py >> gpurun_out/j5_variants.txt 2>> gpurun_out/j5_variants.err; fi
done
JK_NOWAIT=1 JK_PROFILE=1 timeout 300 python tools/phase_profile.py > gpurun_out/j5_phase_nowait.txt 2> gpurun_out/j5_phase.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j5_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j5_status.txt
timeout 900 python bench.py --workload 5b_lyrics --steps 8 --warmup 3 > gpurun_out/j5_bench_c4.json 2> gpurun_out/j5_bench_c4.err
echo "c4 rc=$?" >> gpurun_out/j5_status.txt
cat gpurun_out/j5_variants.txt; tail -6 gpurun_out/j5_allgpu.log; cat gpurun_out/j5_status.txt
