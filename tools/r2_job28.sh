#!/bin/bash
# round-2 GPU job 28 (final validation): full GPU suite, smoke, the driver's bench command, VQ-VAE bench, ncu captures of the decode kernel
mkdir -p gpurun_out; rm -f gpurun_out/j28_*
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j28_allgpu.log 2>&1
echo "allgpu rc=$?" >> gpurun_out/j28_status.txt
tail -4 gpurun_out/j28_allgpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j28_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/j28_status.txt
tail -1 gpurun_out/j28_smoke.log
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j28_bench.json 2> gpurun_out/j28_bench.err ) 2> gpurun_out/j28_bench_time.txt
cut -c1-300 gpurun_out/j28_bench.json; echo; tail -3 gpurun_out/j28_bench_time.txt
timeout 300 python bench.py --workload vqvae_decode --steps 3 --warmup 1 > gpurun_out/j28_bench_vqvae.json 2> gpurun_out/j28_bench_vqvae.err
cut -c1-200 gpurun_out/j28_bench_vqvae.json; echo
timeout 300 python tools/vqvae_profile.py > gpurun_out/j28_vqvae_profile.txt 2>&1
for pos in 500 4000 8000; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:jk_decode_step -s 3 -c 1 -f -o gpurun_out/j28_decode_p$pos python tools/ncu_step.py --pos $pos > gpurun_out/j28_ncu_$pos.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j28_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/j28_launches.log 2>&1
ls -la gpurun_out/j28_* | cut -c30-120; cat gpurun_out/j28_status.txt
