#!/bin/bash
# One gpurun call: parity tests, smoke, bench, tuning sweep, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --kernel simple --no-cpu > gpurun_out/bench_simple.json 2>> gpurun_out/bench.err
for q in 1 4 8 16 32; do
  echo "resolve_min=$q" >> gpurun_out/sweep.log
  RTPBR_RESOLVE_MIN=$q timeout 120 python tools/profile_step.py --passes 4 >> gpurun_out/sweep.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --passes 4 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_pool -s 1 -c 1 -o gpurun_out/prof_r01c -f python tools/profile_step.py --passes 2 --spp 16 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/sweep.log
