#!/bin/bash
# Tests, headline bench, informational workloads, ncu launch list + one full capture at the C1 configuration.
set -x
TAG=${1:-r01d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --workload c3 --steps 3 > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --workload c2 --steps 2 > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --passes 4 > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_pool -s 1 -c 1 -o gpurun_out/prof_$TAG -f python tools/profile_step.py --passes 2 --spp 64 > gpurun_out/ncu_full.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_c3.json; cat gpurun_out/bench_c2.json; tail -n 5 gpurun_out/bench.err
