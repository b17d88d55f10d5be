#!/bin/bash
# Round-2 ncu evidence: launch list of the bench step, one --set full capture of the dominant kernel at C1 (64 spp),
# C2 (64 spp) and C3 (32 spp).  Summaries are extracted here afterwards with tools/ncu_summary.py.
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-blocks > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_pool -s 1 -c 1 -o gpurun_out/${TAG}_prof_c1 -f python tools/profile_step.py --passes 2 --spp 64 > gpurun_out/${TAG}_ncu_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_pool -s 1 -c 1 -o gpurun_out/${TAG}_prof_c2 -f python tools/profile_step.py --scene bunny_glass --bounces 16 --passes 2 --spp 64 > gpurun_out/${TAG}_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pathtrace_pool -s 1 -c 1 -o gpurun_out/${TAG}_prof_c3 -f python tools/profile_step.py --scene tokyo_ibl --passes 2 --spp 32 > gpurun_out/${TAG}_ncu_c3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_c1.log gpurun_out/${TAG}_ncu_c2.log gpurun_out/${TAG}_ncu_c3.log
