#!/bin/bash
# compute-sanitizer over the pool kernels (tiny images): memcheck, racecheck (shared-memory stacks + work-queue chunk), initcheck.
mkdir -p gpurun_out
: > gpurun_out/sanitize.log
for scene in cornell_box_shortest bunny_glass tokyo_ibl src_scene; do
  for tool in memcheck racecheck; do
    echo "== $tool $scene" >> gpurun_out/sanitize.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/profile_step.py --scene $scene --size 40 --spp 3 --bounces 6 --passes 1 2>&1 | grep -v "^$" | tail -12 >> gpurun_out/sanitize.log
  done
done
cat gpurun_out/sanitize.log
