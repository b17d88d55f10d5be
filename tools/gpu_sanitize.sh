#!/bin/bash
# compute-sanitizer over the pool kernels (tiny images): memcheck, racecheck (shared-memory stacks + work-queue chunk).
# Round 2: also the fast-region / regeneration-batch / finish-threshold build forced on for the PBR Cornell scenes, and
# the single-process multi-GPU path is covered by tests/test_multi_gpu.py.
mkdir -p gpurun_out
: > gpurun_out/sanitize.log
run() {
  echo "== $1 $2 ${*:3}" >> gpurun_out/sanitize.log
  env "${@:3}" timeout 600 compute-sanitizer --tool $1 --print-limit 5 python tools/profile_step.py --scene $2 --size 40 --spp 3 --bounces 6 --passes 1 2>&1 | grep -v "^$" | tail -6 >> gpurun_out/sanitize.log
}
for tool in memcheck racecheck; do
  for scene in cornell_box_shortest bunny_glass tokyo_ibl src_scene; do run $tool $scene A=0; done
  run $tool cornell_box RTPBR_JIT_FAST=1 RTPBR_JIT_BBOX=1 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8 RTPBR_FIN_MIN=4
  run $tool cornell_box_v3 RTPBR_JIT_FAST=1 RTPBR_JIT_BBOX=1 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8 RTPBR_FIN_MIN=4
done
cat gpurun_out/sanitize.log
