#!/bin/bash
# Round-2 sweep on the PBR scenes: resident CTAs / block size (registers per thread vs warps per SM).
mkdir -p gpurun_out; : > gpurun_out/sweep_r02b.log
run() { echo "$*" >> gpurun_out/sweep_r02b.log; env "${@:2}" timeout 60 python tools/profile_step.py --passes 3 $1 2>&1 | tail -1 >> gpurun_out/sweep_r02b.log; }
for S in "--scene tokyo_ibl" "--scene scene_demo" "--scene src_scene"; do
run "$S" A=0
run "$S" RTPBR_POOL_MIN_BLOCKS=3
run "$S" RTPBR_POOL_MIN_BLOCKS=2
run "$S" RTPBR_POOL_BLOCK=320 RTPBR_POOL_MIN_BLOCKS=2 RTPBR_POOL_SLOTS=56
run "$S" RTPBR_POOL_BLOCK=192 RTPBR_POOL_MIN_BLOCKS=4
run "$S" RTPBR_POOL_BLOCK=384 RTPBR_POOL_MIN_BLOCKS=2 RTPBR_POOL_SLOTS=56
done
cat gpurun_out/sweep_r02b.log
