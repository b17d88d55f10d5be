#!/bin/bash
# Round-2 sweep: regeneration policy (batch size / idle-lane trigger), pool size; C1 (family A) and cornell_box.py (family B).
mkdir -p gpurun_out; : > gpurun_out/sweep_r02.log
run() { echo "$*" >> gpurun_out/sweep_r02.log; env "${@:2}" timeout 40 python tools/profile_step.py --passes 3 $1 2>&1 | tail -1 >> gpurun_out/sweep_r02.log; }
run "--size 256 --spp 4" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
for S in "--scene cornell_box_shortest" "--scene cornell_box"; do
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=4
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=16
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=24 RTPBR_REGEN_IDLE=8
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=24 RTPBR_REGEN_IDLE=8 RTPBR_POOL_SLOTS=80 RTPBR_POOL_MIN_BLOCKS=3
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=32 RTPBR_REGEN_IDLE=8 RTPBR_POOL_SLOTS=96 RTPBR_POOL_MIN_BLOCKS=3
run "$S" RTPBR_FIN_MIN=6 RTPBR_REGEN_MIN=16 RTPBR_REGEN_IDLE=8 RTPBR_POOL_SLOTS=80 RTPBR_POOL_MIN_BLOCKS=3
done
cat gpurun_out/sweep_r02.log
