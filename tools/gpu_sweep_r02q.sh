#!/bin/bash
# Round 2: missed rays routed to the regeneration batches -- retune the regeneration triggers.
mkdir -p gpurun_out
L=gpurun_out/sweep_r02q.log
: > $L
run() {
  echo "== ${*}" >> $L
  env "${@}" timeout 120 python tools/profile_step.py --scene cornell_box_shortest --spp 64 --passes 3 2>&1 | tail -1 >> $L
}
run A=0
run RTPBR_REGEN_IDLE=10
run RTPBR_REGEN_IDLE=12
run RTPBR_REGEN_IDLE=14
run RTPBR_REGEN_IDLE=16
run RTPBR_REGEN_IDLE=20
run RTPBR_REGEN_IDLE=12 RTPBR_REGEN_MIN=32
run RTPBR_REGEN_IDLE=12 RTPBR_REGEN_MIN=24
run RTPBR_REGEN_IDLE=12 RTPBR_POOL_SLOTS=80
run RTPBR_REGEN_IDLE=12 RTPBR_FIN_MIN=5
run RTPBR_REGEN_IDLE=12 RTPBR_FIN_MIN=7
run RTPBR_REGEN_IDLE=16 RTPBR_REGEN_MIN=32
cat $L
