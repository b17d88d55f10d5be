#!/bin/bash
# Round 2: resolve threshold (runtime parameter) on the final kernels.
mkdir -p gpurun_out
L=gpurun_out/sweep_r02p.log
: > $L
run() {   # scene spp bounces env...
  echo "== $1 spp=$2 b=$3 ${*:4}" >> $L
  env "${@:4}" timeout 120 python tools/profile_step.py --scene $1 --spp $2 --bounces $3 --passes 3 2>&1 | tail -1 >> $L
}
run cornell_box_shortest 64 8 A=0
for r in 12 16 20 24 28 32; do run cornell_box_shortest 64 8 RTPBR_RESOLVE_MIN=$r; done
run cornell_box_shortest 64 8 A=0
for r in 16 24 32; do run tokyo_ibl 64 8 RTPBR_RESOLVE_MIN=$r; done
run tokyo_ibl 64 8 A=0
for r in 16 24 32; do run bunny_glass 32 16 RTPBR_RESOLVE_MIN=$r; done
run bunny_glass 32 16 A=0
cat $L
