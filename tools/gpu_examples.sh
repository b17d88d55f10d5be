#!/bin/bash
# Smoke-run the headless examples (small sizes) and keep their PNGs.
mkdir -p gpurun_out/examples
cd examples
for e in "cornell_box/cornell_box_shortest.py --width 256 --height 256 --spp 64" "cornell_box/cornell_box.py --width 240 --height 240 --spp 64 --bounces 16" \
         "cornell_box/cornell_box_v3.py --width 256 --height 256 --spp 64" "scene_demo/tokyo_ibl.py --width 480 --height 270 --spp 64 --bounces 16" \
         "scene_demo/main.py --width 480 --height 270 --spp 64 --bounces 16" "bunny/bunny_sdf_glass.py --width 256 --height 256 --spp 32 --bounces 16"; do
  set -- $e; name=$(basename $1 .py)
  timeout 300 python $e --out ../gpurun_out/examples/$name.png 2>&1 | tail -n 2
done
timeout 300 python src_main.py --frames 64 --out ../gpurun_out/examples/src_main.png 2>&1 | tail -n 2
ls -la ../gpurun_out/examples
