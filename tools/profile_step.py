"""Minimal driver for ncu: N passes of the hot path on the C1 workload (or a smaller one)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingpbr_b200 import PathTracer, _native as N, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--kernel", default="persistent")
ap.add_argument("--scene", default="cornell_box_shortest", help="preset in raytracingpbr_b200.scenes (bunny_glass, tokyo_ibl, ...)")
a = ap.parse_args()
kern = N.KERNEL_PERSISTENT if a.kernel == "persistent" else N.KERNEL_SIMPLE
cfg, objs, cam, tm = getattr(scenes, a.scene)(a.size, a.size, max_bounces=a.bounces, seed=0, kernel=kern)
with PathTracer(cfg, objs, cam, tm) as pt:
    if a.scene in ("bunny_glass", "tokyo_ibl"):      # same procedural environment as bench.py --workload c2 / c3
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        from raytracingpbr_b200 import ibl
        pt.set_envmap(ibl.process(bench.synthetic_env_u8(), 1.8, 2.2))
    for _ in range(a.passes):
        pt.ctx.flush_l2()
        pt.refresh()
        pt.ctx.set_sample_base(0)
        pt.pathtrace(a.spp)
        pt.post_process()
    pt.sync()
    ms, n = pt.ctx.kernel_time()
    print(f"{n} launches, {ms / n:.3f} ms per launch, {a.size * a.size * a.spp / (ms / n) / 1e3:.1f} Msamples/s")
