"""Single-GPU probe: kernel time of the C1 workload unsharded vs as rank r of an emulated N-rank shard
(same samples per GPU), to separate sharding cost from box-to-box variance in the scaling bench."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingpbr_b200 import PathTracer, _native as N, scenes  # noqa: E402


def run(world, rank, band, spp, passes=3):
    cfg, objs, cam, tm = scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)
    with PathTracer(cfg, objs, cam, tm) as pt:
        if world > 1:
            pt.ctx.set_shard(rank, world, band)
        for i in range(passes + 1):
            pt.ctx.flush_l2()
            pt.refresh()
            pt.ctx.set_sample_base(0)
            pt.pathtrace(spp)
            pt.sync()
            if i == 0:
                pt.ctx.kernel_time()
        ms, n = pt.ctx.kernel_time()
        return ms / n


for world, rank, band, spp in [(1, 0, 4, 64), (2, 0, 4, 128), (2, 1, 4, 128), (8, 3, 4, 512), (2, 0, 64, 128), (1, 0, 4, 128)]:
    print(f"world {world} rank {rank} band {band} spp {spp}: {run(world, rank, band, spp):.2f} ms per launch", flush=True)
