"""torchrun probe: per-launch kernel time of the sharded C1 step (a) with torch.distributed up, (b) after our own
ncclCommInitRank, (c) with the per-step tile reduce -- to find what costs 6 % per rank in bench.py at N > 1."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingpbr_b200 import PathTracer, _native as N, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, objs, cam, tm = scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)
pt = PathTracer(cfg, objs, cam, tm, device=local)
ctx = pt.ctx
ctx.set_shard(rank, world, 4)


def measure(tag, reduce=False, passes=3, flush=True):
    for i in range(passes + 1):
        if flush:
            ctx.flush_l2()
        ctx.refresh()
        ctx.set_sample_base(0)
        ctx.pathtrace(64 * world)
        if reduce:
            ctx.reduce_tiles(0)
        ctx.sync()
        if i == 0:
            ctx.kernel_time()
    ms, n = ctx.kernel_time()
    dist.barrier()
    print(f"rank {rank} {tag}: {ms / n:.2f} ms per launch", flush=True)


measure("torch.distributed up, no rtpbr communicator")
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(N.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
ctx.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
measure("after rtpbr nccl_init, no reduce")
measure("with reduce_tiles each step", reduce=True)
measure("with reduce_tiles, no L2 flush", reduce=True, flush=False)
measure("no reduce again")
dist.destroy_process_group()
