"""Run under torchrun on N GPUs: every rank renders its interleaved column bands, the per-tile sample
sums are combined by rtpbr_reduce_tiles (one NCCL sum at tonemap time) and rank 0 compares the
result with the CPU oracle bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from raytracingpbr_b200 import PathTracer, _native as N, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, SPP, B = 320, 192, 6, 8
cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=B, seed=5)
with PathTracer(cfg, objs, cam, tm, device=local) as pt:
    pt.set_shard(rank, world, 32)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(N.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    pt.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    pt.refresh()
    pt.pathtrace(SPP)
    own = pt.image_buffer.to_numpy()
    pt.reduce_tiles(0)
    pt.post_process()
    img = pt.image_buffer.to_numpy()
mine = ((np.arange(W) // 32) % world) == rank
assert (own[~mine] == 0).all() and (own[mine][..., 3] == SPP).all(), "shard ownership"
ok = 1
if rank == 0:
    import common
    oc, oo = common.to_oracle(cfg, cam, objs)
    want = common.po.pathtrace(oc, oo, SPP)
    ok = int(np.array_equal(img, want))
    print(f"multi-GPU check on {world} ranks: reduced image {'==' if ok else '!='} oracle (bit for bit)", flush=True)
t = torch.tensor([ok], device="cuda")
dist.broadcast(t, 0)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
