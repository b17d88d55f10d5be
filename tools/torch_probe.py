"""Single-GPU probe: does having torch / torch.distributed(NCCL) alive in the process slow the pool kernel?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingpbr_b200 import PathTracer, _native as N, scenes  # noqa: E402

cfg, objs, cam, tm = scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)
pt = PathTracer(cfg, objs, cam, tm, device=0)
ctx = pt.ctx


def measure(tag, passes=3):
    for i in range(passes + 1):
        ctx.flush_l2()
        ctx.refresh()
        ctx.set_sample_base(0)
        ctx.pathtrace(64)
        ctx.sync()
        if i == 0:
            ctx.kernel_time()
    ms, n = ctx.kernel_time()
    print(f"{tag}: {ms / n:.2f} ms per launch", flush=True)


def limits(tag):
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    out = []
    for name, k in (("stack", 0), ("printf_fifo", 1), ("malloc_heap", 2)):
        v = C.c_size_t()
        rt.cudaDeviceGetLimit(C.byref(v), k)
        out.append(f"{name}={v.value}")
    print(tag, " ".join(out), flush=True)


measure("no torch")
try:
    limits("limits:")
except OSError as e:
    print("cudart not loadable:", e)
import torch  # noqa: E402
torch.cuda.init()
x = torch.zeros(1, device="cuda")
measure("torch imported, CUDA initialised")
y = torch.randn(4096, 4096, device="cuda") @ torch.randn(4096, 4096, device="cuda")
torch.cuda.synchronize()
measure("after a cuBLAS GEMM")
import torch.distributed as dist  # noqa: E402
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29544")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
dist.barrier()
measure("torch.distributed (nccl) initialised")
try:
    limits("limits:")
except OSError:
    pass
dist.destroy_process_group()
measure("process group destroyed")
