"""N > 1 on real GPUs (skipped on boxes with one GPU): bit-level checks of the sharded render + NCCL tile reduce.

  * MultiPathTracer: ONE process drives all GPUs (rtpbr_multi_*, communicators created inside the library).
  * one process per GPU under torchrun, the launch the driver uses for bench.py (tools/multi_gpu_check.py).

Both compare the reduced image with the CPU oracle bit for bit, and with the reference-source golden columns of
BASELINE configs[1] at their real size."""
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from raytracingpbr_b200 import MultiPathTracer, PathTracer, _native as N, scenes

pytestmark = pytest.mark.gpu


def _gpus():
    return N.device_count()


@pytest.mark.parametrize("band", [4, 32])
def test_multi_path_tracer_equals_the_oracle_bit_for_bit(band):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    W, H, SPP = 320, 192, 6
    cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=8, seed=5)
    with MultiPathTracer(cfg, objs, cam, tm, devices=range(n), band=band) as mpt:
        mpt.refresh()
        mpt.pathtrace(SPP)
        mpt.sync()
        parts = [c.download(N.BUF_IMAGE_BUFFER) for c in mpt.ctx.ranks]
        mpt.post_process()                       # reduce onto GPU 0 + tonemap
        img = mpt.image_buffer.to_numpy()
        pix = mpt.image_pixels.to_numpy()
        with pytest.raises(N.RtpbrError):        # GPU 0 now holds the sum over ranks
            mpt.pathtrace(1)
        mpt.refresh()
        mpt.pathtrace(1)                         # ... and a refresh clears that state
        mpt.sync()
    owner = (np.arange(W) // band) % n
    for r, part in enumerate(parts):             # every pixel is non-zero on exactly one GPU, +0.0f elsewhere
        assert (part[owner != r] == 0).all() and not np.signbit(part[owner != r]).any()
        assert (part[owner == r][..., 3] == SPP).all()
    oc, oo = common.to_oracle(cfg, cam, objs)
    want = common.po.pathtrace(oc, oo, SPP)
    assert np.array_equal(img, want)
    assert np.isfinite(pix).all() and 0.0 <= pix.min() and pix.max() <= 1.0
    with PathTracer(cfg, objs, cam, tm) as pt:   # and the tone-mapped pixels equal the single-GPU ones
        pt.refresh()
        pt.pathtrace(SPP)
        pt.post_process()
        assert np.array_equal(pt.image_pixels.to_numpy(), pix)


def test_multi_path_tracer_reproduces_the_reference_source_golden_at_real_size():
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    g = np.load(os.path.join(common.GOLDEN_DIR, "c1_columns.npz"))
    W, H = int(g["width"]), int(g["height"])
    cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=int(g["bounces"]), seed=int(g["seed"]))
    with MultiPathTracer(cfg, objs, cam, tm, devices=range(n), band=4) as mpt:
        mpt.refresh()
        mpt.pathtrace(int(g["spp"]) if "spp" in g else 1)
        mpt.reduce(0)
        img = mpt.image_buffer.to_numpy()
    cols = g["columns"]
    assert np.array_equal(img[cols], g["image_buffer_columns"])


def test_one_process_per_gpu_under_torchrun_equals_the_oracle():
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    pytest.importorskip("torch")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(common.ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "== oracle" in r.stdout
