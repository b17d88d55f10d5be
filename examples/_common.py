"""Shared driver of the headless example runners: render `spp` samples, tone-map, write a PNG
(the reference opens a GGUI window instead; that viewer is out of scope)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from raytracingpbr_b200 import MultiPathTracer, PathTracer, ibl, imwrite  # noqa: E402


def find_asset(name: str):
    """The `.hdr` files are not redistributed with this repo; look next to the reference checkout."""
    for base in (os.environ.get("RTPBR_ASSETS", ""), os.path.join(ROOT, "assets"), os.path.join(ROOT, "tests", "assets_local"),
                 "/root/reference/assets", "assets"):
        p = os.path.join(base, name)
        if base and os.path.exists(p):
            return p
    return None


def run(preset, default_res, default_spp, out_name, env=None, frame=None, **preset_kw):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=default_res[0])
    ap.add_argument("--height", type=int, default=default_res[1])
    ap.add_argument("--spp", type=int, default=default_spp)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=out_name)
    ap.add_argument("--gpus", type=int, default=1, help="render on this many GPUs of the box (one process, no torch: "
                                                         "column bands per GPU, NCCL tile reduce at tonemap time)")
    a = ap.parse_args()
    kw = dict(preset_kw)
    if a.bounces is not None:
        kw["max_bounces"] = a.bounces
    cfg, objs, cam, tm = preset(a.width, a.height, seed=a.seed, **kw)
    if a.gpus > 1 and cfg.family == 2:
        raise SystemExit("the src/ integrator keeps per-pixel ray state between launches: run it on one GPU")
    tracer = MultiPathTracer(cfg, objs, cam, tm, devices=range(a.gpus)) if a.gpus > 1 else PathTracer(cfg, objs, cam, tm)
    with tracer as pt:
        if env is not None:
            name, exposure, gamma = env
            path = find_asset(name)
            if path is None:
                print(f"[{out_name}] {name} not found (set RTPBR_ASSETS); using a grey environment")
                import numpy as np
                pt.set_envmap(ibl.process(np.full((16, 8, 3), 128, dtype=np.uint8), exposure, gamma))
            else:
                pt.set_envmap(ibl.load_envmap(path, exposure, gamma))
        if frame is not None:
            pt.set_frame(frame)
        pt.refresh()
        t0 = time.perf_counter()
        pt.pathtrace(a.spp)
        pt.post_process()
        pix = pt.image_pixels.to_numpy()
        dt = time.perf_counter() - t0
        alpha = float(pt.image_buffer.to_numpy()[..., 3].sum())
    imwrite(pix, a.out)
    print(f"[{out_name}] {a.width}x{a.height}, {a.spp} launches, {a.gpus} GPU(s), {alpha / dt / 1e6:.1f} Msamples/s -> {a.out}")
