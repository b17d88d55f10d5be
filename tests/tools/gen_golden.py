#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES under the scalar
Taichi stand-in in tests/tools/taichi_shim (no Taichi available).

    python tests/tools/gen_golden.py [name ...]        # needs /root/reference; run in the authoring container

The reference files are read from /root/reference and executed unmodified except for the
PARAMETER SUBSTITUTIONS listed per fixture below (resolution / bounce count / sample count --
the quantities BASELINE.json's configs vary; every substitution must match the source text
exactly once or generation aborts).  ti.random() is served by the Philox stream contract:
the n-th call a pixel makes in launch L = word n&3 of Philox4x32-10((pixel, L, n>>2, 0), (seed,
"RTPB")), implemented here in pure Python independently of oracle/ and of the CUDA code.

The fixtures are small (tens of pixels) because the stand-in interprets every fp32 operation in
Python; they are committed, and the tests only read them (nothing reads /root/reference at test
time).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RTPBR_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "taichi_shim"))

import taichi as ti  # noqa: E402  (the stand-in)
from taichi._scalar import F  # noqa: E402
from taichi.math import vec2, vec3  # noqa: E402

KEY1 = 0x52545042
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def install_rng(seed: int):
    cache = {}

    def nxt(pixel, launch, n):
        k = (pixel, launch, n >> 2)
        if k not in cache:
            cache.clear()
            cache[k] = philox4x32_10((pixel, launch, n >> 2, 0), (seed, KEY1))
        return (cache[k][n & 3] >> 8) * 2.0 ** -24
    ti.rng.next = nxt


def load_script(relpath: str, name: str, subs):
    path = os.path.join(REF, relpath)
    src = open(path).read()
    for old, new in subs:
        if src.count(old) != 1:
            raise SystemExit(f"{relpath}: substitution target {old!r} found {src.count(old)} times")
        src = src.replace(old, new)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def vlist(v):
    return np.array(v.to_list(), dtype=np.float32)


# ------------------------------------------------------------------------------ family A
def gen_shortest(name, width, height, bounces, spp, seed):
    """examples/cornell_box/cornell_box_shortest.py: whole-image buffers + function-level vectors."""
    subs = [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
            ("for i in range(3):", f"for i in range({bounces}):")]
    m = load_script("examples/cornell_box/cornell_box_shortest.py", f"ref_shortest_{name}", subs)
    install_rng(seed)
    rnd = np.random.default_rng(1234)
    out = {"width": width, "height": height, "bounces": bounces, "spp": spp, "seed": seed}

    # function-level vectors -------------------------------------------------------------
    pts = np.concatenate([rnd.uniform(-1.2, 1.2, (24, 3)), rnd.uniform(-0.9, 0.9, (8, 3)) * [1, 1, 4]]).astype(np.float32)
    sd = np.zeros((len(pts), 8), np.float32)
    for a, p in enumerate(pts):
        for k in range(8):
            sd[a, k] = m.signed_distance(m.objects[k], vec3(*p.tolist()))
    out["sd_points"], out["sd_values"] = pts, sd
    rots = np.array([[0, 0, 0], [90, 0, 0], [0, 90, 0], [0, 112, 0], [0, -197, 0], [13.5, -77.25, 211.0]], np.float32)
    out["angle_deg"] = rots
    out["angle_mat"] = np.array([m.angle(ti.math.radians(vec3(*r.tolist()))).to_list() for r in rots], np.float32)
    npts = np.array([[0.6, -0.8, 0.6, 2], [-0.8, 0.1, 0.2, 3], [0.0, 0.0, -0.8, 0], [0.1, 0.799, 0.05, 7],
                     [-0.3, 0.2, 0.05, 5], [0.3, -0.3, 0.45, 6]], np.float32)
    out["normal_in"] = npts
    out["normal_out"] = np.array([vlist(m.calc_normal(m.objects[int(q[3])], vec3(*q[:3].tolist()))) for q in npts])
    rays = []
    for _ in range(12):
        o = np.array([rnd.uniform(-0.7, 0.7), rnd.uniform(-0.7, 0.7), 3.5], np.float32)
        t = np.array([rnd.uniform(-0.9, 0.9), rnd.uniform(-0.9, 0.9), rnd.uniform(-0.9, 0.5)], np.float32)
        d = vlist(ti.math.normalize(vec3(*(t - o).tolist())))
        rec = m.raycast(m.Ray(vec3(*o.tolist()), vec3(*d.tolist()), vec3(1)))
        rays.append(np.concatenate([o, d, [float(rec.hit), rec.distance], vlist(rec.position),
                                    vlist(rec.object.material.albedo)]))
    out["raycast"] = np.array(rays, np.float32)
    hs = []
    for _ in range(8):
        n = vlist(ti.math.normalize(vec3(*rnd.normal(size=3).tolist())))
        u = rnd.random(2).astype(np.float32)
        seq = iter(u.tolist())
        saved = ti.rng.next
        ti.rng.next = lambda *a: next(seq)
        h = vlist(m.hemispheric_sampling(vec3(*n.tolist())))
        ti.rng.next = saved
        hs.append(np.concatenate([n, u, h]))
    out["hemi"] = np.array(hs, np.float32)

    # whole image: `spp` launches of kernel render() --------------------------------------
    t0 = time.time()
    bufs = []
    for s in range(spp):
        ti.rng.launch = s
        m.render(vec3(0, 0, 3.5), vec3(0, 0, -1), vec3(0, 1, 0))          # the call in main(), shortest:135
        bufs.append(m.image_buffer.to_numpy())
    out["image_buffer"] = bufs[-1]
    out["image_buffer_first"] = bufs[0]
    out["image_pixels"] = m.image_pixels.to_numpy()
    print(f"  {name}: {width}x{height}x{spp} spp in {time.time() - t0:.1f} s")
    return out


FIXTURES = {
    # name: (generator, kwargs)
    "shortest_3b": (gen_shortest, dict(width=12, height=10, bounces=3, spp=2, seed=0)),      # the file as shipped (3 bounces)
    "shortest_8b": (gen_shortest, dict(width=10, height=8, bounces=8, spp=2, seed=7)),       # BASELINE configs[1] bounce count
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    for name in (a.names or FIXTURES):
        gen, kw = FIXTURES[name]
        print(f"generating {name} ...")
        data = gen(name, **kw)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **data)


if __name__ == "__main__":
    main()
