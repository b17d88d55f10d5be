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
    exec(compile(src, path, "exec", dont_inherit=True), mod.__dict__)
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


# ------------------------------------------------------------------------------ families B / C
def _columns_worker(job):
    kind, name, width, height, bounces, seed, cols, spp = job
    want = set(cols)
    install_rng(seed)
    if kind == "shortest":
        subs = [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
                ("for i in range(3):", f"for i in range({bounces}):")]
        m = load_script("examples/cornell_box/cornell_box_shortest.py", f"ref_shortest_{name}_{cols[0]}", subs)
        run = lambda: m.render(vec3(0, 0, 3.5), vec3(0, 0, -1), vec3(0, 1, 0))          # the call in main(), shortest:135
    elif kind == "tokyo":
        env_u8 = synthetic_env()
        ti.tools.imread = lambda path: env_u8
        subs = [("image_resolution = (192*15, 108*15)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 512", f"MAX_RAYTRACE = {bounces}")]
        m = load_script("examples/scene_demo/tokyo_ibl.py", f"ref_tokyo_{name}_{cols[0]}", subs)
        m.init_scene()
        run = lambda: m.sample(vec3(0, -0.2, 4), vec3(0, -0.2, 3), vec3(0, 1, 0))
    elif kind == "cornell_box":
        subs = [("image_resolution = (1920 // 4, 1920 // 4)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 128", f"MAX_RAYTRACE = {bounces}")]
        m = load_script("examples/cornell_box/cornell_box.py", f"ref_cornell_box_{name}_{cols[0]}", subs)
        run = lambda: m.render(vec3(0, 0, 3), vec3(0, 0, -1), vec3(0, 1, 0), False)
    elif kind == "cornell_v2":
        subs = [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 3", f"MAX_RAYTRACE = {bounces}")]
        m = load_script("examples/cornell_box/cornell_box_v2.py", f"ref_cornell_v2_{name}_{cols[0]}", subs)
        run = lambda: m.render(vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), False)
    elif kind == "cornell_v3":
        d = "examples/cornell_box/cornell_box_v3"
        for n in ["config", "dataclass", "util", "scene", "sdf", "pbr", "pathtracer", "postprocessor", "renderer"]:
            sys.modules.pop(n, None)
        load_script(f"{d}/config.py", "config", [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
                                                 ("MAX_RAYTRACE = 3", f"MAX_RAYTRACE = {bounces}")])
        sys.path.insert(0, os.path.join(REF, d))
        try:
            import renderer as rmod
            import scene as m          # image_buffer lives in scene.py
        finally:
            sys.path.pop(0)
        run = lambda: rmod.render(vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), False)
    elif kind == "scene_demo":
        env_u8 = synthetic_env()
        ti.tools.imread = lambda path: env_u8
        subs = [("image_resolution = (1920 // 4, 1080 // 4)", f"image_resolution = ({width}, {height})")]
        m = load_script("examples/scene_demo/main.py", f"ref_scene_demo_{name}_{cols[0]}", subs)
        m.init_scene()
        run = lambda: m.sample(vec3(0, -0.2, 4), vec3(0, -0.2, 3), vec3(0, 1, 0))
    elif kind == "bunny":
        env_u8 = synthetic_env(seed=5)
        ti.tools.imread = lambda path: env_u8
        subs = [("image_resolution = (1920, 1080)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 512", f"MAX_RAYTRACE = {bounces}"),
                ("while True:", "while False:")]
        m = load_script("examples/bunny/bunny_sdf_glass.py", f"ref_bunny_{name}_{cols[0]}", subs)
        run = lambda: m.sample(vec3(0, 0, 4), vec3(0, 0, 3), vec3(0, 1, 0), 0)            # frame 0
    else:
        raise KeyError(kind)
    ti.pixel_filter = lambda i, j: i in want
    for launch in range(spp):          # the reference's main loop: one launch per sample, image_buffer += (colour, 1)
        ti.rng.launch = launch
        run()
    ti.pixel_filter = None
    buf = m.image_buffer.to_numpy()
    return cols, buf[cols]


def gen_columns(name, kind, width, height, bounces, seed, columns, workers=8, spp=1):
    """A BASELINE.json configuration at its REAL resolution and bounce count, launch 0 (1 spp), for a spread subset of
    image columns: the scalar stand-in needs ~0.3 s per sample, pixels are independent and the RNG stream is keyed by
    the global pixel index, so a column subset of the full image is exact."""
    import multiprocessing as mp
    t0 = time.time()
    columns = sorted(columns)
    jobs = [(kind, name, width, height, bounces, seed, columns[k::workers], spp) for k in range(workers) if columns[k::workers]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_columns_worker, jobs)
    img = np.zeros((len(columns), height, 4), np.float32)
    for cols, data in parts:
        for c, d in zip(cols, data):
            img[columns.index(c)] = d
    print(f"  {name}: {len(columns)} columns of {width}x{height} in {time.time() - t0:.1f} s")
    out = {"width": width, "height": height, "bounces": bounces, "spp": spp, "seed": seed,
           "columns": np.asarray(columns, np.int32), "image_buffer_columns": img}
    if kind == "tokyo":
        out["env_u8"] = synthetic_env()
        out["lookfrom"], out["lookat"] = np.array([0, -0.2, 4], np.float32), np.array([0, -0.2, 3], np.float32)
    if kind == "cornell_box":
        out["lookfrom"], out["lookat"] = np.array([0, 0, 3], np.float32), np.array([0, 0, -1], np.float32)
    if kind == "cornell_v3":
        out["lookfrom"], out["lookat"] = np.array([0, 0, 35], np.float32), np.array([0, 0, -10], np.float32)
    if kind == "cornell_v2":
        out["lookfrom"], out["lookat"] = np.array([0, 0, 35], np.float32), np.array([0, 0, -10], np.float32)
    if kind == "scene_demo":
        out["env_u8"] = synthetic_env()
        out["lookfrom"], out["lookat"] = np.array([0, -0.2, 4], np.float32), np.array([0, -0.2, 3], np.float32)
    if kind == "bunny":
        out["env_u8"] = synthetic_env(seed=5)
        out["frame"] = 0
        out["lookfrom"], out["lookat"] = np.array([0, 0, 4], np.float32), np.array([0, 0, 3], np.float32)
    return out


def synthetic_env(w=16, h=8, seed=3):
    """Stand-in for ti.tools.imread(<.hdr>): uint8 (W, H, 3) like stb's LDR conversion returns
    (SURVEY.md 8(c)); small and seeded so the fixture stays tiny."""
    return np.random.default_rng(seed).integers(0, 256, size=(w, h, 3), dtype=np.uint8)


def run_samples(m, fn, args, spp, out):
    bufs = []
    for s in range(spp):
        ti.rng.launch = s
        fn(*args)
        bufs.append(m.image_buffer.to_numpy())
    out["image_buffer"] = bufs[-1]
    out["image_buffer_first"] = bufs[0]


def probe_functions(m, out, nobj, pts, normal_pts, rays, raycast_fields):
    """Function-level vectors shared by the family-B scripts (signed_distance / calc_normal / raycast)."""
    sd = np.zeros((len(pts), nobj), np.float32)
    for a, p in enumerate(pts):
        for k in range(nobj):
            sd[a, k] = m.signed_distance(m.objects[k], vec3(*p.tolist()))
    out["sd_points"], out["sd_values"] = pts, sd
    out["normal_in"] = normal_pts
    out["normal_out"] = np.array([vlist(m.calc_normal(m.objects[int(q[3])], vec3(*q[:3].tolist()))) for q in normal_pts])
    rows = []
    for o, d in rays:
        rec = m.raycast(m.Ray(vec3(*o.tolist()), vec3(*d.tolist()), vec3(1)))
        rows.append(np.concatenate([o, d, raycast_fields(rec)]))
    out["raycast"] = np.array(rows, np.float32)


def random_rays(rnd, n, origin_z, spread, target_box):
    rays = []
    for _ in range(n):
        o = np.array([rnd.uniform(-spread, spread), rnd.uniform(-spread, spread), origin_z], np.float32)
        t = np.array([rnd.uniform(-1, 1) * target_box, rnd.uniform(-1, 1) * target_box, rnd.uniform(-1, 0.5) * target_box], np.float32)
        d = vlist(ti.math.normalize(vec3(*(t - o).tolist())))
        rays.append((o, d))
    return rays


def gen_cornell_box(name, width, height, bounces, spp, seed, v2=False):
    """examples/cornell_box/cornell_box.py (family B, plain marcher, PBR materials) or cornell_box_v2.py."""
    if v2:
        subs = [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 3", f"MAX_RAYTRACE = {bounces}")]
        m = load_script("examples/cornell_box/cornell_box_v2.py", f"ref_cornell_box_{name}", subs)
    else:
        subs = [("image_resolution = (1920 // 4, 1920 // 4)", f"image_resolution = ({width}, {height})"),
                ("MAX_RAYTRACE = 128", f"MAX_RAYTRACE = {bounces}")]
        m = load_script("examples/cornell_box/cornell_box.py", f"ref_cornell_box_{name}", subs)
    install_rng(seed)
    rnd = np.random.default_rng(77)
    k = 10.0 if v2 else 1.0
    out = {"width": width, "height": height, "bounces": bounces, "spp": spp, "seed": seed,
           "lookfrom": np.array([0, 0, 35 if v2 else 3], np.float32), "lookat": np.array([0, 0, -10 if v2 else -1], np.float32)}
    pts = (rnd.uniform(-1.1, 1.1, (10, 3)) * k).astype(np.float32)
    npts = np.array([[0.6, -0.8, 0.6, 2], [-0.8, 0.1, 0.2, 3], [0.1, 0.799, 0.05, 7], [0.3, -0.3, 0.45, 6]], np.float32)
    npts[:, :3] *= k
    probe_functions(m, out, 8, pts, npts, random_rays(rnd, 6, 3.0 * (35 / 3 if v2 else 1), 0.5 * k, 0.9 * k),
                    lambda rec: np.concatenate([[float(rec.hit), rec.distance], vlist(rec.position)]))
    t0 = time.time()
    cam_args = (vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), False) if v2 else (vec3(0, 0, 3), vec3(0, 0, -1), vec3(0, 1, 0), False)
    run_samples(m, m.render, cam_args, spp, out)
    out["image_pixels"] = m.image_pixels.to_numpy()
    print(f"  {name}: {width}x{height}x{spp} spp in {time.time() - t0:.1f} s")
    return out


def gen_cornell_v3(name, width, height, bounces, spp, seed):
    """examples/cornell_box/cornell_box_v3/ (family B, x10 world, rounded boxes, enhanced marcher)."""
    d = "examples/cornell_box/cornell_box_v3"
    names = ["config", "dataclass", "util", "scene", "sdf", "pbr", "pathtracer", "postprocessor", "renderer"]
    for n in names:
        sys.modules.pop(n, None)
    load_script(f"{d}/config.py", "config", [("image_resolution = (512, 512)", f"image_resolution = ({width}, {height})"),
                                             ("MAX_RAYTRACE = 3", f"MAX_RAYTRACE = {bounces}")])
    sys.path.insert(0, os.path.join(REF, d))
    try:
        import renderer as m
        import pathtracer as pt
        import scene as sc
        import sdf as sdfm
    finally:
        sys.path.pop(0)
    install_rng(seed)
    rnd = np.random.default_rng(78)
    out = {"width": width, "height": height, "bounces": bounces, "spp": spp, "seed": seed,
           "lookfrom": np.array([0, 0, 35], np.float32), "lookat": np.array([0, 0, -10], np.float32)}
    pts = (rnd.uniform(-1.1, 1.1, (8, 3)) * 10).astype(np.float32)
    sd = np.zeros((len(pts), 8), np.float32)
    for a, p in enumerate(pts):
        for k in range(8):
            sd[a, k] = sdfm.signed_distance(sc.objects[k], vec3(*p.tolist()))
    out["sd_points"], out["sd_values"] = pts, sd
    rows = []
    for o, dd in random_rays(rnd, 6, 35.0, 5.0, 9.0):
        rec = pt.raycast(pt.Ray(vec3(*o.tolist()), vec3(*dd.tolist()), vec3(1)))
        rows.append(np.concatenate([o, dd, [float(rec.hit)], vlist(rec.position)]))
    out["raycast"] = np.array(rows, np.float32)
    t0 = time.time()
    bufs = []
    for s in range(spp):
        ti.rng.launch = s
        m.render(vec3(0, 0, 35), vec3(0, 0, -10), vec3(0, 1, 0), False)
        bufs.append(sc.image_buffer.to_numpy())
    out["image_buffer"], out["image_buffer_first"] = bufs[-1], bufs[0]
    out["image_pixels"] = sc.image_pixels.to_numpy()
    for n in names:
        sys.modules.pop(n, None)
    print(f"  {name}: {width}x{height}x{spp} spp in {time.time() - t0:.1f} s")
    return out


def gen_tokyo(name, width, height, spp, seed, script="tokyo_ibl"):
    """examples/scene_demo/tokyo_ibl.py (or main.py): 7-object scene, enhanced marcher, IBL / gradient sky."""
    env_u8 = synthetic_env()
    ti.tools.imread = lambda path: env_u8
    if script == "tokyo_ibl":
        subs = [("image_resolution = (192*15, 108*15)", f"image_resolution = ({width}, {height})")]
    else:
        subs = [("image_resolution = (1920 // 4, 1080 // 4)", f"image_resolution = ({width}, {height})")]
    m = load_script(f"examples/scene_demo/{'tokyo_ibl' if script == 'tokyo_ibl' else 'main'}.py", f"ref_{script}_{name}", subs)
    m.init_scene()
    install_rng(seed)
    rnd = np.random.default_rng(79)
    out = {"width": width, "height": height, "spp": spp, "seed": seed, "env_u8": env_u8,
           "lookfrom": np.array([0, -0.2, 4], np.float32), "lookat": np.array([0, -0.2, 3], np.float32)}
    if script == "tokyo_ibl":
        out["env_table"] = m.hdr_map.img.to_numpy()
        dirs = rnd.normal(size=(12, 3))
        dirs = np.array([vlist(ti.math.normalize(vec3(*d.tolist()))) for d in dirs] + [[0, -1, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1]], np.float32)   # (0,1,0) and (-1,0,0) index out of bounds in the reference (v == 1, u == 1)
        out["sky_dirs"] = dirs
        out["sky_vals"] = np.array([vlist(m.sky_color(m.Ray(vec3(0), vec3(*d.tolist()), vec3(1)))) for d in dirs], np.float32)
    pts = (rnd.uniform(-1.5, 1.5, (10, 3)) * [1, 0.5, 2]).astype(np.float32)
    sd = np.zeros((len(pts), 7), np.float32)
    near = np.zeros((len(pts), 2), np.float32)
    for a, p in enumerate(pts):
        for k in range(7):
            sd[a, k] = m.signed_distance(m.objects[k], vec3(*p.tolist()))
        idx, dis = m.nearest_object(vec3(*p.tolist()))
        near[a] = [idx, dis]
    out["sd_points"], out["sd_values"], out["nearest"] = pts, sd, near
    npts = np.array([[0.0, 0.5, 0.0, 1], [1.0, 0.1, 0.0, 2], [-1.0, 0.1, 0.0, 4], [0.5, 0.3, -1.77, 6], [0.3, -0.501, 0.7, 0]], np.float32)
    out["normal_in"] = npts
    out["normal_out"] = np.array([vlist(m.calc_normal(m.objects[int(q[3])], vec3(*q[:3].tolist()))) for q in npts])
    rows = []
    for o, d in random_rays(rnd, 8, 4.0, 0.3, 1.2):
        obj, position, hit = m.raycast(m.Ray(vec3(*o.tolist()), vec3(*d.tolist()), vec3(1)))
        rows.append(np.concatenate([o, d, [float(hit)], vlist(position), vlist(obj.material.albedo)]))
    out["raycast"] = np.array(rows, np.float32)
    t0 = time.time()
    run_samples(m, m.sample, (vec3(0, -0.2, 4), vec3(0, -0.2, 3), vec3(0, 1, 0)), spp, out)
    m.render()
    out["image_pixels"] = m.image_pixels.to_numpy()
    print(f"  {name}: {width}x{height}x{spp} spp in {time.time() - t0:.1f} s")
    return out


def gen_bunny(name, width, height, bounces, spp, seed, frame):
    """examples/bunny/bunny_sdf_glass.py: neural SDF, glass, w = 0.5 marcher, env map ^2.2 at lookup."""
    env_u8 = synthetic_env(seed=5)
    ti.tools.imread = lambda path: env_u8
    subs = [("image_resolution = (1920, 1080)", f"image_resolution = ({width}, {height})"),
            ("MAX_RAYTRACE = 512", f"MAX_RAYTRACE = {bounces}"),
            ("while True:", "while False:")]
    m = load_script("examples/bunny/bunny_sdf_glass.py", f"ref_bunny_{name}", subs)
    install_rng(seed)
    rnd = np.random.default_rng(80)
    out = {"width": width, "height": height, "bounces": bounces, "spp": spp, "seed": seed, "frame": frame, "env_u8": env_u8,
           "lookfrom": np.array([0, 0, 4], np.float32), "lookat": np.array([0, 0, 3], np.float32)}
    pts = rnd.uniform(-0.7, 0.7, (24, 3)).astype(np.float32)
    pts[:4] *= 2.5                                       # some outside the unit sphere
    out["bunny_points"] = pts
    out["bunny_sd"] = np.array([m.sd_bunny(vec3(*p.tolist())) for p in pts], np.float32)
    m.u_frame[None] = frame
    out["sd_values"] = np.array([[m.signed_distance(m.objects[0], vec3(*p.tolist()))] for p in pts], np.float32)
    out["sd_points"] = pts
    dirs = np.array([vlist(ti.math.normalize(vec3(*d.tolist()))) for d in rnd.normal(size=(8, 3))], np.float32)
    out["sky_dirs"] = dirs
    out["sky_vals"] = np.array([vlist(m.sky_color(m.Ray(vec3(0), vec3(*d.tolist()), vec3(1)))) for d in dirs], np.float32)
    t0 = time.time()
    run_samples(m, m.sample, (vec3(0, 0, 4), vec3(0, 0, 3), vec3(0, 1, 0), frame), spp, out)
    print(f"  {name}: {width}x{height}x{spp} spp in {time.time() - t0:.1f} s")
    return out


def gen_bunny_inner(name, script, width, height, bounces, inner, seed, frame, launches, camera_z):
    """examples/bunny/bunny_sdf.py / bunny_sdf_v2.py: kernel render() with its in-kernel SAMPLE_PER_PIXEL loop (one ti.random
    stream per pixel and launch, image_buffer overwritten by every launch), white / black camera-ray background."""
    env_u8 = synthetic_env(seed=15)
    ti.tools.imread = lambda path: env_u8
    spp_line = {"bunny_sdf_v2": "SAMPLE_PER_PIXEL = 12", "bunny_sdf": "SAMPLE_PER_PIXEL = 4"}[script]
    subs = [("image_resolution = (3840, 2160)", f"image_resolution = ({width}, {height})"),
            ("MAX_RAYTRACE = 128", f"MAX_RAYTRACE = {bounces}"),
            (spp_line, f"SAMPLE_PER_PIXEL = {inner}"),
            ("while window.running:", "while False:")]
    m = load_script(f"examples/bunny/{script}.py", f"ref_{script}_{name}", subs)
    install_rng(seed)
    out = {"width": width, "height": height, "bounces": bounces, "inner_spp": inner, "launches": launches, "seed": seed, "frame": frame,
           "env_u8": env_u8, "lookfrom": np.array([0, 0, camera_z], np.float32), "lookat": np.array([0, 0, camera_z - 1], np.float32)}
    rnd = np.random.default_rng(81)
    pts = rnd.uniform(-0.7, 0.7, (16, 3)).astype(np.float32)
    pts[:3] *= 2.5
    m.u_frame[None] = frame
    out["sd_points"] = pts
    out["sd_values"] = np.array([[m.signed_distance(m.objects[0], vec3(*p.tolist()))] for p in pts], np.float32)
    t0 = time.time()
    bufs = []
    for L in range(launches):
        ti.rng.launch = L
        m.render(vec3(0, 0, camera_z), vec3(0, 0, camera_z - 1), vec3(0, 1, 0), False, frame)
        bufs.append(m.image_buffer.to_numpy())
    out["image_buffer"] = bufs[-1]
    out["image_buffer_first"] = bufs[0]
    out["image_pixels"] = m.image_pixels.to_numpy()
    print(f"  {name}: {width}x{height} x {inner} samples per launch x {launches} launches in {time.time() - t0:.1f} s")
    return out


class _SrcFinder:
    """Imports `src.*` from the reference tree, applying parameter substitutions to src/config.py."""
    def __init__(self, subs):
        self.subs = subs

    def find_spec(self, fullname, path=None, target=None):
        import importlib.util
        if fullname != "src" and not fullname.startswith("src."):
            return None
        rel = fullname.replace(".", "/")
        base = os.path.join(REF, rel)
        if os.path.isdir(base):
            spec = importlib.util.spec_from_loader(fullname, self, is_package=True)
            spec.submodule_search_locations = [base]
            spec.origin = os.path.join(base, "__init__.py")
            return spec
        spec = importlib.util.spec_from_loader(fullname, self)
        spec.origin = base + ".py"
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        path = module.__spec__.origin
        if not os.path.exists(path):
            return                                          # namespace-like package without __init__.py
        src = open(path).read()
        for old, new in self.subs.get(module.__name__, []):
            if src.count(old) != 1:
                raise SystemExit(f"{path}: substitution target {old!r} found {src.count(old)} times")
            src = src.replace(old, new)
        module.__file__ = path
        exec(compile(src, path, "exec", dont_inherit=True), module.__dict__)


def gen_src(name, width, height, launches, seed, spp_per_launch=1, adaptive=False, noise_threshold=None):
    """src/ package (family C): kernel pathtrace() x launches, ray_buffer state persisted between launches."""
    env_u8 = synthetic_env(seed=9)
    ti.tools.imread = lambda path: env_u8
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    subs = {"src.config": [("image_resolution = (1920 * 4 // 10, 1080 * 4 // 10)", f"image_resolution = ({width}, {height})"),
                           ("SAMPLES_PER_PIXEL = 1  #", f"SAMPLES_PER_PIXEL = {spp_per_launch}  #")]}
    if adaptive:
        subs["src.config"] += [("ADAPTIVE_SAMPLING = False", "ADAPTIVE_SAMPLING = True"),
                               ("NOISE_THRESHOLD = 1e-4", f"NOISE_THRESHOLD = {noise_threshold}")]
    finder = _SrcFinder(subs)
    sys.meta_path.insert(0, finder)
    try:
        import src.renderer as rend
        import src.pathtracer as pt
        import src.scene as sc
        import src.camera as cam
        import src.fileds as fld
        import src.ibl as ibl
        import src.pbr  # noqa: F401
    finally:
        sys.meta_path.remove(finder)
    cam.smooth.position[None] = vec3(0, -0.2, 4.0)             # src/main.py:17 camera.position(0, -0.2, 4.0)
    cam.smooth.lookat[None] = vec3(0, -0.2, 3.0)
    cam.smooth.up[None] = vec3(0, 1, 0)
    sc.build_scene()
    install_rng(seed)
    rnd = np.random.default_rng(81)
    out = {"width": width, "height": height, "launches": launches, "seed": seed, "env_u8": env_u8, "spp_per_launch": spp_per_launch,
           "lookfrom": np.array([0, -0.2, 4], np.float32), "lookat": np.array([0, -0.2, 3], np.float32),
           "env_table": ibl.hdr_map.img.to_numpy()}
    pts = (rnd.uniform(-1.5, 1.5, (10, 3)) * [1, 0.5, 2]).astype(np.float32)
    near = np.zeros((len(pts), 2), np.float32)
    for a, p in enumerate(pts):
        idx, dis = sc.nearest(vec3(*p.tolist()))
        near[a] = [idx, dis]
    out["sd_points"], out["nearest"] = pts, near
    npts = np.array([[0.0, 0.5, 0.0, 1], [1.0, 0.1, 0.0, 2], [-1.0, 0.1, 0.0, 4], [0.5, 0.3, -1.77, 6], [0.3, -0.501, 0.7, 0]], np.float32)
    out["normal_in"] = npts
    out["normal_out"] = np.array([vlist(sc.calc_normal(sc.objects[int(q[3])], vec3(*q[:3].tolist()))) for q in npts])
    rows = []
    for o, d in random_rays(rnd, 8, 4.0, 0.3, 1.2):
        ray, obj, hit = sc.raycast(pt.Ray(vec3(*o.tolist()), vec3(*d.tolist()), vec3(1), 0))
        rows.append(np.concatenate([o, d, [float(hit)], vlist(ray.origin), vlist(obj.material.albedo)]))
    out["raycast"] = np.array(rows, np.float32)
    t0 = time.time()
    rend.refresh()
    snaps = {}
    sampled = []
    for L in range(launches):
        ti.rng.launch = L
        if adaptive:                                   # src/renderer.py:29-32: pathtrace(); post_process()
            sampled.append(int((fld.diff_pixels.to_numpy() > noise_threshold).sum()))
        pt.pathtrace()
        if adaptive:
            import src.postprocessor as pp
            pp.post_process()
        if L in (0, launches // 2):
            snaps[L] = fld.image_buffer.to_numpy()
    if adaptive:
        out["sampled_per_launch"] = np.array(sampled, np.int32)
        out["noise_threshold"] = np.float32(noise_threshold)
        out["diff_pixels"] = fld.diff_pixels.to_numpy()
        out["diff_buffer"] = fld.diff_buffer.to_numpy()
    out["image_buffer"] = fld.image_buffer.to_numpy()
    out["image_buffer_first"] = snaps[0]
    rb = np.zeros((width, height, 10), np.float32)
    rb[..., 0:3] = fld.ray_buffer.member("origin")
    rb[..., 3:6] = fld.ray_buffer.member("direction")
    rb[..., 6:9] = fld.ray_buffer.member("color")
    rb[..., 9] = fld.ray_buffer.member("depth").astype(np.int32).view(np.float32)
    out["ray_buffer"] = rb
    if not adaptive:
        rend.post_process()
    out["image_pixels"] = fld.image_pixels.to_numpy()
    print(f"  {name}: {width}x{height} x {launches} launches in {time.time() - t0:.1f} s")
    return out


def _src_columns_worker(job):
    name, width, height, launches, seed, cols = job
    env_u8 = synthetic_env(seed=9)
    ti.tools.imread = lambda path: env_u8
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    subs = {"src.config": [("image_resolution = (1920 * 4 // 10, 1080 * 4 // 10)", f"image_resolution = ({width}, {height})")]}
    finder = _SrcFinder(subs)
    sys.meta_path.insert(0, finder)
    try:
        import src.renderer as rend
        import src.pathtracer as pt
        import src.scene as sc
        import src.camera as cam
        import src.fileds as fld
        import src.pbr  # noqa: F401
    finally:
        sys.meta_path.remove(finder)
    cam.smooth.position[None] = vec3(0, -0.2, 4.0)             # src/main.py:17 camera.position(0, -0.2, 4.0)
    cam.smooth.lookat[None] = vec3(0, -0.2, 3.0)
    cam.smooth.up[None] = vec3(0, 1, 0)
    sc.build_scene()
    install_rng(seed)
    want = set(cols)
    ti.pixel_filter = lambda i, j: i in want
    rend.refresh()
    for L in range(launches):
        ti.rng.launch = L
        pt.pathtrace()
    ti.pixel_filter = None
    rb = np.zeros((width, height, 10), np.float32)
    rb[..., 0:3] = fld.ray_buffer.member("origin")
    rb[..., 3:6] = fld.ray_buffer.member("direction")
    rb[..., 6:9] = fld.ray_buffer.member("color")
    rb[..., 9] = fld.ray_buffer.member("depth").astype(np.int32).view(np.float32)
    return cols, fld.image_buffer.to_numpy()[cols], rb[cols]


def gen_src_columns(name, width, height, launches, seed, columns, workers=8):
    """The src/ package (family C) at the resolution src/config.py ships (768 x 432): `launches` launches of kernel
    pathtrace() (one bounce per pixel per launch, state carried in ray_buffer) for a spread subset of image columns."""
    import multiprocessing as mp
    t0 = time.time()
    columns = sorted(columns)
    jobs = [(name, width, height, launches, seed, columns[k::workers]) for k in range(workers) if columns[k::workers]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_src_columns_worker, jobs)
    img = np.zeros((len(columns), height, 4), np.float32)
    rb = np.zeros((len(columns), height, 10), np.float32)
    for cols, a, b in parts:
        for c, x, y in zip(cols, a, b):
            img[columns.index(c)], rb[columns.index(c)] = x, y
    print(f"  {name}: {len(columns)} columns of {width}x{height} x {launches} launches in {time.time() - t0:.1f} s")
    return {"width": width, "height": height, "launches": launches, "seed": seed, "env_u8": synthetic_env(seed=9), "spp_per_launch": 1,
            "lookfrom": np.array([0, -0.2, 4], np.float32), "lookat": np.array([0, -0.2, 3], np.float32),
            "columns": np.asarray(columns, np.int32), "image_buffer_columns": img, "ray_buffer_columns": rb}


FIXTURES = {
    # name: (generator, kwargs)
    "shortest_3b": (gen_shortest, dict(width=12, height=10, bounces=3, spp=2, seed=0)),      # the file as shipped (3 bounces)
    "shortest_8b": (gen_shortest, dict(width=10, height=8, bounces=8, spp=2, seed=7)),       # BASELINE configs[1] bounce count
    # BASELINE.json configs[0] at its real size: 32 spread columns of the 256 x 256 x 1 spp x 4 bounce image
    "c0_columns": (gen_columns, dict(kind="shortest", width=256, height=256, bounces=4, seed=0,
                                     columns=[0, 1, 7, 15, 31, 40, 63, 64, 77, 90, 100, 111, 120, 127, 128, 129, 140, 150, 160, 170,
                                              180, 191, 192, 200, 210, 220, 230, 240, 250, 253, 254, 255])),
    # configs[1] / configs[4]: the same scene at 1024^2 and 4096^2 with 8 bounces; configs[3]: tokyo_ibl at 1920 x 1080, 8 bounces
    "c1_columns": (gen_columns, dict(kind="shortest", width=1024, height=1024, bounces=8, seed=0,
                                     columns=[0, 255, 400, 511, 512, 640, 900, 1023])),
    "c1_columns_4spp": (gen_columns, dict(kind="shortest", width=1024, height=1024, bounces=8, seed=0, spp=4,
                                          columns=[64, 200, 330, 470, 555, 690, 820, 960])),
    "c4_columns": (gen_columns, dict(kind="shortest", width=4096, height=4096, bounces=8, seed=0, columns=[1500, 2047])),
    # configs[2]: bunny_sdf_glass.py at 1024 x 1024, 16 bounces, frame 0 (the neural SDF costs seconds per sample here)
    "c2_columns": (gen_columns, dict(kind="bunny", width=1024, height=1024, bounces=16, seed=0,
                                     columns=[300, 420, 480, 511, 512, 560, 640, 760])),
    "c3_columns": (gen_columns, dict(kind="tokyo", width=1920, height=1080, bounces=8, seed=0,
                                     columns=[0, 480, 800, 959, 960, 1100, 1500, 1919])),
    # the other example scripts exactly as shipped (resolution and bounce cap of the files)
    "cornell_box_columns": (gen_columns, dict(kind="cornell_box", width=480, height=480, bounces=128, seed=0, columns=[0, 120, 239, 240, 300, 479])),
    "cornell_v2_columns": (gen_columns, dict(kind="cornell_v2", width=512, height=512, bounces=3, seed=0, columns=[0, 100, 255, 256, 300, 400, 450, 511])),
    "cornell_v3_columns": (gen_columns, dict(kind="cornell_v3", width=512, height=512, bounces=3, seed=0, columns=[0, 100, 255, 256, 300, 400, 450, 511])),
    "scene_demo_columns": (gen_columns, dict(kind="scene_demo", width=480, height=270, bounces=128, seed=0, columns=[0, 100, 200, 239, 240, 300, 400, 479])),
    # the src/ package at its shipped resolution, 16 launches of pathtrace()
    "src_columns": (gen_src_columns, dict(width=768, height=432, launches=16, seed=0, columns=[0, 100, 250, 383, 384, 500, 640, 767])),
    "cornell_box": (gen_cornell_box, dict(width=8, height=8, bounces=6, spp=2, seed=1)),
    "cornell_v2": (gen_cornell_box, dict(width=8, height=8, bounces=3, spp=2, seed=9, v2=True)),
    "cornell_v3": (gen_cornell_v3, dict(width=8, height=8, bounces=3, spp=2, seed=2)),
    "tokyo_ibl": (gen_tokyo, dict(width=12, height=8, spp=3, seed=3)),
    "scene_demo": (gen_tokyo, dict(width=8, height=6, spp=2, seed=4, script="main")),
    "bunny_glass": (gen_bunny, dict(width=8, height=6, bounces=16, spp=1, seed=5, frame=7)),
    # the two other bunny scripts: in-kernel sample loop on one RNG stream, white / black background for camera rays
    "bunny_sdf_v2": (gen_bunny_inner, dict(script="bunny_sdf_v2", width=8, height=6, bounces=12, inner=3, seed=12, frame=5, launches=2, camera_z=4)),
    "bunny_sdf": (gen_bunny_inner, dict(script="bunny_sdf", width=8, height=6, bounces=12, inner=2, seed=13, frame=9, launches=2, camera_z=5)),
    "src_scene": (gen_src, dict(width=10, height=6, launches=12, seed=6)),
    "src_adaptive": (gen_src, dict(width=8, height=6, launches=14, seed=8, adaptive=True, noise_threshold=0.2)),
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
