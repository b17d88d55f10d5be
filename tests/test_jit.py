"""Scene-specialised kernels (NVRTC): the generated nearest() must be bit-identical to the generic
one.  CPU part: the generated function is compiled for the host next to the generic code and
compared on random points; NVRTC is run on the full translation unit (no GPU needed).  GPU part:
whole images with the JIT kernel on and off."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from raytracingpbr_b200 import _native as N, scenes

PRESETS = {
    "cornell_box_shortest": (scenes.cornell_box_shortest, "Variant<FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, false>", 1.5),
    "cornell_box": (scenes.cornell_box, "Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_PLAIN, false>", 1.5),
    "cornell_box_v3": (scenes.cornell_box_v3, "Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, false>", 15.0),
    "tokyo_ibl": (scenes.tokyo_ibl, "Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, false>", 3.0),
    "bunny_glass": (scenes.bunny_glass, "Variant<FAMILY_B, 0, SHAPESET_BUNNY, MARCH_ENHANCED, false>", 1.2),
    "src_scene": (scenes.src_scene, "Variant<FAMILY_C, 0, SHAPESET_ANALYTIC, MARCH_SRC, false>", 3.0),
}

HARNESS = r'''
#include <cstring>
#include <vector>
#include "%(csrc)s/host_setup.h"
#include "%(csrc)s/rt_integrator.cuh"
namespace rt {
%(func)s
}
using namespace rt;
extern "C" __attribute__((visibility("default"))) int jit_check(const RtpbrConfig* cfg, const RtpbrObject* objs, int n, int frame,
                                                                 const float* pts, int npts, float* out_best, int* out_idx)
{
    KParams P;
    memset(&P, 0, sizeof(P));
    fill_config(P, *cfg);
    fill_objects(P, objs, n);
    fill_frame(P, frame);
    int bad = 0;
    for (int k = 0; k < npts; ++k) {
        vec3 p = V3(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]);
        int i0, i1;
        float a = nearest<%(variant)s>(P, p, i0);
        float b = jit_nearest(P, p, i1);
        out_best[k] = b; out_idx[k] = i1;
        if (memcmp(&a, &b, 4) != 0 || i0 != i1) ++bad;
#if defined(RT_JIT_SPLIT_BUNNY)
        // two-stage march of the bunny scenes: cheap part + MLP must reassemble to jit_nearest_dist()
        bool need; vec3 pb;
        float c = jit_nearest_partial(P, p, need, pb);
        if (need) c = fminf(c, fabsf(sd_bunny(pb)));
        float d = jit_nearest_dist(P, p);
        if (memcmp(&c, &d, 4) != 0 || memcmp(&a, &d, 4) != 0) ++bad;
        if (need) ++out_idx[npts];      // how many probe points needed the MLP
#endif
    }
    return bad;
}
'''


def specialised_function(cfg, objs):
    src = N.jit_source(cfg, [o.to_native() for o in objs])
    body = src[src.index("namespace rt {") + len("namespace rt {"):src.index("}  // namespace rt")]
    assert "jit_nearest" in body
    return src, body


@pytest.mark.parametrize("name", list(PRESETS))
def test_generated_nearest_is_bit_identical_on_host(name, tmp_path):
    preset, variant, extent = PRESETS[name]
    cfg, objs, cam, _ = preset(32, 32)
    src, body = specialised_function(cfg, objs)
    csrc = os.path.join(common.ROOT, "raytracingpbr_b200", "csrc")
    cu = tmp_path / "jit_check.cu"
    cu.write_text(HARNESS % dict(csrc=csrc, func=body, variant=variant))
    so = tmp_path / "libjit_check.so"
    split = ["-DRT_JIT_SPLIT_BUNNY=1"] if "jit_nearest_partial" in body else []
    assert bool(split) == (name == "bunny_glass")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                           "-fPIC,-ffp-contract=off,-fno-fast-math,-mfma,-fvisibility=hidden", "-o", str(so), str(cu)] + split,
                          stderr=subprocess.DEVNULL)
    L = C.CDLL(str(so))
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-extent, extent, (6000, 3)), rng.normal(size=(2000, 3)) * extent * 0.3]).astype(np.float32)
    pts[:8] = 0.0
    nat = [o.to_native() for o in objs]
    arr = (N.RtpbrObject * len(nat))(*nat)
    best = np.zeros(len(pts), np.float32)
    idx = np.zeros(len(pts) + 1, np.int32)
    bad = L.jit_check(C.byref(cfg), arr, len(nat), 7, pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts),
                      best.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int)))
    assert bad == 0
    assert len(set(idx[:-1].tolist())) >= min(3, len(objs))  # the probe points reach several objects
    if split:
        assert 500 < idx[-1] < len(pts) - 500                # both stages of the split march are exercised


def _check_on_host(cfg, objs, variant, extent, tmp_path, tag):
    src, body = specialised_function(cfg, objs)
    csrc = os.path.join(common.ROOT, "raytracingpbr_b200", "csrc")
    cu = tmp_path / f"jit_check_{tag}.cu"
    cu.write_text(HARNESS % dict(csrc=csrc, func=body, variant=variant))
    so = tmp_path / f"libjit_check_{tag}.so"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                           "-fPIC,-ffp-contract=off,-fno-fast-math,-mfma,-fvisibility=hidden", "-o", str(so), str(cu)],
                          stderr=subprocess.DEVNULL)
    L = C.CDLL(str(so))
    rng = np.random.default_rng(17)
    pts = np.concatenate([rng.uniform(-extent, extent, (3000, 3)), rng.normal(size=(1000, 3)) * extent * 0.3]).astype(np.float32)
    pts[:4] = 0.0
    for k, o in enumerate(objs[:16]):                      # points exactly on object centres and axes
        pts[4 + k] = np.asarray(o.transform.position, np.float32)
    nat = [o.to_native() for o in objs]
    arr = (N.RtpbrObject * len(nat))(*nat)
    best = np.zeros(len(pts), np.float32)
    idx = np.zeros(len(pts) + 1, np.int32)
    return L.jit_check(C.byref(cfg), arr, len(nat), 3, pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts),
                       best.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int))), src


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_generated_nearest_on_random_scenes(seed, tmp_path):
    """The exactness claims of jit_codegen.h (elided zero / unit matrix terms, ranged sqrt, packed pairs, doubled
    distances) on scenes nobody tuned: random primitives, rotations drawn from a mix of arbitrary angles and
    multiples of 90 degrees, zero and non-zero offsets, tiny and large extents."""
    from raytracingpbr_b200.dataclass import Material, SDFObject, Transform
    from raytracingpbr_b200.tmath import vec3
    rng = np.random.default_rng(seed)
    cfg, _, _, _ = scenes.tokyo_ibl(32, 32)               # family B, analytic shape set, enhanced marcher
    objs = []
    for _ in range(int(rng.integers(3, 13))):
        angles = [float(rng.choice([0.0, 90.0, -90.0, 180.0, 270.0, rng.uniform(-360, 360)])) for _ in range(3)]
        pos = [float(rng.choice([0.0, rng.uniform(-2, 2)])) for _ in range(3)]
        scale = [float(rng.choice([1.0, 0.5, rng.uniform(0.01, 1.5), 1e-9, 3e4])) for _ in range(3)]
        kind = int(rng.choice([scenes.SHAPE_SPHERE, scenes.SHAPE_BOX, scenes.SHAPE_BOX, scenes.SHAPE_BOX, scenes.SHAPE_CYLINDER,
                               scenes.SHAPE_CONE, scenes.SHAPE_PLANE]))
        objs.append(SDFObject(type=kind, transform=Transform(vec3(*pos), vec3(*angles), vec3(*scale)),
                              material=Material(vec3(0.5), vec3(1), 0.5, 0.0, 0.0, 1.5)))
    bad, src = _check_on_host(cfg, objs, "Variant<FAMILY_B, 0, SHAPESET_ANALYTIC, MARCH_ENHANCED, false>", 3.0, tmp_path, f"rand{seed}")
    assert bad == 0
    # family A on random boxes only (sharp boxes, first object seeds the minimum)
    cfg_a, _, _, _ = scenes.cornell_box_shortest(32, 32)
    boxes = [o for o in objs if o.type == scenes.SHAPE_BOX] or objs[:1]
    for o in boxes:
        o.type = scenes.SHAPE_BOX
    bad, _ = _check_on_host(cfg_a, boxes, "Variant<FAMILY_A, 0, SHAPESET_BOX, MARCH_PLAIN, false>", 3.0, tmp_path, f"randA{seed}")
    assert bad == 0


def test_specialised_source_drops_zero_terms():
    cfg, objs, _, _ = scenes.cornell_box_shortest(32, 32)
    src, body = specialised_function(cfg, objs)
    first = body[body.index("object 0"):body.index("object 1")]
    assert "vec3 p0 = V3(dx0, dy0, dz0);" in first and "pos.x;" in first    # identity rotation, zero offsets
    assert "P.geom" not in body                                               # no parameter-block loads in the march loop


def test_specialised_source_carries_the_right_variant_switches():
    """march constants as literals everywhere; out-of-line resolve helpers for the PBR families without the bunny;
    the two-stage march loop for bunny scenes with the enhanced marcher (jit_codegen.h)."""
    def src_of(preset):
        cfg, objs, _, _ = preset(32, 32)
        return N.jit_source(cfg, [o.to_native() for o in objs]), cfg
    a, cfg = src_of(scenes.cornell_box_shortest)
    assert "#define RT_K_MAX_STEPS %d\n" % cfg.max_steps in a and "#define RT_K_HIT_EPS" in a and "#define RT_K_T_FAR" in a
    assert "RT_RESOLVE_OOL" not in a and "RT_JIT_SPLIT_BUNNY" not in a and "jit_nearest_partial" not in a
    assert a.count("sd_box2_ranged_x2<false>") == 10                     # 4 packed pairs x (jit_nearest, jit_nearest_dist) + 1 each in jit_nearest_fast / _fast_idx
    assert "#define RT_JIT_FAST 1" in a and "#define RT_JIT_BBOX 1" in a  # five walls as planes; bounded scene
    for preset in (scenes.tokyo_ibl, scenes.cornell_box, scenes.cornell_box_v3, scenes.src_scene):
        b, _ = src_of(preset)
        assert "#define RT_RESOLVE_OOL 1" in b and "RT_JIT_SPLIT_BUNNY" not in b
        assert "RT_JIT_FAST" not in b                                    # automatic policy: fast region for family A only
        assert ("#define RT_JIT_BBOX 1" in b) == (preset is scenes.tokyo_ibl)   # t_stop where the scene has a sky (not src/)
    c, _ = src_of(scenes.bunny_glass)
    assert "#define RT_JIT_SPLIT_BUNNY 1" in c and "RT_RESOLVE_OOL" not in c
    assert "jit_nearest_partial(const KParams& P, vec3 pos, bool& need_mlp, vec3& pb)" in c and "rt_inf()" in c


# ---- the specialised MARCH (fast region = walls as planes, t_stop = provable misses), compiled for the host ----------
_JIT_HC = {}


def jit_hostcheck(name, tmp_path_factory, scene=None):
    """tests/native/hostcheck.cu compiled together with the scene-specialised translation unit of preset `name` (or of the
    explicit `scene` = (cfg, objs, cam)), with both analyses forced on wherever they apply (the automatic policy keeps
    them for the scenes where they pay)."""
    if name in _JIT_HC:
        return _JIT_HC[name]
    if scene is not None:
        cfg, objs, cam = scene
    else:
        preset = PRESETS[name][0]
        cfg, objs, cam, _ = preset(40, 32, seed=5, max_bounces=6)
    old = {k: os.environ.get(k) for k in ("RTPBR_JIT_FAST", "RTPBR_JIT_BBOX")}
    os.environ.update(RTPBR_JIT_FAST="1", RTPBR_JIT_BBOX="1")
    try:
        src = N.jit_source(cfg, [o.to_native() for o in objs])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    d = tmp_path_factory.mktemp("jit_hc_" + name)
    pre = src[:src.index('#include "pool_kernel.cuh"')]
    body = src[src.index("namespace rt {") + len("namespace rt {"):src.index("}  // namespace rt")]
    (d / "preamble.h").write_text(pre)
    (d / "body.inc").write_text(body)
    so = d / "libhostcheck_jit.so"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Wno-deprecated-gpu-targets",
                           f'-DHC_JIT_PREAMBLE="{d}/preamble.h"', f'-DHC_JIT_BODY="{d}/body.inc"', "-Xcompiler",
                           "-fPIC,-ffp-contract=off,-fno-fast-math,-mfma,-fopenmp,-fvisibility=hidden", "-o", str(so),
                           os.path.join(common.ROOT, "tests", "native", "hostcheck.cu"), "-lgomp"], stderr=subprocess.DEVNULL)
    L = C.CDLL(str(so))
    _JIT_HC[name] = (L, src, cfg, objs, cam)
    return _JIT_HC[name]


def _native_objects(objs):
    nat = [o.to_native() for o in objs]
    return (N.RtpbrObject * len(nat))(*nat), len(nat)


@pytest.mark.parametrize("name", list(PRESETS))
def test_specialised_march_ends_like_the_generic_one(name, tmp_path_factory):
    """Random rays through every preset: status and hit position of march_to_end_jit() (the pieces the pool kernel
    interleaves: t_stop from the scene bounds, full-code steps outside the fast region, the fast step inside) equal the
    generic march's, bit for bit; where the scene is bounded the specialised march takes fewer steps."""
    L, src, cfg, objs, cam = jit_hostcheck(name, tmp_path_factory)
    extent = PRESETS[name][2]
    rng = np.random.default_rng(23)
    n = 12000 if name != "bunny_glass" else 2500
    o = rng.uniform(-extent, extent, (n, 3))
    o[: n // 4] = np.asarray(cam.lookfrom, np.float64)                      # camera rays
    o[n // 4: n // 2] *= 0.3                                                 # origins well inside the scene
    d = rng.normal(size=(n, 3))
    d[: n // 4] = np.asarray(cam.lookat, np.float64) - np.asarray(cam.lookfrom, np.float64) + rng.normal(size=(n // 4, 3)) * 0.25
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[-50:, rng.integers(0, 3, 50)] = 0.0                                    # axis-parallel components
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    arr, nobj = _native_objects(objs)
    out = (C.c_int * 3)()
    bad = L.hostcheck_jit_march(C.byref(cfg), arr, nobj, 3, rays.ctypes.data_as(C.POINTER(C.c_float)), n, out)
    assert bad == 0
    hits, steps_generic, steps_jit = out[0], out[1], out[2]
    assert 0.05 * n < hits < 0.98 * n                                        # both outcomes are exercised
    if "#define RT_JIT_BBOX 1" in src:
        assert steps_jit < steps_generic                                     # provable misses end early
    else:
        assert steps_jit == steps_generic


@pytest.mark.parametrize("name", ["cornell_box_shortest", "cornell_box", "cornell_box_v3"])
def test_walls_as_planes_have_the_bits_of_the_box_distance(name, tmp_path_factory):
    L, src, cfg, objs, cam = jit_hostcheck(name, tmp_path_factory)
    assert "#define RT_JIT_FAST 1" in src
    scale = 10.0 if name == "cornell_box_v3" else 1.0
    rng = np.random.default_rng(29)
    pts = rng.uniform(-1.05, 1.05, (40000, 3))
    pts[:10000] = np.clip(pts[:10000], -0.8, 0.8)                            # inside the room ...
    k = rng.integers(0, 3, 10000)
    pts[np.arange(10000), k] = np.sign(pts[np.arange(10000), k]) * (0.8 - np.abs(rng.normal(size=10000)) * 1e-5)   # ... hugging a wall
    pts = (pts * scale).astype(np.float32)
    arr, nobj = _native_objects(objs)
    n_ok = C.c_int(0)
    bad = L.hostcheck_jit_fast(C.byref(cfg), arr, nobj, 0, pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts), C.byref(n_ok))
    assert bad == 0
    assert n_ok.value > 15000                                                # the region covers the room


@pytest.mark.parametrize("name", list(PRESETS))
def test_specialised_march_renders_the_same_image_on_host(name, tmp_path_factory):
    """Whole paths: trace_sample() with the specialised march against the generic build of the same harness."""
    L, src, cfg, objs, cam = jit_hostcheck(name, tmp_path_factory)
    if cfg.family == N.FAMILY_C:
        pytest.skip("family C: the specialised kernel changes nearest() only (covered above)")
    env = None
    if cfg.sky == N.SKY_ENVMAP:
        env = common.env_table(np.random.default_rng(1).integers(0, 256, (16, 8, 3), dtype=np.uint8), 1.4, 2.2)
    want = common.hostcheck_pathtrace(cfg, cam, objs, 3, env=env, frame=3)
    f32p = C.POINTER(C.c_float)
    L.hostcheck_pathtrace_ex.restype = C.c_int
    L.hostcheck_pathtrace_ex.argtypes = common.hostcheck().hostcheck_pathtrace_ex.argtypes
    arr, nobj = _native_objects(objs)
    got = np.zeros((cfg.width, cfg.height, 4), np.float32)
    ncam = cam.to_native()
    rc = L.hostcheck_pathtrace_ex(C.byref(cfg), C.byref(ncam), arr, nobj, got.ctypes.data_as(f32p), None,
                                  env.ctypes.data_as(f32p) if env is not None else None,
                                  env.shape[0] if env is not None else 0, env.shape[1] if env is not None else 0,
                                  3, 3, 0, 0, 1, 32, None)
    assert rc == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", [11, 12])
def test_scene_analyses_on_random_rooms(seed, tmp_path_factory):
    """The code generator's analyses on scenes nobody tuned: rooms made of axis-aligned slabs with random extents,
    offsets and quarter-turn rotations (so that the near-permutation matrices carry the reference's -4.4e-8 entries in
    every position and sign), plus freely rotated boxes inside; family A and the PBR family with rounded boxes.  The fast
    function must have the bits of the full one wherever it says ok, and whole marches must end like the generic ones."""
    from raytracingpbr_b200.dataclass import Material, SDFObject, Transform
    from raytracingpbr_b200.tmath import vec3
    rng = np.random.default_rng(seed)
    half = float(rng.uniform(0.6, 3.0))
    thick = float(rng.uniform(0.05, 0.3)) * half
    centre = rng.uniform(-0.5, 0.5, 3) * float(rng.choice([0.0, 1.0]))
    quarter = [0.0, 90.0, 180.0, 270.0, -90.0]
    objs = []
    for axis in range(3):
        for sign in (-1.0, 1.0):
            if axis == 2 and sign > 0 and seed % 2:
                continue                                                    # open front, like the Cornell box
            pos = centre.copy()
            pos[axis] += sign * half
            # the slab's thin extent must end up along `axis` after the rotation: build it in world axes, then pick a
            # rotation made of quarter turns and permute the scale accordingly (any of them is a valid description)
            scale = [half * 1.2, half * 1.2, half * 1.2]
            scale[axis] = thick
            rot = [float(rng.choice(quarter)) if rng.random() < 0.5 else 0.0 for _ in range(3)]
            m = np.zeros(9, np.float32)
            common.hostcheck().hostcheck_euler((C.c_float * 3)(*rot), m.ctypes.data_as(C.POINTER(C.c_float)))
            perm = np.abs(m.reshape(3, 3)).argmax(axis=1)                   # local axis r looks along world axis perm[r]
            local = [scale[perm[r]] for r in range(3)]
            objs.append(SDFObject(type=scenes.SHAPE_BOX, transform=Transform(vec3(*pos), vec3(*rot), vec3(*local)),
                                  material=Material(vec3(0.6), vec3(1), 1.0, 0.0, 0.0, 1.5)))
    for _ in range(int(rng.integers(1, 4))):
        objs.append(SDFObject(type=scenes.SHAPE_BOX,
                              transform=Transform(vec3(*(centre + rng.uniform(-0.5, 0.5, 3) * half)), vec3(*rng.uniform(-180, 180, 3)),
                                                  vec3(*rng.uniform(0.1, 0.35, 3) * half)),
                              material=Material(vec3(0.5), vec3(1), 1.0, 0.0, 0.0, 1.5)))
    objs.append(SDFObject(type=scenes.SHAPE_BOX, transform=Transform(vec3(*(centre + np.array([0, 0.8 * half, 0]))), vec3(0, 0, 0),
                                                                     vec3(0.2 * half, 0.01 * half, 0.2 * half)),
                          material=Material(vec3(1), vec3(50), 1.0, 0.0, 0.0, 1.0)))
    for family in ("A", "B"):
        if family == "A":
            cfg, _, cam, _ = scenes.cornell_box_shortest(40, 32, max_bounces=5, seed=seed)
        else:
            cfg, _, cam, _ = scenes.cornell_box_v2(40, 32, max_bounces=5, seed=seed)      # plain marcher, rounded boxes (0.01)
        cam.lookfrom = vec3(float(centre[0]), float(centre[1]), float(centre[2] + 3.5 * half))
        cam.lookat = vec3(float(centre[0]), float(centre[1]), float(centre[2]))
        L, src, cfg, objs_, cam = jit_hostcheck(f"room{seed}{family}", tmp_path_factory, scene=(cfg, objs, cam))
        assert "#define RT_JIT_FAST 1" in src and "#define RT_JIT_BBOX 1" in src, src[:600]
        arr, nobj = _native_objects(objs)
        pts = (centre + rng.uniform(-1.3, 1.3, (30000, 3)) * half).astype(np.float32)
        n_ok = C.c_int(0)
        bad = L.hostcheck_jit_fast(C.byref(cfg), arr, nobj, 0, pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts), C.byref(n_ok))
        assert bad == 0 and n_ok.value > 3000, (bad, n_ok.value)
        n = 6000
        o = centre + rng.uniform(-1.0, 1.0, (n, 3)) * half * 0.95
        o[: n // 3] = np.asarray(cam.lookfrom, np.float64)
        d = rng.normal(size=(n, 3))
        d[: n // 3] = np.asarray(cam.lookat, np.float64) - np.asarray(cam.lookfrom, np.float64) + rng.normal(size=(n // 3, 3)) * 0.3 * half
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([o, d], axis=1).astype(np.float32)
        out = (C.c_int * 3)()
        bad = L.hostcheck_jit_march(C.byref(cfg), arr, nobj, 0, rays.ctypes.data_as(C.POINTER(C.c_float)), n, out)
        assert bad == 0
        assert out[2] <= out[1]                                            # never more steps than the reference's march


@pytest.mark.parametrize("name", list(PRESETS))
def test_nvrtc_compiles_the_specialised_kernel(name):
    preset = PRESETS[name][0]
    cfg, objs, _, _ = preset(32, 32)
    try:
        N.jit_compile_check(cfg, [o.to_native() for o in objs])
    except N.RtpbrError as e:
        if "dlopen" in str(e):
            pytest.skip("NVRTC not available on this machine")
        raise


def test_specialised_kernel_does_not_depend_on_who_loaded_an_nvrtc_first(tmp_path):
    """PyTorch wheels bundle their own (older) libnvrtc.so.12; once torch is imported a bare-soname dlopen would pick
    that copy and the specialised kernel came out 5.5 % slower on the multi-GPU bench.  jit.cu loads the toolkit's
    NVRTC by absolute path: the cubin must be byte-identical with and without torch in the process."""
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n%s"
            "from raytracingpbr_b200 import _native as N, scenes\n"
            "cfg, objs, cam, tm = scenes.cornell_box_shortest(64, 64, max_bounces=8)\n"
            "N.jit_compile_check(cfg, [o.to_native() for o in objs])\n")
    cubins = []
    for tag, pre in (("plain", ""), ("torch", "import torch\n")):
        env = dict(os.environ, RTPBR_JIT_DUMP=str(tmp_path / tag))
        r = subprocess.run([sys.executable, "-c", code % (common.ROOT, pre)], env=env, capture_output=True, text=True)
        if r.returncode != 0 and "dlopen" in r.stderr:
            pytest.skip("NVRTC not available on this machine")
        assert r.returncode == 0, r.stderr[-2000:]
        cubins.append((tmp_path / (tag + ".cubin")).read_bytes())
    assert cubins[0] == cubins[1]


@pytest.mark.gpu
@pytest.mark.parametrize("forced", [False, True])
@pytest.mark.parametrize("name", list(PRESETS))
def test_jit_and_aot_kernels_give_the_same_bits(name, forced, monkeypatch):
    """forced: fast region and scene bounds compiled in wherever the analysis permits, not only where the automatic
    policy keeps them (jit_codegen.h) -- with the regeneration batches and the finish threshold the fast kernels use."""
    from raytracingpbr_b200 import PathTracer
    preset = PRESETS[name][0]
    if forced:
        for k, v in dict(RTPBR_JIT_FAST="1", RTPBR_JIT_BBOX="1", RTPBR_REGEN_MIN="16", RTPBR_REGEN_IDLE="8", RTPBR_FIN_MIN="4").items():
            monkeypatch.setenv(k, v)
    cfg, objs, cam, tm = preset(96, 64, seed=3, max_bounces=8)
    out = {}
    for jit in (True, False):
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.ctx.set_jit(jit)
            if cfg.sky == N.SKY_ENVMAP:
                pt.set_envmap(common.env_table(np.random.default_rng(1).integers(0, 256, (16, 8, 3), dtype=np.uint8), 1.4, 2.2))
            pt.refresh()
            pt.pathtrace(6)
            out[jit] = pt.image_buffer.to_numpy()
            active, msg = pt.ctx.jit_status()
            if jit and not active and "dlopen" in msg:
                pytest.skip("NVRTC not available on this machine: " + msg)
            assert active == jit, msg
    assert np.array_equal(out[True], out[False])


@pytest.mark.gpu
def test_device_bit_identity_along_real_rays(tmp_path):
    """Every scene evaluation of 512 x 512 x 16 spp Cornell paths: generic nearest() vs the generated
    jit_nearest() / jit_nearest_dist() ON THE DEVICE (ranged sqrt, packed f32x2 pairs, doubled distances)."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import jit_device_check
    exe = jit_device_check.build(size=512, spp=16, out_dir=str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert " 0 mismatches" in out.stdout
