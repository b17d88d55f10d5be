"""Golden-vector tests: the oracle (CPU, always) and the CUDA path (-m gpu) against fixtures
produced by executing the reference's own source files under tests/tools/taichi_shim
(tests/tools/gen_golden.py).  Everything is compared BIT FOR BIT."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import common
from common import po

GOLDEN = os.path.join(common.ROOT, "tests", "golden")
SHORTEST = sorted(glob.glob(os.path.join(GOLDEN, "shortest_*.npz")))


def f32p(a):
    return np.ascontiguousarray(a, dtype=np.float32).ctypes.data_as(C.POINTER(C.c_float))


def params(g):
    return tuple(int(g[k]) for k in ("width", "height", "bounces", "spp", "seed"))


def test_fixtures_present():
    assert len(SHORTEST) >= 2


@pytest.mark.parametrize("path", SHORTEST, ids=os.path.basename)
def test_oracle_image_buffer_matches_reference_source(path):
    g = np.load(path)
    W, H, B, S, seed = params(g)
    cfg = po.cornell_shortest_config(W, H, B, seed)
    objs = po.objects_array(po.cornell_shortest_objects())
    assert np.array_equal(po.pathtrace(cfg, objs, 1), g["image_buffer_first"])
    assert np.array_equal(po.pathtrace(cfg, objs, S), g["image_buffer"])
    # as-written mode (matrices recomputed per evaluation, shortest:43) gives the same bits
    assert np.array_equal(po.pathtrace(cfg, objs, S, hoisted=False), g["image_buffer"])


@pytest.mark.parametrize("path", SHORTEST[:1], ids=os.path.basename)
def test_oracle_functions_match_reference_source(path):
    g = np.load(path)
    W, H, B, S, seed = params(g)
    cfg = po.cornell_shortest_config(W, H, B, seed)
    objs = po.objects_array(po.cornell_shortest_objects())
    L = po.lib()
    for p, row in zip(g["sd_points"], g["sd_values"]):            # signed_distance, shortest:41-45
        for k in range(8):
            for hoisted in (0, 1):
                assert np.float32(L.orc_signed_distance(C.byref(cfg), objs, 8, k, f32p(p), hoisted)) == row[k]
    out = np.zeros(9, np.float32)
    for r, m in zip(g["angle_deg"], g["angle_mat"]):              # angle(radians(.)), shortest:34-39
        L.orc_angle_deg(f32p(r), f32p(out))
        assert np.array_equal(out.reshape(3, 3), m)
    n = np.zeros(3, np.float32)
    for q, want in zip(g["normal_in"], g["normal_out"]):          # calc_normal, shortest:55-61
        L.orc_calc_normal(C.byref(cfg), objs, 8, int(q[3]), f32p(q[:3]), f32p(n))
        assert np.array_equal(n, want)
    rec = np.zeros(7, np.float32)
    albedo = np.array([list(o.albedo) for o in objs], np.float32)
    for row in g["raycast"]:                                      # raycast, shortest:63-72
        L.orc_raycast(C.byref(cfg), objs, 8, f32p(row[0:3]), f32p(row[3:6]), f32p(rec))
        assert rec[0] == row[6] and rec[3] == row[7]
        assert np.array_equal(rec[4:7], row[8:11])
        if row[6]:
            assert np.array_equal(albedo[int(rec[1])], row[11:14])
    h = np.zeros(3, np.float32)
    for row in g["hemi"]:                                         # hemispheric_sampling, shortest:74-79
        L.orc_hemispheric_sampling(f32p(row[0:3]), float(row[3]), float(row[4]), f32p(h))
        assert np.array_equal(h, row[5:8])


@pytest.mark.parametrize("path", SHORTEST, ids=os.path.basename)
def test_product_host_code_matches_reference_source(path):
    # the product's __host__ __device__ integrator compiled for the CPU (tests/native/hostcheck.cu)
    from raytracingpbr_b200 import scenes
    g = np.load(path)
    W, H, B, S, seed = params(g)
    cfg, objs, cam, _ = scenes.cornell_box_shortest(W, H, max_bounces=B, seed=seed)
    assert np.array_equal(common.hostcheck_pathtrace(cfg, cam, objs, S), g["image_buffer"])


# ---------------------------------------------------------------- BASELINE.json configurations at their real size
# Spread columns of launch 0 (1 spp; c1_columns_4spp: launches 0-3) of the full-resolution images, rendered by the reference's own source files under
# the stand-in (tests/tools/gen_golden.py *_columns): configs[0] 256^2 / 4 bounces, configs[1] 1024^2 / 8 bounces,
# configs[4] 4096^2 / 8 bounces (cornell_box_shortest.py), configs[3] 1920 x 1080 / 8 bounces (tokyo_ibl.py) and
# configs[2] 1024^2 / 16 bounces, frame 0 (bunny_sdf_glass.py).
REAL_SIZE = {"c0_columns": (256, 256, 4), "c1_columns": (1024, 1024, 8), "c1_columns_4spp": (1024, 1024, 8), "c4_columns": (4096, 4096, 8), "c3_columns": (1920, 1080, 8),
             "c2_columns": (1024, 1024, 16),
             # the remaining example scripts exactly as shipped (resolution and bounce cap of the files)
             "cornell_box_columns": (480, 480, 128), "cornell_v2_columns": (512, 512, 3), "cornell_v3_columns": (512, 512, 3),
             "scene_demo_columns": (480, 270, 128)}
REAL_SIZE_PRESENT = [n for n in REAL_SIZE if os.path.exists(os.path.join(GOLDEN, n + ".npz"))]


def real_size_case(name):
    from raytracingpbr_b200 import scenes
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, B, S, seed = params(g)
    assert (W, H, B) == REAL_SIZE[name]                         # exactly the configuration BASELINE.json names
    cols, want = g["columns"].astype(int), g["image_buffer_columns"]
    assert want.shape == (len(cols), H, 4) and (want[..., 3] == float(S)).all()
    assert len(np.unique(want[..., :3])) > 10 and (want[..., :3].sum(-1) > 0).mean() > 0.05     # a real image, not zeros
    env = None
    if name == "c3_columns":
        cfg, objs, cam, tm = scenes.tokyo_ibl(W, H, max_bounces=B, seed=seed)
        cam.lookfrom, cam.lookat = g["lookfrom"], g["lookat"]
        env = common.env_table(g["env_u8"], 1.8, 2.2)               # tokyo_ibl.py:60
    elif name in ("cornell_box_columns", "cornell_v2_columns", "cornell_v3_columns", "scene_demo_columns"):
        preset = {"cornell_box_columns": scenes.cornell_box, "cornell_v2_columns": scenes.cornell_box_v2,
                  "cornell_v3_columns": scenes.cornell_box_v3, "scene_demo_columns": scenes.scene_demo}[name]
        cfg, objs, cam, tm = preset(W, H, max_bounces=B, seed=seed)
        cam.lookfrom, cam.lookat = g["lookfrom"], g["lookat"]
    elif name == "c2_columns":
        cfg, objs, cam, tm = scenes.bunny_glass(W, H, max_bounces=B, seed=seed, frame=int(g["frame"]))
        cam.lookfrom, cam.lookat = g["lookfrom"], g["lookat"]
        env = common.env_table(g["env_u8"], 1.8, 2.2)               # bunny_sdf_glass.py:279-280
    else:
        cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=B, seed=seed)
    return cfg, objs, cam, tm, env, cols, want, S


def test_configs0_fixture_present():
    assert "c0_columns" in REAL_SIZE_PRESENT


@pytest.mark.parametrize("name", REAL_SIZE_PRESENT)
def test_oracle_matches_reference_source_at_real_size(name):
    cfg, objs, cam, tm, env, cols, want, spp = real_size_case(name)
    oc, oo = common.to_oracle(cfg, cam, objs)
    buf = np.zeros((cfg.width, cfg.height, 4), np.float32)
    for c0 in cols:                                             # the oracle renders single columns of the full image
        po.pathtrace(oc, oo, spp, env=env, i0=int(c0), i1=int(c0) + 1, image_buffer=buf)
    assert np.array_equal(buf[cols], want), name


@pytest.mark.parametrize("name", [n for n in REAL_SIZE_PRESENT if n not in ("c4_columns", "c2_columns")])   # (whole images: keep the CPU suite short)
def test_product_host_code_matches_reference_source_at_real_size(name):
    cfg, objs, cam, tm, env, cols, want, spp = real_size_case(name)
    assert np.array_equal(common.hostcheck_pathtrace(cfg, cam, objs, spp, env=env)[cols], want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", REAL_SIZE_PRESENT)
def test_cuda_matches_reference_source_at_real_size(name):
    from raytracingpbr_b200 import PathTracer, _native as N
    cfg, objs, cam, tm, env, cols, want, spp = real_size_case(name)
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg.kernel = kernel
        with PathTracer(cfg, objs, cam, tm) as pt:
            if env is not None:
                pt.set_envmap(env)
            pt.refresh()
            pt.pathtrace(spp)                                   # all the reference's launches in ONE kernel launch
            buf = pt.image_buffer.to_numpy()
        assert np.array_equal(buf[cols], want), (name, kernel)


SRC_COLUMNS = os.path.join(GOLDEN, "src_columns.npz")


def src_columns_case():
    """The src/ package (family C) at its shipped 768 x 432, 16 launches of pathtrace(), 8 spread columns."""
    from raytracingpbr_b200 import scenes
    g = np.load(SRC_COLUMNS)
    W, H, seed, launches = int(g["width"]), int(g["height"]), int(g["seed"]), int(g["launches"])
    assert (W, H) == (768, 432)                                     # src/config.py:7
    cfg, objs, cam, tm = scenes.src_scene(W, H, seed=seed)
    cam.lookfrom, cam.lookat = g["lookfrom"], g["lookat"]
    env = common.env_table(g["env_u8"], 1.4, 2.2)                   # src/ibl.py:33
    cols = g["columns"].astype(int)
    assert g["image_buffer_columns"][..., 3].sum() > 0 and (g["ray_buffer_columns"][..., 9].view(np.int32) < 0).any()
    return cfg, objs, cam, tm, env, launches, cols, g["image_buffer_columns"], g["ray_buffer_columns"]


@pytest.mark.skipif(not os.path.exists(SRC_COLUMNS), reason="fixture not generated")
def test_oracle_family_c_matches_reference_source_at_real_size():
    cfg, objs, cam, tm, env, launches, cols, want_img, want_rb = src_columns_case()
    oc, oo = common.to_oracle(cfg, cam, objs)
    img = np.zeros((cfg.width, cfg.height, 4), np.float32)
    rb = np.zeros((cfg.width, cfg.height, 10), np.float32)
    for c0 in cols:
        po.pathtrace(oc, oo, launches, env=env, ray_buffer=rb, image_buffer=img, i0=int(c0), i1=int(c0) + 1)
    assert np.array_equal(img[cols], want_img)
    assert np.array_equal(rb[cols].view(np.int32), want_rb.view(np.int32))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SRC_COLUMNS), reason="fixture not generated")
def test_cuda_family_c_matches_reference_source_at_real_size():
    from raytracingpbr_b200 import PathTracer, _native as N
    cfg, objs, cam, tm, env, launches, cols, want_img, want_rb = src_columns_case()
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg.kernel = kernel
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.set_envmap(env)
            pt.refresh()
            pt.pathtrace(5)                                         # 16 reference launches in two kernel launches
            pt.pathtrace(launches - 5)
            img = pt.image_buffer.to_numpy()
            rb = pt.ray_buffer.to_numpy()
        assert np.array_equal(img[cols], want_img), kernel
        assert np.array_equal(rb[cols].view(np.int32), want_rb.view(np.int32)), kernel


FAMILY_B = ["cornell_box", "cornell_v2", "cornell_v3", "tokyo_ibl", "scene_demo", "bunny_glass"]


def oracle_of(name):
    g, cfg, objs, cam, tm, env = common.golden_case(name)
    oc, oo = common.to_oracle(cfg, cam, objs)
    if "frame" in g:
        oc.frame = int(g["frame"])
    return g, oc, oo, env


@pytest.mark.parametrize("name", FAMILY_B)
def test_oracle_family_b_image_buffer_matches_reference_source(name):
    g, oc, oo, env = oracle_of(name)
    assert np.array_equal(po.pathtrace(oc, oo, 1, env=env), g["image_buffer_first"])
    assert np.array_equal(po.pathtrace(oc, oo, int(g["spp"]), env=env), g["image_buffer"])


@pytest.mark.parametrize("name", FAMILY_B + ["src_scene"])
def test_oracle_family_bc_functions_match_reference_source(name):
    g, oc, oo, env = oracle_of(name)
    L, n = po.lib(), len(oo)
    if "sd_values" in g:                                       # signed_distance of every object
        for p, row in zip(g["sd_points"], g["sd_values"]):
            for k in range(len(row)):
                assert np.float32(L.orc_signed_distance_g(C.byref(oc), oo, n, k, f32p(p))) == row[k]
    if "nearest" in g:                                         # nearest / nearest_object
        d = C.c_float()
        for p, row in zip(g["sd_points"], g["nearest"]):
            assert L.orc_nearest_g(C.byref(oc), oo, n, f32p(p), C.byref(d)) == int(row[0])
            assert np.float32(d.value) == row[1]
    if "normal_out" in g:                                      # calc_normal
        out = np.zeros(3, np.float32)
        for q, want in zip(g["normal_in"], g["normal_out"]):
            L.orc_calc_normal_g(C.byref(oc), oo, n, int(q[3]), f32p(q[:3]), f32p(out))
            assert np.array_equal(out, want)
    if "raycast" in g:                                         # raycast: hit flag + final position
        rec = np.zeros(7, np.float32)
        for row in g["raycast"]:
            L.orc_raycast_g(C.byref(oc), oo, n, f32p(row[0:3]), f32p(row[3:6]), f32p(rec))
            pos = row[8:11] if name in ("cornell_box", "cornell_v2") else row[7:10]
            assert rec[0] == row[6] and np.array_equal(rec[4:7], pos)
    if "bunny_sd" in g:                                        # sd_bunny, bunny_sdf_glass.py:149-203
        for p, want in zip(g["bunny_points"], g["bunny_sd"]):
            assert np.float32(L.orc_sd_bunny(f32p(p))) == want
    if "sky_vals" in g:                                        # sky_color -> sample_spherical_map -> texture
        out = np.zeros(3, np.float32)
        for d_, want in zip(g["sky_dirs"], g["sky_vals"]):
            L.orc_sky_envmap(f32p(env), env.shape[0], env.shape[1], f32p(d_), f32p(out))
            assert np.array_equal(out, want)
    if "env_table" in g:                                       # Image.process
        assert np.array_equal(env, g["env_table"])


def test_oracle_family_c_matches_reference_source():
    # src/: kernel pathtrace() x launches with ray_buffer state carried between launches
    g, oc, oo, env = oracle_of("src_scene")
    W, H = int(g["width"]), int(g["height"])
    rb = np.zeros((W, H, 10), np.float32)
    img = po.pathtrace(oc, oo, 1, env=env, ray_buffer=rb)
    assert np.array_equal(img, g["image_buffer_first"])
    img = po.pathtrace(oc, oo, int(g["launches"]) - 1, sample_base=1, env=env, ray_buffer=rb, image_buffer=img)
    assert np.array_equal(img, g["image_buffer"])
    assert np.array_equal(rb.view(np.int32), g["ray_buffer"].view(np.int32))
    assert img[..., 3].sum() > 0 and (g["ray_buffer"][..., 9].view(np.int32) < 0).any()


@pytest.mark.parametrize("name", FAMILY_B + ["src_scene"])
def test_product_host_code_families_bc_match_reference_source(name):
    # the product's __host__ __device__ integrator (rt_integrator.cuh) compiled for the CPU
    g, cfg, objs, cam, tm, env = common.golden_case(name)
    frame = int(g["frame"]) if "frame" in g else 0
    if name == "src_scene":
        rb = np.zeros((cfg.width, cfg.height, 10), np.float32)
        img = common.hostcheck_pathtrace(cfg, cam, objs, 1, ray_buffer=rb, env=env)
        assert np.array_equal(img, g["image_buffer_first"])
        img = common.hostcheck_pathtrace(cfg, cam, objs, int(g["launches"]) - 1, sample_base=1, image=img, ray_buffer=rb, env=env)
        assert np.array_equal(img, g["image_buffer"])
        assert np.array_equal(rb.view(np.int32), g["ray_buffer"].view(np.int32))
    else:
        img = common.hostcheck_pathtrace(cfg, cam, objs, int(g["spp"]), env=env, frame=frame)
        assert np.array_equal(img, g["image_buffer"])


def test_product_bunny_sdf_matches_reference_source():
    g = np.load(os.path.join(GOLDEN, "bunny_glass.npz"))
    H = common.hostcheck()
    for p, want in zip(g["bunny_points"], g["bunny_sd"]):
        assert np.float32(H.hostcheck_sd_bunny(f32p(p))) == want


def _adaptive_run(pathtrace_fn, g, cfg):
    """refresh(); then launches x (pathtrace(); post_process()) like src/renderer.py:25-32, with the oracle's
    restatement of post_process(); `pathtrace_fn(L, image, ray_buffer, diff_pixels)` renders one launch."""
    W, H = cfg.width, cfg.height
    img = np.zeros((W, H, 4), np.float32)
    rb = np.zeros((W, H, 10), np.float32)
    pix = np.zeros((W, H, 3), np.float32)
    dbuf = np.ones((W, H, 2), np.float32)                     # refresh(): diff_buffer = vec2(1), diff_pixels = 1e32
    dpix = np.full((W, H), 1e32, np.float32)
    sampled = []
    for L in range(int(g["launches"])):
        sampled.append(int((dpix > np.float32(g["noise_threshold"])).sum()))
        pathtrace_fn(L, img, rb, dpix)
        po.post_process_src(img, pix, dbuf, dpix, 1.0, 2.2, True)
    return img, rb, pix, dbuf, dpix, sampled


def _check_adaptive(g, img, rb, pix, dbuf, dpix, sampled):
    assert sampled == g["sampled_per_launch"].tolist() and min(sampled) == 0 and max(sampled) > 0
    assert np.array_equal(img, g["image_buffer"])
    assert np.array_equal(rb.view(np.int32), g["ray_buffer"].view(np.int32))
    assert np.array_equal(dbuf, g["diff_buffer"]) and np.array_equal(dpix, g["diff_pixels"])
    assert np.array_equal(pix, g["image_pixels"])


def test_oracle_adaptive_sampling_matches_reference_source():
    # src/ with ADAPTIVE_SAMPLING = True: pathtrace() skips converged pixels (src/pathtracer.py:97-101),
    # post_process() maintains diff_buffer / diff_pixels (src/postprocessor.py:40-43)
    g, oc, oo, env = oracle_of("src_adaptive")
    run = lambda L, img, rb, dpix: po.pathtrace_adaptive(oc, oo, 1, img, rb, dpix, env, sample_base=L)
    _check_adaptive(g, *_adaptive_run(run, g, oc))


def test_product_host_code_adaptive_sampling_matches_reference_source():
    g, cfg, objs, cam, tm, env = common.golden_case("src_adaptive")
    run = lambda L, img, rb, dpix: common.hostcheck_pathtrace(cfg, cam, objs, 1, sample_base=L, image=img, ray_buffer=rb,
                                                              env=env, diff_pixels=dpix)
    _check_adaptive(g, *_adaptive_run(run, g, cfg))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 1])
def test_cuda_adaptive_sampling_matches_reference_source(kernel):
    from raytracingpbr_b200 import PathTracer, _native as N
    g, cfg, objs, cam, tm, env = common.golden_case("src_adaptive")
    cfg.kernel = kernel
    with PathTracer(cfg, objs, cam, tm) as pt:
        pt.set_envmap(env)
        pt.pathtrace(1)                                   # before refresh(): diff_pixels = 0, nothing is sampled
        assert pt.image_buffer.to_numpy()[..., 3].sum() == 0
        pt.ctx.set_sample_base(0)
        pt.refresh()
        for _ in range(int(g["launches"])):
            pt.render(1)                                  # pathtrace(); post_process()
        img, rb = pt.image_buffer.to_numpy(), pt.ray_buffer.to_numpy()
        dbuf = pt.ctx.download(N.BUF_DIFF_BUFFER)
        dpix = pt.ctx.download(N.BUF_DIFF_PIXELS)[..., 0]
        pix = pt.image_pixels.to_numpy()
    assert np.array_equal(img, g["image_buffer"])
    assert np.array_equal(rb.view(np.int32), g["ray_buffer"].view(np.int32))
    assert np.array_equal(dbuf, g["diff_buffer"]) and np.array_equal(dpix, g["diff_pixels"])
    assert np.array_equal(pix, g["image_pixels"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SHORTEST, ids=os.path.basename)
def test_cuda_image_buffer_matches_reference_source(path):
    from raytracingpbr_b200 import PathTracer, _native as N, scenes
    g = np.load(path)
    W, H, B, S, seed = params(g)
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg, objs, cam, tm = scenes.cornell_box_shortest(W, H, max_bounces=B, seed=seed, kernel=kernel)
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.refresh()
            for _ in range(S):                      # one launch per sample, like the reference's main loop
                pt.pathtrace(1)
            pt.post_process()
            buf = pt.image_buffer.to_numpy()
            pix = pt.image_pixels.to_numpy()
        assert np.array_equal(buf, g["image_buffer"])
        # tone mapping uses pow(): tolerance, not bits (shortest:124-129)
        np.testing.assert_allclose(pix, g["image_pixels"], atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FAMILY_B)
def test_cuda_family_b_matches_reference_source(name):
    from raytracingpbr_b200 import PathTracer, _native as N
    g, cfg, objs, cam, tm, env = common.golden_case(name)
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg.kernel = kernel
        with PathTracer(cfg, objs, cam, tm) as pt:
            if env is not None:
                pt.set_envmap(env)
            if "frame" in g:
                pt.ctx.set_frame(int(g["frame"]))
            pt.refresh()
            pt.pathtrace(1)
            first = pt.image_buffer.to_numpy()
            if int(g["spp"]) > 1:
                pt.pathtrace(int(g["spp"]) - 1)
            buf = pt.image_buffer.to_numpy()
            pt.post_process()
            pix = pt.image_pixels.to_numpy()
        assert np.array_equal(first, g["image_buffer_first"]), (name, kernel)
        assert np.array_equal(buf, g["image_buffer"]), (name, kernel)
        if "image_pixels" in g:      # tone mapping (pow): tolerance, not bits; cornell_box.py:372-379 and variants
            ref = g["image_pixels"]
            ok = np.isfinite(ref)       # cornell_box.py takes pow() of slightly negative ACES output: NaN there; the kernel clamps to 0
            want = np.clip(np.where(ok, ref, 0.0), 0.0, 1.0)
            np.testing.assert_allclose(pix, want, atol=3e-5, err_msg=name)


@pytest.mark.gpu
def test_cuda_family_c_matches_reference_source():
    from raytracingpbr_b200 import PathTracer, _native as N
    g, cfg, objs, cam, tm, env = common.golden_case("src_scene")
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg.kernel = kernel
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.set_envmap(env)
            pt.refresh()
            pt.pathtrace(1)                                  # one reference launch
            first = pt.image_buffer.to_numpy()
            for _ in range(int(g["launches"]) - 1):          # launch by launch, like src/renderer.py:29-30
                pt.pathtrace(1)
            buf = pt.image_buffer.to_numpy()
            rb = pt.ray_buffer.to_numpy()
            pt.post_process()                                # src/postprocessor.py:24-38
            np.testing.assert_allclose(pt.image_pixels.to_numpy(), g["image_pixels"], atol=3e-5)
        assert np.array_equal(first, g["image_buffer_first"]), kernel
        assert np.array_equal(buf, g["image_buffer"]), kernel
        assert np.array_equal(rb.view(np.int32), g["ray_buffer"].view(np.int32)), kernel
        # several reference launches replayed inside ONE kernel launch give the same bits
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.set_envmap(env)
            pt.refresh()
            pt.pathtrace(int(g["launches"]))
            assert np.array_equal(pt.image_buffer.to_numpy(), g["image_buffer"]), kernel
            assert np.array_equal(pt.ray_buffer.to_numpy().view(np.int32), g["ray_buffer"].view(np.int32)), kernel


# ---------------------------------------------------------------- bunny_sdf.py / bunny_sdf_v2.py: in-kernel sample loop
INNER = ["bunny_sdf_v2", "bunny_sdf"]


@pytest.mark.parametrize("name", INNER)
def test_oracle_inner_sample_loop_matches_reference_source(name):
    """kernel render() of examples/bunny/bunny_sdf_v2.py:397-432 / bunny_sdf.py: SAMPLE_PER_PIXEL samples per launch on ONE
    ti.random stream per pixel, image_buffer overwritten by every launch, white / black background for camera rays,
    w = 1.6 -> 0.7 relaxation, metal bunny; bunny_sdf.py animates without the bob."""
    g, oc, oo, env = oracle_of(name)
    assert oc.inner_spp == int(g["inner_spp"]) and oc.primary_miss in (1, 2)
    assert np.array_equal(po.pathtrace(oc, oo, 1, env=env), g["image_buffer_first"])
    got = po.pathtrace(oc, oo, int(g["launches"]), env=env)
    assert np.array_equal(got, g["image_buffer"])
    assert (got[..., 3] == float(g["inner_spp"])).all()                        # overwritten, not accumulated
    L = po.lib()
    for p, row in zip(g["sd_points"], g["sd_values"]):                          # signed_distance incl. the animation (with / without bob)
        assert np.float32(L.orc_signed_distance_g(C.byref(oc), oo, len(oo), 0, f32p(p))) == row[0]
    if name == "bunny_sdf_v2":
        assert (got[..., :3] == float(g["inner_spp"])).all(axis=-1).any()      # pure white background pixels exist


@pytest.mark.parametrize("name", INNER)
def test_product_host_code_inner_sample_loop_matches_reference_source(name):
    g, cfg, objs, cam, tm, env = common.golden_case(name)
    img = common.hostcheck_pathtrace(cfg, cam, objs, int(g["launches"]), env=env, frame=int(g["frame"]))
    assert np.array_equal(img, g["image_buffer"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", INNER)
def test_cuda_inner_sample_loop_matches_reference_source(name):
    from raytracingpbr_b200 import PathTracer
    g, cfg, objs, cam, tm, env = common.golden_case(name)
    with PathTracer(cfg, objs, cam, tm) as pt:            # the preset's tonemap dict carries the frame
        pt.set_envmap(env)
        pt.refresh()
        pt.pathtrace(1)
        first = pt.image_buffer.to_numpy()
        pt.pathtrace(int(g["launches"]) - 1)
        buf = pt.image_buffer.to_numpy()
        pt.post_process()
        pix = pt.image_pixels.to_numpy()
    assert np.array_equal(first, g["image_buffer_first"])
    assert np.array_equal(buf, g["image_buffer"])
    np.testing.assert_allclose(pix, np.clip(g["image_pixels"], 0.0, 1.0), atol=3e-5)     # (the file does not clamp; the kernel does)
