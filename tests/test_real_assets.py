"""The reference's own environment maps (assets/*.hdr, 3200 x 1600 Radiance files).

They are not part of this repository.  __graft_entry__.build() stages copies under tests/assets_local/ (git-ignored)
whenever /root/reference/assets is present, so that these tests -- and bench.py's C2 / C3 blocks -- run with the real
maps on the GPU box too; everything here skips when the files are absent.

  * read_rgbe() against OpenCV's independent Radiance decoder (exact equality), imread()'s layout;
  * the product's integrator compiled for the host against the oracle on columns of the FULL-SIZE C3 / C2 images lit by
    the real maps;
  * (GPU) both CUDA kernels against the oracle on the same columns."""
import os

import numpy as np
import pytest

import common
from oracle import pyoracle as po
from raytracingpbr_b200 import ibl, scenes

ASSET_DIRS = [os.environ.get("RTPBR_ASSETS"), os.path.join(common.ROOT, "tests", "assets_local")]   # staged by __graft_entry__.build()
FILES = {"tokyo": "Tokyo_BigSight_3k.hdr", "limpopo": "limpopo_golf_course_3k.hdr"}


def asset(name):
    for d in ASSET_DIRS:
        if d and os.path.exists(os.path.join(d, FILES[name])):
            return os.path.join(d, FILES[name])
    pytest.skip(f"{FILES[name]} not available (reference assets are not redistributed)")


@pytest.mark.parametrize("name", list(FILES))
def test_rgbe_decoder_equals_opencv_on_the_real_assets(name):
    cv2 = pytest.importorskip("cv2")
    path = asset(name)
    mine = ibl.read_rgbe(path)                                          # (H, W, 3) RGB, top row first
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)                        # (H, W, 3) BGR float32
    assert ref is not None and ref.dtype == np.float32
    assert mine.shape == (1600, 3200, 3)
    assert np.array_equal(mine, ref[..., ::-1])
    assert float(mine.max()) > 50.0                                     # a real HDR range (55.25 / 86016)


@pytest.mark.parametrize("name", list(FILES))
def test_imread_layout_on_the_real_assets(name):
    cv2 = pytest.importorskip("cv2")
    path = asset(name)
    u8 = ibl.imread(path)
    assert u8.shape == (3200, 1600, 3) and u8.dtype == np.uint8         # width first, y up (src/ibl.py:15-16)
    assert u8.max() == 255 and len(np.unique(u8)) > 200
    # pixel (x, y) of the table is column x of the file's row H-1-y (first file row = top of the panorama = largest y)
    rgb = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1]
    for x, y in ((0, 0), (17, 1599), (3199, 800), (1234, 321)):
        assert np.array_equal(u8[x, y], ibl.hdr_to_ldr_stb(rgb[1599 - y, x][None, None])[0, 0])


CASES = {
    # name: (preset, width, height, bounces, asset, exposure, gamma, columns)
    "c3_real_env": ("tokyo_ibl", 1920, 1080, 8, "tokyo", 1.8, 2.2, (3, 601, 960, 1300, 1917)),
    "c2_real_env": ("bunny_glass", 1024, 1024, 16, "limpopo", 1.8, 2.2, (300, 512, 700)),
}
_ENV = {}


def real_case(name):
    preset, w, h, bounces, a, exposure, gamma, cols = CASES[name]
    if a not in _ENV:
        _ENV[a] = ibl.load_envmap(asset(a), exposure, gamma)
    cfg, objs, cam, tm = getattr(scenes, preset)(w, h, max_bounces=bounces, seed=11)
    return cfg, objs, cam, tm, _ENV[a], list(cols)


def oracle_columns(cfg, objs, cam, env, cols, spp):
    oc, oo = common.to_oracle(cfg, cam, objs)
    buf = np.zeros((cfg.width, cfg.height, 4), np.float32)
    for c0 in cols:
        po.pathtrace(oc, oo, spp, env=env, i0=int(c0), i1=int(c0) + 1, image_buffer=buf)
    return buf[cols]


def test_processed_real_map_is_clamped_at_exposure_to_the_gamma():
    cfg, objs, cam, tm, env, cols = real_case("c3_real_env")
    assert env.shape == (3200, 1600, 3) and env.dtype == np.float32
    assert abs(float(env.max()) - 1.8 ** 2.2) < 1e-4                    # 8-bit LDR re-linearised


@pytest.mark.parametrize("name", ["c3_real_env"])
def test_product_host_code_equals_the_oracle_under_the_real_map(name):
    cfg, objs, cam, tm, env, cols = real_case(name)
    want = oracle_columns(cfg, objs, cam, env, cols[:2], 1)
    # (the host-compiled integrator renders whole images: one band of columns keeps the CPU suite short)
    got = np.zeros((cfg.width, cfg.height, 4), np.float32)
    for c0 in cols[:2]:
        band = c0 // 32
        nb = (cfg.width + 31) // 32
        common.hostcheck_pathtrace(cfg, cam, objs, 1, image=got, rank=band, nranks=nb, band=32, env=env)
    assert np.array_equal(got[cols[:2]], want)
    assert (want[..., :3].sum(-1) > 0).mean() > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_equals_the_oracle_under_the_real_map(name):
    from raytracingpbr_b200 import PathTracer, _native as N
    cfg, objs, cam, tm, env, cols = real_case(name)
    spp = 2
    want = oracle_columns(cfg, objs, cam, env, cols, spp)
    for kernel in (N.KERNEL_PERSISTENT, N.KERNEL_SIMPLE):
        cfg.kernel = kernel
        with PathTracer(cfg, objs, cam, tm) as pt:
            pt.set_envmap(env)
            pt.refresh()
            pt.pathtrace(spp)
            got = pt.image_buffer.to_numpy()[cols]
        assert np.array_equal(got, want), (name, kernel)
    assert len(np.unique(want[..., :3])) > 1000                          # lit by a real sky, not a constant
