"""CPU-only tests: the C-ABI library loads and exports every symbol include/rtpbr.h declares
(no compute calls without a GPU), the product's host-side setup and its __host__ __device__
integrator code (compiled for the CPU by tests/native/hostcheck.cu) agree bit for bit with the
independent C oracle, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from common import po
from conftest import HAVE_GPU
from raytracingpbr_b200 import RtpbrError, _native as N, scenes
from raytracingpbr_b200.dataclass import Camera, Material, SDFObject, Transform
from raytracingpbr_b200.tmath import vec3


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(common.ROOT, "include", "rtpbr.h")).read()
    declared = set(re.findall(r"RTPBR_API\s+[\w\s\*]+?\b(rtpbr_\w+)\s*\(", hdr))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    L = N.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.rtpbr_version() == 2        # RTPBR_VERSION: 2 = RtpbrConfig grew inner_spp / primary_miss / bunny_bob


def test_struct_layouts_match_header():
    L = N.lib()
    assert L.rtpbr_sizeof_config() == C.sizeof(N.RtpbrConfig)
    assert L.rtpbr_sizeof_object() == C.sizeof(N.RtpbrObject) == 80
    assert L.rtpbr_sizeof_camera() == C.sizeof(N.RtpbrCamera) == 52


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    cfg = scenes.cornell_box_shortest(8, 8)[0]
    with pytest.raises(RtpbrError) as e:
        N.Context(cfg)
    assert e.value.code == N.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation_without_gpu():
    L = N.lib()
    h = C.c_void_p()
    assert L.rtpbr_create(None, 0, C.byref(h)) == N.ERR_ARG
    bad = scenes.cornell_box_shortest(8, 8)[0]
    bad.max_bounces = 0
    assert L.rtpbr_create(C.byref(bad), 0, C.byref(h)) == N.ERR_ARG
    assert b"max_bounces" in L.rtpbr_last_error()
    assert L.rtpbr_pathtrace(None, 1) == N.ERR_ARG
    assert L.rtpbr_destroy(None) == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(common.ROOT, "raytracingpbr_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                for needle in ("pyoracle", "liboracle", "import oracle", "from oracle", "oracle.c", "oracle/"):
                    assert needle not in txt, (f, needle)


def test_dataclass_surface_positional_like_reference():
    # src/scene.py:13-14 construction style
    o = SDFObject(2, Transform(vec3(0, 0, -1), vec3(0), vec3(1, 1, 0.2)), Material(vec3(1, 1, 1) * 0.6, vec3(1), 1.0, 1.0, 0, 1.100))
    n = o.to_native()
    assert n.type == 2 and list(n.scale) == [1.0, 1.0, np.float32(0.2)] and list(n.albedo) == [np.float32(0.6)] * 3
    assert n.ior == np.float32(1.1) and n.metallic == 1.0
    c = Camera(vec3(0, 0, 3.5), vec3(0, 0, -1), vec3(0, 1, 0), 35.0, 1.0, 0.0, 1.0).to_native()
    assert list(c.lookfrom) == [0, 0, 3.5] and c.vfov == 35.0


def test_host_transcendentals_match_oracle_bitwise():
    H, L = common.hostcheck(), po.lib()
    s1, c1, s2, c2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    for x in np.linspace(-40, 40, 4001).astype(np.float32):
        H.hostcheck_sincos(float(x), C.byref(s1), C.byref(c1))
        L.orc_sincosf(float(x), C.byref(s2), C.byref(c2))
        assert s1.value == s2.value and c1.value == c2.value
    rng = np.random.default_rng(3)
    for y, x in rng.normal(size=(2000, 2)).astype(np.float32):
        assert H.hostcheck_atan2(float(y), float(x)) == L.orc_atan2f(float(y), float(x))
    for x in np.linspace(-1.1, 1.1, 1001).astype(np.float32):
        assert H.hostcheck_asin(float(x)) == L.orc_asinf(float(x))


def test_host_euler_matrix_matches_oracle_bitwise():
    H, L = common.hostcheck(), po.lib()
    a, b = np.zeros(9, np.float32), np.zeros(9, np.float32)
    f32p = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))
    for rot in ([0, 0, 0], [90, 0, 0], [0, 112, 0], [0, -197, 0], [13.5, -77.25, 211.0]):
        r = np.array(rot, np.float32)
        H.hostcheck_euler(f32p(r), f32p(a))
        L.orc_angle_deg(f32p(r), f32p(b))
        assert np.array_equal(a, b), rot


@pytest.mark.parametrize("w,h,spp,b,seed", [(48, 40, 2, 4, 0), (33, 17, 3, 8, 5), (5, 3, 4, 16, 9)])
def test_product_integrator_on_host_matches_oracle_bitwise(w, h, spp, b, seed):
    cfg, objs, cam, _ = scenes.cornell_box_shortest(w, h, max_bounces=b, seed=seed)
    got = common.hostcheck_pathtrace(cfg, cam, objs, spp)
    oc, oo = common.to_oracle(cfg, cam, objs)
    want = po.pathtrace(oc, oo, spp)
    assert np.array_equal(got, want)


def test_product_shard_mapping_on_host():
    cfg, objs, cam, _ = scenes.cornell_box_shortest(70, 20, max_bounces=4, seed=1)
    full = common.hostcheck_pathtrace(cfg, cam, objs, 2)
    acc = np.zeros_like(full)
    for r in range(4):
        part = common.hostcheck_pathtrace(cfg, cam, objs, 2, rank=r, nranks=4, band=8)
        own = ((np.arange(70) // 8) % 4) == r
        assert (part[~own] == 0).all() and (part[own][..., 3] == 2).all()
        acc += part
    assert np.array_equal(acc, full)


def test_src_shaped_surface_imports_without_side_effects():
    # module layout of the reference's src/ package; importing creates no GPU context (lazy runtime)
    from raytracingpbr_b200.src import _runtime, camera, config, fileds, pathtracer, postprocessor, renderer, scene, sdf
    assert _runtime._pt is None
    assert sdf.SHAPE.SPHERE == 1 and sdf.SHAPE.PLANE == 5                       # src/sdf.py:12-18
    assert len(scene.OBJECTS) == 7 and [o.type for o in scene.OBJECTS] == sorted(o.type for o in scene.OBJECTS)   # src/scene.py:33
    assert fileds.image_buffer.shape == tuple(config.image_resolution) == (768, 432)
    assert camera.camera_vfov[None] == 35.0 and camera.camera_aperture[None] == 0.01 and camera.camera_focus[None] == 4.0
    for fn in (renderer.render, renderer.refresh, pathtracer.pathtrace, postprocessor.post_process, scene.build_scene):
        assert callable(fn)


def test_smooth_camera_easing_follows_reference_update():
    # src/camera.py:82-112: position += (target - position) * clamp(velocity * dt, 0, 1); moving flag; u_frame += 1
    from raytracingpbr_b200.src import camera, fileds

    class Cam:
        curr_position = np.array([1, -0.2, 4], np.float32)
        curr_lookat = np.array([0, -0.2, 3], np.float32)
        curr_up = np.array([0, 1, 0], np.float32)
    s = camera.SmoothCamera()
    f0 = fileds.u_frame[None]
    s.update(0.05, Cam)
    np.testing.assert_allclose(s.position[None], [0.5, -0.2, 4.0], atol=1e-6)
    assert s.moving[None] == 1 and fileds.u_frame[None] == f0 + 1
    for _ in range(60):
        s.update(0.05, Cam)
    np.testing.assert_allclose(s.position[None], [1, -0.2, 4], atol=1e-5)
    assert s.moving[None] == 0
    s.update(0.5, Cam)                                # velocity * dt > 1 clamps to 1
    np.testing.assert_allclose(s.position[None], [1, -0.2, 4], atol=1e-6)


def test_bench_keeps_library_output_off_stdout():
    """bench.py's contract is ONE JSON line on stdout; NCCL prints its version banner to file descriptor 1 when the box sets
    NCCL_DEBUG.  After _only_json_on_stdout() raw writes to fd 1 land on stderr and only Python's prints reach stdout."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, bench; bench._only_json_on_stdout(); os.write(1, b'NCCL version x\\n'); "
            "print('{\"metric\": 1}', flush=True)")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": 1}\n'
    assert "NCCL version x" in r.stderr


def test_bench_reference_arm_sample_fits_its_budget():
    """--impl reference: the per-step sample shrinks with --steps so that all timed steps fit ~150 s of CPU time."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    per_band = bench.H * bench.CPU_SPP * bench.CPU_COLS / 1e6            # Msamples of one band
    assert bench.reference_bands(0.30, 5) == bench.CPU_BANDS             # the GPU box's 16 cores, default K: the full 64 columns
    assert bench.reference_bands(0.30, 20) == 16
    assert bench.reference_bands(0.0, 20) == bench.CPU_BANDS             # no calibration: the default sample
    for rate in (0.02, 0.07, 0.3, 1.0, 5.0):
        for steps in (1, 2, 5, 20, 100):
            b = bench.reference_bands(rate, steps)
            assert 4 <= b <= bench.CPU_BANDS and (b & (b - 1)) == 0
            assert b == 4 or b * per_band / rate * steps <= bench.REF_BUDGET_S * 1.0001
