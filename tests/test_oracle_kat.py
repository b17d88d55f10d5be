"""Known-answer tests that pin the CPU oracle (oracle/oracle.c) to the reference's formulas
 and Philox4x32-10 to the Random123 vectors."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import common
from common import po


def f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def arr(*v):
    return np.array(v, dtype=np.float32)


@pytest.mark.parametrize("ctr,key,expect", [
    # Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_known_answers(ctr, key, expect):
    out = (C.c_uint32 * 4)()
    po.lib().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    assert list(out) == expect


def test_uniform_conversion_is_taichi_rule():
    # f32 = (u32 >> 8) * 2^-24 in [0, 1)
    out = np.zeros(4, dtype=np.float32)
    raw = (C.c_uint32 * 4)()
    L = po.lib()
    L.orc_draw4(5, 11, 3, 2, 0, f32p(out))
    L.orc_philox4x32_10((C.c_uint32 * 4)(11, 3, 2, 0), (C.c_uint32 * 2)(5, 0x52545042), raw)
    want = [np.float32((u >> 8) * 2.0 ** -24) for u in raw]
    assert out.tolist() == want
    assert all(0.0 <= v < 1.0 for v in out)


def test_sequential_stream_contract():
    # n-th ti.random() of (pixel, launch) = word n & 3 of Philox(counter = (pixel, launch, n >> 2, 0))
    L = po.lib()
    raw = (C.c_uint32 * 4)()
    for n in range(11):
        L.orc_philox4x32_10((C.c_uint32 * 4)(77, 5, n >> 2, 0), (C.c_uint32 * 2)(9, 0x52545042), raw)
        assert L.orc_random(9, 77, 5, n) == np.float32((raw[n & 3] >> 8) * 2.0 ** -24)


def test_sincos_accuracy():
    L = po.lib()
    s, c = C.c_float(), C.c_float()
    xs = np.concatenate([np.linspace(-60, 60, 40001), np.linspace(0, 2 * math.pi, 20001)]).astype(np.float32)
    err = 0.0
    for x in xs:
        L.orc_sincosf(float(x), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - math.sin(float(x))), abs(c.value - math.cos(float(x))))
    assert err < 1.5e-7


def test_atan2_asin_accuracy():
    L = po.lib()
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(20000, 2)).astype(np.float32)
    err = max(abs(L.orc_atan2f(float(y), float(x)) - math.atan2(float(y), float(x))) for y, x in pts)
    assert err < 5e-7
    xs = np.linspace(-1, 1, 20001).astype(np.float32)
    err = max(abs(L.orc_asinf(float(x)) - math.asin(float(x))) for x in xs)
    assert err < 5e-7
    assert L.orc_asinf(1.5) == pytest.approx(math.pi / 2, abs=1e-6)      # clamped, not NaN
    assert L.orc_atan2f(0.0, -1.0) == pytest.approx(math.pi, abs=1e-6)
    assert L.orc_atan2f(0.0, 0.0) == 0.0


def test_sd_box_axis_points():
    # src/sdf.py:31-34 / cornell_box_shortest.py:44-45 at points where the answer is obvious
    L = po.lib()
    b = arr(1, 2, 3)
    assert L.orc_sd_box(f32p(arr(2, 0, 0)), f32p(b), 0.0) == 1.0           # outside along x
    assert L.orc_sd_box(f32p(arr(0, 0, 0)), f32p(b), 0.0) == -1.0          # centre: -min half extent
    assert L.orc_sd_box(f32p(arr(0, 2.5, 0)), f32p(b), 0.0) == 0.5
    assert L.orc_sd_box(f32p(arr(4, 6, 3)), f32p(b), 0.0) == 5.0           # corner distance (3,4,0)
    assert L.orc_sd_box(f32p(arr(2, 0, 0)), f32p(b), 0.03) == pytest.approx(0.97, abs=1e-7)  # rounded box


def test_rotation_matrix_90_degrees():
    # src/util.py:36-42 rotate / shortest:34-39 angle: Rz @ Ry @ Rx, row-major
    L = po.lib()
    out = np.zeros(9, dtype=np.float32)
    L.orc_angle_deg(f32p(arr(0, 0, 0)), f32p(out))
    assert np.array_equal(out.reshape(3, 3), np.eye(3, dtype=np.float32))
    L.orc_angle_deg(f32p(arr(90, 0, 0)), f32p(out))
    np.testing.assert_allclose(out.reshape(3, 3), [[1, 0, 0], [0, 0, 1], [0, -1, 0]], atol=1e-7)
    L.orc_angle_deg(f32p(arr(0, 90, 0)), f32p(out))
    np.testing.assert_allclose(out.reshape(3, 3), [[0, 0, -1], [0, 1, 0], [1, 0, 0]], atol=1e-7)
    L.orc_angle_deg(f32p(arr(0, 0, 90)), f32p(out))
    np.testing.assert_allclose(out.reshape(3, 3), [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], atol=1e-7)
    # general angles against a float64 evaluation of the same product
    rot = np.radians([20.0, -197.0, 33.0])
    s, c = np.sin(rot), np.cos(rot)
    A = np.array([[c[2], s[2], 0], [-s[2], c[2], 0], [0, 0, 1]])
    B = np.array([[c[1], 0, -s[1]], [0, 1, 0], [s[1], 0, c[1]]])
    Cm = np.array([[1, 0, 0], [0, c[0], s[0]], [0, -s[0], c[0]]])
    L.orc_angle_deg(f32p(arr(20, -197, 33)), f32p(out))
    np.testing.assert_allclose(out.reshape(3, 3), A @ B @ Cm, atol=3e-7)


def test_camera_centre_ray_and_corners():
    # shortest:107-118 with the camera of shortest:135: centre ray looks down -z
    cfg = po.cornell_shortest_config(512, 512)
    ro, rd = np.zeros(3, np.float32), np.zeros(3, np.float32)
    po.lib().orc_camera_ray(C.byref(cfg), 256, 256, 0.0, 0.0, f32p(ro), f32p(rd))
    assert ro.tolist() == [0.0, 0.0, 3.5]
    np.testing.assert_allclose(rd, [0, 0, -1], atol=1e-7)
    po.lib().orc_camera_ray(C.byref(cfg), 0, 0, 0.0, 0.0, f32p(ro), f32p(rd))
    h = math.tan(math.radians(35) / 2)
    want = np.array([-h, -h, -1.0]) / math.sqrt(2 * h * h + 1)
    np.testing.assert_allclose(rd, want, atol=2e-7)


def test_rr_probability_table():
    # shortest:84-85: p_i = 1 - 1/exp(i/128); p_0 = 0 exactly
    cfg = po.cornell_shortest_config()
    L = po.lib()
    assert L.orc_rr_prob(C.byref(cfg), 0) == 0.0
    for i in (1, 2, 7, 100):
        assert L.orc_rr_prob(C.byref(cfg), i) == pytest.approx(1 - math.exp(-i / 128.0), abs=1e-7)


def test_hemispheric_sampling_unit_and_in_hemisphere():
    L = po.lib()
    rng = np.random.default_rng(1)
    n = arr(0, 1, 0)
    out = np.zeros(3, np.float32)
    for u1, u2 in rng.random((500, 2)):
        L.orc_hemispheric_sampling(f32p(n), float(u1), float(u2), f32p(out))
        assert abs(float(np.linalg.norm(out.astype(np.float64))) - 1.0) < 1e-6
        assert out[1] >= -1e-6
    # (sin, cos) order for (x, y): u2 = 0 -> a = 0 -> xy = s * (0, 1)
    L.orc_hemispheric_sampling(f32p(arr(0, 0, 0.0)), 0.5, 0.0, f32p(out))
    np.testing.assert_allclose(out, [0, 1, 0], atol=1e-7)


def test_nearest_and_normal_in_cornell_box():
    cfg = po.cornell_shortest_config()
    objs = po.objects_array(po.cornell_shortest_objects())
    L = po.lib()
    d = C.c_float()
    # centre of the room, just below the light (light bottom face at y = 0.809 - 0.01)
    idx = L.orc_nearest(C.byref(cfg), objs, 8, f32p(arr(0, 0.7, 0)), C.byref(d))
    assert idx == 7 and d.value == pytest.approx(0.099, abs=1e-6)
    # near the floor (top face of wall 3 at y = -0.8)
    idx = L.orc_nearest(C.byref(cfg), objs, 8, f32p(arr(0.6, -0.75, 0.6)), C.byref(d))
    assert idx == 2 and d.value == pytest.approx(0.05, abs=1e-6)
    n = np.zeros(3, np.float32)
    L.orc_calc_normal(C.byref(cfg), objs, 8, 2, f32p(arr(0.6, -0.8, 0.6)), f32p(n))
    np.testing.assert_allclose(n, [0, 1, 0], atol=1e-5)
    L.orc_calc_normal(C.byref(cfg), objs, 8, 3, f32p(arr(-0.8, 0.1, 0.2)), f32p(n))   # red wall faces +x
    np.testing.assert_allclose(n, [1, 0, 0], atol=1e-5)


def test_raycast_hits_back_wall():
    cfg = po.cornell_shortest_config()
    objs = po.objects_array(po.cornell_shortest_objects())
    out = np.zeros(7, np.float32)
    po.lib().orc_raycast(C.byref(cfg), objs, 8, f32p(arr(0, 0.5, 3.5)), f32p(arr(0, 0, -1)), f32p(out))
    hit, index, steps, dist = out[:4]
    assert hit == 1.0 and index == 0                       # wall 1 (back wall, front face z = -0.8)
    assert dist == pytest.approx(4.3, abs=1e-4) and steps < 256
    po.lib().orc_raycast(C.byref(cfg), objs, 8, f32p(arr(0, 0, 3.5)), f32p(arr(0, 0, 1)), f32p(out))
    assert out[0] == 0.0                                   # looking away: miss


def test_as_written_and_hoisted_agree_bitwise():
    cfg = po.cornell_shortest_config(24, 20, 4, seed=3)
    objs = po.cornell_shortest_objects()
    a = po.pathtrace(cfg, objs, 2, hoisted=True)
    b = po.pathtrace(cfg, objs, 2, hoisted=False)
    assert np.array_equal(a, b)
    assert (a[..., 3] == 2.0).all()


def test_oracle_is_thread_count_independent_and_progressive():
    cfg = po.cornell_shortest_config(32, 16, 8, seed=9)
    objs = po.cornell_shortest_objects()
    a = po.pathtrace(cfg, objs, 4, nthreads=1)
    b = po.pathtrace(cfg, objs, 4, nthreads=5)
    assert np.array_equal(a, b)
    c = po.pathtrace(cfg, objs, 3, sample_base=0)
    c = po.pathtrace(cfg, objs, 1, sample_base=3, image_buffer=c)
    assert np.array_equal(a, c)


def test_white_furnace_radiance_is_bounded():
    # closed box of albedo-1 walls, emission 1 everywhere: every path carries radiance 1 until
    # RR scales it by p_i < 1 -> pixel means in (0, 1]; shortest:84-99
    objs = po.cornell_shortest_objects()
    for o in objs:
        o.albedo[:] = [1, 1, 1]
        o.emission[:] = [1, 1, 1]
    cfg = po.cornell_shortest_config(16, 16, 8)
    img = po.pathtrace(cfg, objs, 8)
    mean = img[..., :3] / img[..., 3:]
    # camera sits outside the box front (open side): escaping paths are black, inside ones <= 1
    assert mean.max() <= 1.0 + 1e-6 and mean.min() >= 0.0


def test_two_ahead_draws_are_the_same_stream():
    """rng_next2_ahead (hit shading of family A) returns exactly the numbers three rng_next calls return, for every
    position of the draws inside a Philox block, and matches the oracle's stream."""
    import ctypes as C
    L = common.hostcheck()
    L.hostcheck_rng3.argtypes = [C.c_uint32] * 4 + [C.POINTER(C.c_float)]
    L.hostcheck_rng3.restype = None
    rng = np.random.default_rng(3)
    out = (C.c_float * 7)()
    for n in list(range(0, 12)) + [int(x) for x in rng.integers(0, 2 ** 20, 40)]:
        seed, pixel, launch = (int(x) for x in rng.integers(0, 2 ** 32, 3, dtype=np.uint64))
        L.hostcheck_rng3(seed, pixel, launch, n, out)
        v = np.frombuffer(out, dtype=np.float32).copy()
        assert v[6] == 3.0
        assert np.array_equal(v[0:3].view(np.uint32), v[3:6].view(np.uint32)), (n, v)
        want = []
        for k in range(3):                                   # the oracle's Philox + Taichi's u32 -> f32 rule
            raw = (C.c_uint32 * 4)()
            po.lib().orc_philox4x32_10((C.c_uint32 * 4)(pixel, launch, (n + k) >> 2, 0), (C.c_uint32 * 2)(seed, 0x52545042), raw)
            want.append(np.float32(raw[(n + k) & 3] >> 8) * np.float32(2.0 ** -24))
        assert np.array_equal(np.asarray(want, np.float32).view(np.uint32), v[0:3].view(np.uint32))


def test_denoise_restatement_against_a_python_transcription():
    """orc_denoise vs a line-by-line Python transcription of kernel denoise() (examples/denoise/denoise_test_1.py:86-118)
    with the reads taken from the previous output (the deterministic form), on a small field."""
    import numpy as np
    from oracle import pyoracle as po
    f32 = np.float32
    rng = np.random.default_rng(3)
    W, H, thr = 7, 5, f32(0.3)
    pin = (rng.random((W, H, 3)) ** 2).astype(f32)
    pin[rng.random((W, H)) < 0.4] *= f32(0.05)
    prev = rng.random((W, H, 3)).astype(f32)

    def fma(a, b, c):                    # one rounding, like fmaf
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    def bright(c):                       # dot contract: fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))
        return fma(c[2], f32(0.114), fma(c[1], f32(0.587), f32(c[0] * f32(0.299))))

    want = np.empty_like(pin)
    for i in range(W):
        for j in range(H):
            p1, p2 = pin[i, j], prev[i, j]
            col = np.array([f32(f32(a * f32(f32(1.0) - f32(0.2))) + f32(b * f32(0.2))) for a, b in zip(p1, p2)], f32)   # mix
            if bright(p1) < thr:
                sur = [prev[min(i + 1, W - 1), j], prev[max(i - 1, 0), j], prev[i, min(j + 1, H - 1)], prev[i, min(j + 1, H - 1)]]
                s, n = np.zeros(3, f32), f32(0)
                for q in sur:
                    if bright(q) > thr:
                        s, n = (s + q).astype(f32), f32(n + 1)
                with np.errstate(invalid="ignore", divide="ignore"):
                    col = (s / n).astype(f32)
            want[i, j] = col
    got = po.denoise(pin, prev, float(thr))
    assert np.array_equal(got, want, equal_nan=True)


def test_bunny_mlp_sine_arguments_are_bounded():
    """Precondition of the device's packed sine (sin2_rt): every pre-activation of the neural bunny stays far below
    2^22 * pi/2.  Bound from the weight tables (csrc/bunny_weights.h): inputs |p| <= 1 per coordinate (the MLP is only
    evaluated inside the unit sphere), first activations in [-1, 1], second in [-2, 2] (sin + residual)."""
    import re
    import numpy as np
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytracingpbr_b200", "csrc", "bunny_weights.h")).read()
    txt = txt.replace("\\\n", " ")

    def table(name):
        m = re.search(r"#define BUNNY_%s_INIT (.*)" % name, txt)
        return np.array([float(v) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?(?=f)", m.group(1))])
    wy, wz, wx, b1 = table("WY"), table("WZ"), table("WX"), table("B1")
    m2, b2, m3, b3 = table("M2").reshape(4, 4, 16), table("B2"), table("M3").reshape(4, 4, 16), table("B3")
    assert len(wy) == 16 and m2.size == 256 and m3.size == 256
    x1 = np.abs(wy) + np.abs(wz) + np.abs(wx) + np.abs(b1)
    # out[4g + j] = sum_h sum_k in[4h + k] * M[g][h][4k + j] + B[4g + j]
    col2 = np.abs(m2).reshape(4, 4, 4, 4).sum(axis=(1, 2)).reshape(16)      # sum over (h, k) for every (g, j)
    col3 = np.abs(m3).reshape(4, 4, 4, 4).sum(axis=(1, 2)).reshape(16)
    x2 = 1.0 * col2 + np.abs(b2)
    x3 = 2.0 * col3 + np.abs(b3)
    worst = max(x1.max(), x2.max(), x3.max()) * 1.001
    assert worst < 64.0, worst
    assert worst * 2 / np.pi < 2 ** 22 / 1e4
