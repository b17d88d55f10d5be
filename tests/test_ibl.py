"""Environment-map input side (raytracingpbr_b200/ibl.py): RGBE decode, stb-style LDR conversion, process()."""
import numpy as np

from raytracingpbr_b200 import ibl


def test_rgbe_round_trip_flat(tmp_path):
    rng = np.random.default_rng(0)
    img = (rng.random((6, 9, 3)) ** 4 * 50).astype(np.float32)
    p = str(tmp_path / "t.hdr")
    ibl.write_rgbe(p, img)
    back = ibl.read_rgbe(p)
    assert back.shape == img.shape
    # 8-bit mantissa shared exponent: relative error of the largest channel < 1/128
    m = img.max(axis=2)
    assert (np.abs(back - img).max(axis=2) <= m / 128 + 1e-6).all()


def test_rgbe_rle_scanlines(tmp_path):
    # hand-built new-style RLE file: one 8-pixel scanline, every channel a single run
    p = tmp_path / "rle.hdr"
    body = bytes([2, 2, 0, 8]) + bytes([128 + 8, 64]) + bytes([128 + 8, 32]) + bytes([128 + 8, 16]) + bytes([128 + 8, 129])
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 1 +X 8\n" + body)
    img = ibl.read_rgbe(str(p))
    assert img.shape == (1, 8, 3)
    np.testing.assert_allclose(img[0, 3], [64 * 2.0 ** -7, 32 * 2.0 ** -7, 16 * 2.0 ** -7])


def test_ldr_conversion_and_orientation(tmp_path):
    img = np.zeros((2, 3, 3), np.float32)
    img[0, 0] = [1.0, 0.25, 0.0]          # top-left pixel
    p = str(tmp_path / "o.hdr")
    ibl.write_rgbe(p, img)
    u8 = ibl.imread(p)                    # (W, H, 3), y up
    assert u8.shape == (3, 2, 3) and u8.dtype == np.uint8
    assert u8[0, 1, 0] == 255             # top-left ends up at x = 0, y = H-1
    assert u8[0, 1, 1] == int(0.25 ** (1 / 2.2) * 255 + 0.5)
    assert u8[0, 0].sum() == 0


def test_process_is_double_pow_rounded_once():
    u8 = np.arange(0, 256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    t = ibl.process(u8, 1.8, 2.2)
    x = (u8.astype(np.float32) / np.float32(255)) * np.float32(1.8)
    want = np.array([[[float(v) ** float(np.float32(2.2)) for v in px] for px in row] for row in x.astype(np.float64)]).astype(np.float32)
    assert np.array_equal(t, want) and t.dtype == np.float32
    assert abs(float(t.max()) - 1.8 ** 2.2) < 1e-5      # environment radiance is clamped at exposure^gamma
