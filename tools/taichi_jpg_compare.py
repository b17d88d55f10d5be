"""Background of the glass-bunny picture (others/sdf_bunny_glass.jpg, made by Taichi) against a render with the real
limpopo map: 8 x 8 region means of the regions the bunny never covers.  python tools/taichi_jpg_compare.py [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raytracingpbr_b200 import PathTracer, ibl, scenes

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(root, "tests", "golden", "taichi_bunny_jpg_regions.npz"))
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg, objs, cam, tm = scenes.bunny_glass(1920, 1080, max_bounces=16, seed=1)
path = os.path.join(root, "tests", "assets_local", "limpopo_golf_course_3k.hdr")
hdr = ibl.read_rgbe(path)                                    # (H, W, 3) float, top row first
lay = lambda a: np.ascontiguousarray(a.swapaxes(0, 1)[:, ::-1, :])      # -> (W, H, 3), y up (ibl.imread's layout)


def table(kind):
    """Hypotheses for what `ti.tools.imread(.hdr)` hands to the script (then / 255, * 1.8, ** 2.2 as the script does)."""
    if kind == "stb_gamma":        # the contract: stb_image's 8-bit path, clamp(x^(1/2.2) * 255 + 0.5)
        return ibl.load_envmap(path, 1.8, 2.2)
    if kind == "linear_clamp":     # 8-bit without the gamma: clamp(x * 255 + 0.5)
        u8 = np.clip(hdr * 255.0 + 0.5, 0, 255).astype(np.uint8)
        return ibl.process(lay(u8), 1.8, 2.2)
    if kind == "gamma_noclamp":    # float image, gamma-encoded, not clamped (values / 255 above 1 allowed)
        x = np.power(np.maximum(hdr, 0), 1 / 2.2).astype(np.float32) * np.float32(1.8)
        return lay(np.power(x.astype(np.float64), 2.2).astype(np.float32))
    raise KeyError(kind)


ring = np.ones((8, 8), bool)
ring[1:7, 2:6] = False
np.set_printoptions(precision=1, suppress=True, linewidth=200)
for kind in sys.argv[2:] or ["stb_gamma"]:
    tone = dict(tm)
    if ":" in kind:                                          # kind:exposure overrides camera_exposure
        kind, e = kind.split(":")
        tone["exposure"] = float(e)
    with PathTracer(cfg, objs, cam, tone) as pt:
        pt.set_envmap(table(kind))
        pt.refresh()
        pt.pathtrace(spp)
        pt.post_process()
        pix = pt.image_pixels.to_numpy()
    img = np.floor(np.clip(pix, 0, 1).transpose(1, 0, 2)[::-1] * 255.0)
    means = img.reshape(8, 135, 8, 240, 3).mean(axis=(1, 3))
    d = means - g["region_means"]
    print(f"{kind} exposure {tone['exposure']}: global mean ours {img.mean(axis=(0, 1)).round(1)} jpg {g['global_mean'].round(1)}; "
          f"background ring: mean abs diff {np.abs(d[ring]).mean():.2f}, max {np.abs(d[ring]).max():.2f}, mean signed {d[ring].mean(axis=0).round(1)}")
    if os.environ.get("SHOW"):
        print(d.mean(axis=2))
