import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from raytracingpbr_b200 import PathTracer, _native as N, scenes
import common

def render(kernel, jit, spp=64, size=1024):
    cfg, objs, cam, tm = scenes.cornell_box_shortest(size, size, max_bounces=8, seed=0, kernel=kernel)
    with PathTracer(cfg, objs, cam, tm) as pt:
        pt.ctx.set_jit(jit)
        pt.refresh(); pt.pathtrace(spp)
        return pt.image_buffer.to_numpy()

a = render(N.KERNEL_PERSISTENT, True)
b = render(N.KERNEL_PERSISTENT, False)
c = render(N.KERNEL_SIMPLE, False)
print("jit vs aot-pool:", (a != b).any(-1).sum(), "aot-pool vs simple:", (b != c).any(-1).sum(), "jit vs simple:", (a != c).any(-1).sum())
idx = np.argwhere((a != c).any(-1))
print(idx[:10])
for (i, j) in idx[:3]:
    print(i, j, a[i, j], c[i, j])
    # find the sample: render 1 spp at a time is expensive; use the oracle for this pixel column
    oc, oo = common.to_oracle(*[x for x in (scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)[0],)], scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)[2], scenes.cornell_box_shortest(1024, 1024, max_bounces=8, seed=0)[1])
    want = common.po.pathtrace(oc, oo, 64, i0=int(i), i1=int(i) + 1)
    print("oracle", want[i, j], "jit==oracle", np.array_equal(a[i, j], want[i, j]), "simple==oracle", np.array_equal(c[i, j], want[i, j]))
