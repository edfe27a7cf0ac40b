"""diagnostic: where does a tile-partitioned 4K progressive render differ from the 1-context render?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from minotert_b200 import capi, scenes
from test_gpu_spheres import as_capi, setup_sky
import test_gpu_group as G

pk = int(sys.argv[1]) if len(sys.argv) > 1 else 0
atmo = O.earth(); bn = O.load_blue_noise()
pos, idx, alb, view = scenes.hall_260k()
w, h, spp, bounces, frames = 3840, 2160, 8, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 8
cam = O.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
orig = capi.Context.__init__
def init(self, *a, **k):
    orig(self, *a, **k)
    self.set_option("path_kernel", pk)
capi.Context.__init__ = init
a = G.single_context_image(O, atmo, cam, bn, (pos, idx, alb), w, h, spp, bounces, frames)
b = G.single_context_image(O, atmo, cam, bn, (pos, idx, alb), w, h, spp, bounces, frames)
print("path_kernel", pk, "1-ctx determinism: acc equal", np.array_equal(a[1], b[1]), "diff px", (a[1] != b[1]).any(-1).sum())
for n in (2,):
    g = capi.Group([0] * n, transport="p2p")
    got = G.group_image(g, O, atmo, cam, bn, (pos, idx, alb), w, h, spp, bounces, frames, 8)
    g.close()
    d = (got[1] != a[1]).any(-1)
    print(n, "ranks: ldr eq", np.array_equal(got[0], a[0]), "vis eq", np.array_equal(got[2], a[2]), "acc diff px", d.sum(), "rows", np.unique(np.where(d)[0])[:20],
          "max abs", np.abs(got[1] - a[1]).max())
    ys, xs = np.where(d)
    for k in range(min(5, len(ys))):
        print(ys[k], xs[k], got[1][ys[k], xs[k]], a[1][ys[k], xs[k]])
