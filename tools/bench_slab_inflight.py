#!/usr/bin/env python
"""Would frames in flight help tile mode?  K frame contexts share one scene; each renders rank 0's slabs of N (4K, 8 spp x 3
bounces, independent frames), issued round-robin.  Wall-clock ms per frame for K = 1, 2, 3 and N = 1, 8, with the traversal
grids capped at ceil(6 / K) CTAs per SM or not.
usage: tools/bench_slab_inflight.py [scene]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from PIL import Image
from minotert_b200 import capi, host, scenes

scene = sys.argv[1] if len(sys.argv) > 1 else "scene_10m"
w, h, spp, bounces = 3840, 2160, 8, 3
pos, idx, alb, view = getattr(scenes, scene)()
bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
ctxs = []
KMAX = 6
for k in range(KMAX):
    c = capi.Context(0)
    c.upload_blue_noise(bn)
    if k == 0:
        c.upload_mesh(pos, idx, alb); c.build()
    else:
        c.share_scene(ctxs[0])
    c.atmosphere(host.atmosphere_earth())
    c.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
    ctxs.append(c)
out = {}
for n, pk in ((8, 0), (8, 1)):
    for c in ctxs: c.set_partition(0, n, 8)
    for c in ctxs: c.set_option("path_kernel", pk)
    for K, cap in ((1, 0), (2, 3), (3, 2), (3, 3), (4, 2), (4, 1), (6, 1), (6, 2)):
        if True:
            if K == 1 and cap: continue
            for c in ctxs: c.set_option("trace_ctas_per_sm", cap)
            def frame(i):
                c = ctxs[i % K]
                pc, sc = host.camera_constants(cam, cam, i + 1)
                c.primary_rays(w, h, pc)
                c.secondary_rays(sc, spp, bounces, 0)
                c.tonemap("amd", 1.0, (16.0, 2.0, 1.0, 0.18, 0.18), capi.BUF_ACCUM)
            for i in range(2 * K): frame(i)
            for c in ctxs: c.sync()
            steps = 12 if n > 1 else 6
            t0 = time.perf_counter()
            for i in range(steps): frame(i)
            for c in ctxs: c.sync()
            out[f"N{n}_pk{pk}_K{K}_cap{cap}"] = round(1e3 * (time.perf_counter() - t0) / steps, 3)
            print(f"N{n}_pk{pk}_K{K}_cap{cap}", out[f"N{n}_pk{pk}_K{K}_cap{cap}"], flush=True)
print(json.dumps(out))
