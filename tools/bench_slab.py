#!/usr/bin/env python
"""One rank's share of a tile-partitioned 4K progressive frame, timed on one GPU: rank 0 of N (interleaved 8-row slabs),
8 spp x 3 bounces, for the wavefront and the path kernel.  Shows what limits tile-mode strong scaling without needing
N GPUs: ms(N) x N / ms(1) is the inverse efficiency of the render itself (no exchange).
usage: tools/bench_slab.py [scene] [--opt name=value ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from PIL import Image
from minotert_b200 import capi, host, scenes

scene = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "hall_260k"
opts = [a for a in sys.argv[1:] if "=" in a]
w, h, spp, bounces = 3840, 2160, 8, 3
pos, idx, alb, view = getattr(scenes, scene)()
bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
ctx = capi.Context(0)
ctx.upload_blue_noise(bn)
ctx.upload_mesh(pos, idx, alb)
ctx.build()
ctx.atmosphere(host.atmosphere_earth())
ctx.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
for kv in opts:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
stream = torch.cuda.ExternalStream(ctx.stream())
out = {"scene": scene, "options": opts}
for pk in (0, 1):
    ctx.set_option("path_kernel", pk)
    res = {}
    for n in (1, 2, 4, 8, 16):
        ctx.set_partition(0, n, 8)
        def frame(i):
            pc, sc = host.camera_constants(cam, cam, i + 1)
            ctx.primary_rays(w, h, pc)
            ctx.secondary_rays(sc, spp, bounces, capi.SECONDARY_ACCUMULATE if i else 0)
            ctx.tonemap("amd", 1.0, (16.0, 2.0, 1.0, 0.18, 0.18), capi.BUF_ACCUM)
        for i in range(2):
            frame(i)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        steps = 4
        for i in range(steps):
            frame(i)
        e1.record(stream)
        torch.cuda.synchronize()
        res[n] = e0.elapsed_time(e1) / steps
    out["path_kernel" if pk else "wavefront"] = {"ms_per_frame_rank0_of_N": res,
                                                  "render_efficiency": {n: res[1] / (n * res[n]) for n in res}}
print(json.dumps(out))
