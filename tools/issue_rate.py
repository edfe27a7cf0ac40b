#!/usr/bin/env python
"""Host issue time vs device time of one frame (is the frame loop launch-bound?).
usage: tools/issue_rate.py [scene] [w] [h] [spp] [bounces] [name=value ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from minotert_b200 import capi, host, scenes
import oracle_lib as O

scene = sys.argv[1] if len(sys.argv) > 1 else "hall_260k"
w, h, spp, bounces = (int(x) for x in (sys.argv[2:6] or (1920, 1080, 1, 2)))
opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in sys.argv[6:]]
pos, idx, alb, v = getattr(scenes, scene)()
c = capi.Context(0)
c.upload_blue_noise(O.load_blue_noise())
for k, val in opts: c.set_option(k, val)
c.upload_mesh(pos, idx, alb); c.build()
cam = host.make_camera(w, h, v["position"], v["yaw_deg"], v["pitch_deg"])
c.atmosphere(host.atmosphere_earth())
c.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
AMD = (16.0, 2.0, 1.0, 0.18, 0.18)
def frame(i):
    pc, sc = host.camera_constants(cam, cam, i)
    c.primary_rays(w, h, pc); c.secondary_rays(sc, spp, bounces, 0); c.tonemap("amd", 1.0, AMD, capi.BUF_ACCUM)
for i in range(5): frame(i + 1)
c.sync()
for timing in (1, 0):
    c.set_option("trace_timing", timing)
    for n in (1, 50):
        c.sync(); c.stats_reset(); c.sync()
        t0 = time.perf_counter()
        for i in range(n): frame(i + 1)
        t1 = time.perf_counter()
        c.sync()
        t2 = time.perf_counter()
        st = c.stats()
        print(f"trace_timing={timing} frames={n}: host issue {1e3*(t1-t0)/n:.3f} ms/frame, issue+drain {1e3*(t2-t0)/n:.3f} ms/frame, launches/frame {st.kernel_launches/n:.0f}")
c.close()
