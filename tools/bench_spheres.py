#!/usr/bin/env python
"""The reference's own configuration (second half): the 5 compiled-in spheres at 960x540,
8 spp x 8 bounces, full frame = sky view -> primary -> secondary -> AMD tonemap -> RGBA8 readback, on the GPU path
(spheres.cu) and on the CPU oracle in faithful mode (all host threads).  One JSON line.

    python tools/bench_spheres.py [--frames 50] [--no-cpu]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=50)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import oracle_lib as O
    from minotert_b200 import capi
    w, h = args.width, args.height
    cam = O.default_camera(w, h)
    atmo = O.earth()
    bn = O.load_blue_noise()

    def as_capi(x, T):
        return T.from_buffer_copy(bytes(x))

    ctx = capi.Context(0)
    ctx.upload_blue_noise(bn)
    ctx.set_spheres(O.REFERENCE_SPHERES)
    ctx.atmosphere(atmo)
    amd = (16.0, 2.0, 1.0, 0.18, 0.18)

    def frame(f):
        pc, sc = O.constants(cam, frame=f)
        ctx.sky_view(cam.position[:], O.SUN_DIRECTION, O.SUN_ILLUMINANCE)
        ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 8, 8)
        ctx.tonemap("amd", 1.0, amd, capi.BUF_COLOR)

    for f in range(1, 4):
        frame(f)
    ctx.sync()
    fb = np.empty((h, w, 4), np.uint8)
    ctx.stats_reset()
    rays = 0
    t0 = time.perf_counter()
    for f in range(args.frames):
        frame(4 + f)
        ctx.readback(capi.BUF_LDR, out=fb)
    dt = time.perf_counter() - t0
    st = ctx.stats()
    rays_per_frame = int(st.primary_rays) + int(st.secondary_rays)
    out = {"workload": f"reference scene: 5 spheres, {w}x{h}, 8 spp x 8 bounces (primaryRay.comp + secondaryRays.comp + amd.comp)",
           "frames": args.frames, "ms_per_frame": dt / args.frames * 1e3, "rays_per_frame": rays_per_frame,
           "Mrays_per_s": rays_per_frame * args.frames / dt / 1e6,
           "device_ms": {"primary": st.ms_primary, "secondary": st.ms_secondary, "tonemap": st.ms_tonemap},
           "timing": "wall clock, sky view + primary + secondary + tonemap + blocking RGBA8 readback every frame"}
    if not args.no_cpu:
        sp = O.spheres_array()
        pc, sc = O.constants(cam, frame=1)
        trans, multi, view = O.sky_luts(atmo, cam.position[:])
        t0 = time.perf_counter()
        vis, depth, normal, motion = O.primary_spheres(w, h, pc, sp)
        c16, c32, crays = O.secondary_spheres(w, h, sc, sp, vis, depth, normal, bn, atmo, trans, view)
        O.tonemap("amd", c16)
        cdt = time.perf_counter() - t0
        out["cpu_oracle"] = {"ms_per_frame": cdt * 1e3, "Mrays_per_s": (w * h + crays) / cdt / 1e6, "cores": O.lib().orc_num_threads(),
                             "kind": "port (faithful mode, fp16 G-buffer round trip); sky LUT generation not included"}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
