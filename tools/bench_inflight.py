#!/usr/bin/env python
"""Frames in flight: K contexts on one GPU (each with its own stream and frame buffers), frames issued round-robin,
so the drain phase of one frame's persistent traversal launches is filled by another frame's kernels (the
reference keeps 3 frames in flight, src/gfx/renderer.ixx:36).  Wall clock over --frames frames, one JSON line per K.

    python tools/bench_inflight.py [--workload hall_260k_1080p] [--frames 60] [--contexts 1,2,3] [--share 0|1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

SUN_DIRECTION = (-0.435286462, 0.818654716, 0.374606609)  # src/gfx/modules/sky.ixx:193
SUN_ILLUMINANCE = (8.0, 8.0, 8.0)                          # src/gfx/modules/sky.ixx:194


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hall_260k_1080p")
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--contexts", default="1,2,3")
    ap.add_argument("--share", type=int, default=0, help="1: contexts 1.. borrow context 0's scene (mrt_scene_share)")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    import torch
    from PIL import Image
    from minotert_b200 import capi, host
    gen, w, h, spp, bounces = B.WORKLOADS[args.workload]
    pos, idx, alb, view = B.make_scene(gen)
    bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    amd = (16.0, 2.0, 1.0, 0.18, 0.18)
    atmo = host.atmosphere_earth()
    for K in [int(x) for x in args.contexts.split(",")]:
        ctxs = []
        for k in range(K):
            c = capi.Context(0)
            c.set_option("trace_timing", 0)
            for kv in args.opt:
                name, value = kv.split("=")
                c.set_option(name, int(value))
            c.upload_blue_noise(bn)
            c.atmosphere(atmo)
            if args.share and k > 0:
                c.share_scene(ctxs[0])
            else:
                c.upload_mesh(pos, idx, alb)
                c.build()
            c.sky_view(list(cam.position), SUN_DIRECTION, SUN_ILLUMINANCE)
            ctxs.append(c)

        def frame(f):
            c = ctxs[f % K]
            pc, sc = host.camera_constants(cam, cam, f + 1)
            c.primary_rays(w, h, pc)
            c.secondary_rays(sc, spp, bounces, 0)
            c.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)

        for f in range(3 * K):
            frame(f)
        for c in ctxs:
            c.sync()
            c.stats_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in range(args.frames):
            frame(f)
        t_issue = time.perf_counter() - t0
        for c in ctxs:
            c.sync()
        dt = time.perf_counter() - t0
        rays = sum(int(c.stats().total_rays) for c in ctxs)
        print(json.dumps({"workload": args.workload, "contexts": K, "shared_scene": bool(args.share), "frames": args.frames,
                          "ms_per_frame": dt / args.frames * 1e3, "host_issue_ms_per_frame": t_issue / args.frames * 1e3,
                          "Mrays_per_s": rays / dt / 1e6, "timing": "wall clock, no L2 flush, no readback"}), flush=True)
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
