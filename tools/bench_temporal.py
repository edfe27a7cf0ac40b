#!/usr/bin/env python
"""Temporal reprojection (temporal.cu) on BASELINE config 2 under a slowly turning camera: device time of
mrt_temporal_accumulate per 1080p frame, the history it keeps, and its algorithmic bandwidth.  One JSON line.

    python tools/bench_temporal.py [--frames 8] [--workload hall_260k_1080p]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

SUN_DIRECTION = (-0.435286462, 0.818654716, 0.374606609)  # src/gfx/modules/sky.ixx:193
SUN_ILLUMINANCE = (8.0, 8.0, 8.0)                          # src/gfx/modules/sky.ixx:194
BYTES_PER_PIXEL = 16 + 4 + 4 + 4 * 24 * 0.25 + 16 + 4 + 4  # accumulator, visibility, motion, history taps (each texel serves ~4 pixels), 3 outputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--workload", default="hall_260k_1080p")
    args = ap.parse_args()
    from PIL import Image
    from minotert_b200 import capi, host
    gen, w, h, spp, bounces = B.WORKLOADS[args.workload]
    pos, idx, alb, view = B.make_scene(gen)
    bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
    c = capi.Context(0)
    c.upload_blue_noise(bn)
    c.atmosphere(host.atmosphere_earth())
    c.upload_mesh(pos, idx, alb)
    c.build()
    prev = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    c.sky_view(list(prev.position), SUN_DIRECTION, SUN_ILLUMINANCE)
    ms = []
    for f in range(1, args.frames + 1):
        cur = host.Camera.from_buffer_copy(bytes(prev))
        host.load().minote_camera_rotate(C.byref(cur), 3.0, 0.5)
        pc, sc = host.camera_constants(cur, prev, f)
        c.primary_rays(w, h, pc)
        c.secondary_rays(sc, spp, bounces)
        c.temporal_accumulate(32.0)
        ms.append(c.stats().ms_temporal)
        prev = cur
    cnt = c.readback(capi.BUF_TEMPORAL_COUNT)
    warm = float(np.median(ms[2:])) if len(ms) > 2 else float(ms[-1])
    print(json.dumps({"workload": args.workload, "frames": args.frames, "ms_temporal": warm, "ms_all": [round(x, 4) for x in ms],
                      "mean_history_length": float(cnt.mean()), "algorithmic_bytes_per_pixel": BYTES_PER_PIXEL,
                      "algorithmic_GBps": BYTES_PER_PIXEL * w * h / (warm * 1e-3) / 1e9}))
    c.close()


if __name__ == "__main__":
    main()
