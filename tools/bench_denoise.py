#!/usr/bin/env python
"""Times the bilateral denoiser (mrt_denoise_bilateral, denoise.cu) on a rendered frame and, optionally, the CPU
oracle's restatement beside it.  Input: the reference's sphere scene (default) or BASELINE configs[1]'s
260k-triangle scene, path-traced on the GPU first.  One JSON line.

    python tools/bench_denoise.py [--scene spheres|hall] [--width 1920 --height 1080] [--frames 20] [--cpu]

Work per pixel: taps x (2 shared-memory texel reads, 7 lerps, rcp, ex2, 4 accumulations); the 18 B/px of input are
read once per CTA tile, so the kernel is reported against the issue-slot peak (ncu), not against HBM."""
import argparse
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
    ap.add_argument("--scene", default="spheres", choices=["spheres", "hall"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--sigma", type=float, default=5.0)
    ap.add_argument("--ksigma", type=float, default=2.0)
    ap.add_argument("--threshold", type=float, default=0.12)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (all host threads) on the same images")
    args = ap.parse_args()
    import oracle_lib as O
    from minotert_b200 import capi, scenes
    w, h = args.width, args.height
    atmo, bn = O.earth(), O.load_blue_noise()

    def as_capi(x, T):
        return T.from_buffer_copy(bytes(x))

    ctx = capi.Context(0)
    ctx.upload_blue_noise(bn)
    ctx.atmosphere(atmo)
    if args.scene == "spheres":
        cam = O.default_camera(w, h)
        ctx.set_spheres(O.REFERENCE_SPHERES)
        spp, bounces = 8, 8
    else:
        pos, idx, alb, view = scenes.hall_260k()
        cam = O.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        spp, bounces = 1, 2
    pc, sc = O.constants(cam, frame=1)
    ctx.sky_view(cam.position[:], O.SUN_DIRECTION, O.SUN_ILLUMINANCE)
    ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces)
    params = (args.sigma, args.ksigma, args.threshold)
    for _ in range(3):
        ctx.denoise_bilateral(params, cam.nearPlane, 1)
    ctx.sync()
    ms = []
    for f in range(args.frames):
        ctx.denoise_bilateral(params, cam.nearPlane, f + 1)
        ms.append(ctx.stats().ms_denoise)  # CUDA events around the launch; stats() synchronises
    radius = int(np.floor(args.ksigma * args.sigma + 0.5))
    taps = sum(int(np.floor(2 * np.sqrt(radius * radius - dx * dx))) + 1 for dx in range(-radius, radius + 1))
    med = float(np.median(ms))
    out = {"kernel": "k_denoise_bilateral", "scene": args.scene, "resolution": [w, h], "params": params, "taps": taps,
           "frames": args.frames, "ms_median": med, "ms_min": float(np.min(ms)),
           "Gtaps_per_s": w * h * taps / (med * 1e-3) / 1e9, "Mpixels_per_s": w * h / (med * 1e-3) / 1e6,
           "lib_dir": capi.LIB_DIR}
    if args.cpu:
        col, dep, nor = (ctx.readback(b) for b in (capi.BUF_COLOR, capi.BUF_DEPTH, capi.BUF_NORMAL))
        t0 = time.perf_counter()
        want = O.denoise_bilateral(col, dep, nor, params, cam.nearPlane, args.frames)
        dt = time.perf_counter() - t0
        got = ctx.readback(capi.BUF_DENOISED)
        d = np.abs(got.astype(int) - want.astype(int)).max(-1)
        out["cpu_oracle"] = {"ms": dt * 1e3, "cores": O.lib().orc_num_threads(), "kind": "port"}
        out["parity"] = {"max_code_diff": int(d.max()), "frac_diff_gt0": float((d > 0).mean()), "frac_diff_gt1": float((d > 1).mean())}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
