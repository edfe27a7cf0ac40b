#!/usr/bin/env python
"""BASELINE.json configs[4]: animated ~1M-triangle scene, per-frame BVH refit (full rebuild every K frames)
under scripted freecam motion, 1920x1080, through the host modules (Renderer::updateMesh + draw).

Prints one JSON line: ms per frame split into upload+refit / rebuild / render, and Mrays/s.
usage: python tools/bench_animated.py [--frames 40] [--rebuild-every 10] [--scene scene_1m]
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


def run_animated(frames=200, rebuild_every=10, frames_in_flight=3, async_update=1, scene="scene_1m", spp=1, bounces=1):
    """The measurement as a function (bench.py's `configs` entry animated_1m_1080p): returns the JSON line as a dict."""
    args = argparse.Namespace(frames=frames, rebuild_every=rebuild_every, frames_in_flight=frames_in_flight,
                              async_update=async_update, scene=scene, spp=spp, bounces=bounces, opt=[])
    return measure(args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200, help="the camera flies on: compare runs of the SAME length only")
    ap.add_argument("--rebuild-every", type=int, default=10)
    ap.add_argument("--scene", default="scene_1m")
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--bounces", type=int, default=1)
    ap.add_argument("--frames-in-flight", type=int, default=1, help="> 1 or --async-update: pipelined loop (async readbacks)")
    ap.add_argument("--async-update", type=int, default=0, help="mrt_set_option async_update: upload + refit without host stalls")
    ap.add_argument("--opt", action="append", default=[], help="name=value for mrt_set_option (A/B of build paths), repeatable")
    args = ap.parse_args()
    print(json.dumps(measure(args)))


def measure(args):
    import torch
    from PIL import Image
    from minotert_b200 import host, scenes
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device (no CPU fallback)")
    w, h = 1920, 1080
    pos, idx, alb, view = getattr(scenes, args.scene)()
    bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
    r = host.Renderer(w, h, bn, frames_in_flight=args.frames_in_flight)
    for o in getattr(args, "opt", []) or []:  # before the build: some of these invalidate the BVH
        name, value = o.split("=")
        r.set_option(name, int(value))
    r.set_mesh(pos, idx, alb)
    r.configure(samples=args.spp, bounces=args.bounces)
    r.set_option("async_update", args.async_update)
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    fb = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    fb_ptr = C.c_void_p(fb.data_ptr())
    # pre-generate the animated vertex arrays (host-side animation is not what is measured)
    # ... in pinned host memory, as a renderer that streams vertex data every frame would keep them
    pinned = [torch.from_numpy(np.ascontiguousarray(scenes.animate(pos, f / 60.0))).pin_memory() for f in range(8)]
    frames = [t.numpy() for t in pinned]
    r.draw(cam)
    r.read_framebuffer_into(fb_ptr, fb.numel())

    if args.frames_in_flight > 1 or args.async_update:
        return pipelined(args, r, cam, frames, idx, w, h)

    t_update = t_rebuild = t_render = 0.0
    n_refit = n_rebuild = 0
    rays = 0
    frame_time = 1.0 / 60.0
    t_all = time.perf_counter()
    for f in range(args.frames):
        # scripted freecam: hold W, drag the mouse slowly to the right (src/freecam.ixx:51-68)
        host.freecam_update(cam, frame_time, up=True, moving=True, cursor=(2.0, 0.0))
        t0 = time.perf_counter()
        full = args.rebuild_every > 0 and f % args.rebuild_every == args.rebuild_every - 1
        r.update_mesh(frames[f % len(frames)], refit=not full)   # H2D of the vertex array + refit / rebuild (synchronous)
        t1 = time.perf_counter()
        r.draw(cam)
        r.read_framebuffer_into(fb_ptr, fb.numel())
        t2 = time.perf_counter()
        st = r.stats()
        rays += st.primary_rays + st.secondary_rays
        if full:
            t_rebuild += t1 - t0
            n_rebuild += 1
        else:
            t_update += t1 - t0
            n_refit += 1
        t_render += t2 - t1
        frame_time = t2 - t0
    total = time.perf_counter() - t_all
    st = r.stats()
    out = ({
        "workload": f"animated {args.scene} ({idx.shape[0]} triangles), {w}x{h}, {args.spp} spp, {args.bounces} bounce(s), freecam",
        "mode": "synchronous: upload, refit / rebuild, draw and readback one after the other",
        "frames": args.frames, "ms_per_frame": 1e3 * total / args.frames, "fps": args.frames / total,
        "ms_upload_plus_refit": 1e3 * t_update / max(1, n_refit), "ms_upload_plus_rebuild": 1e3 * t_rebuild / max(1, n_rebuild),
        "ms_render_plus_readback": 1e3 * t_render / args.frames, "rebuild_every": args.rebuild_every,
        "Mrays_per_s_e2e": rays / total / 1e6, "h2d_bytes_per_frame": int(frames[0].nbytes) + 676,
        "d2h_bytes_per_frame": int(fb.numel()), "wide_nodes": int(st.num_wide_nodes), "stack_overflows": int(st.stack_overflows)})
    r.close()
    return out


def pipelined(args, r, cam, frames, idx, w, h):
    """The same frames with the host running ahead: asynchronous readbacks into a ring of pinned framebuffers, up to
    frames_in_flight - 1 of them pending; with --async-update the vertex upload and the refit do not stall either."""
    import torch
    from minotert_b200 import host
    K = args.frames_in_flight
    fbs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(K)]
    nwarm = 2 * K + max(0, args.rebuild_every)  # includes one rebuild: the second copy of the tree is allocated on first use
    for warm in range(nwarm):
        r.update_mesh(frames[warm % len(frames)], refit=not (args.rebuild_every > 0 and warm == nwarm - K - 1))
        r.draw(cam)
        r.read_framebuffer_async(C.c_void_p(fbs[warm % K].data_ptr()), fbs[warm % K].numel())
        r.wait_framebuffer(K - 1)
    r.wait_framebuffer(0)
    r.stats_reset()
    t_all = time.perf_counter()
    for f in range(args.frames):
        host.freecam_update(cam, 1.0 / 60.0, up=True, moving=True, cursor=(2.0, 0.0))
        full = args.rebuild_every > 0 and f % args.rebuild_every == args.rebuild_every - 1
        r.update_mesh(frames[f % len(frames)], refit=not full)
        r.draw(cam)
        r.read_framebuffer_async(C.c_void_p(fbs[f % K].data_ptr()), fbs[f % K].numel())
        r.wait_framebuffer(K - 1)
    r.wait_framebuffer(0)
    total = time.perf_counter() - t_all
    st = r.stats()
    out = ({
        "workload": f"animated {args.scene} ({idx.shape[0]} triangles), {w}x{h}, {args.spp} spp, {args.bounces} bounce(s), freecam",
        "mode": f"pipelined: {K} frame(s) in flight, async_update {args.async_update}",
        "frames": args.frames, "ms_per_frame": 1e3 * total / args.frames, "fps": args.frames / total,
        "rebuild_every": args.rebuild_every, "Mrays_per_s_e2e": st.total_rays / total / 1e6,
        "h2d_bytes_per_frame": int(frames[0].nbytes) + 676, "d2h_bytes_per_frame": int(fbs[0].numel()),
        "wide_nodes": int(st.num_wide_nodes), "stack_overflows": int(st.stack_overflows)})
    r.close()
    return out


if __name__ == "__main__":
    main()
