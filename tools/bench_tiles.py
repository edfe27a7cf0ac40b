#!/usr/bin/env python
"""BASELINE.json configs[3]: progressive 4K accumulation to 64 spp (8 frames x 8 spp, seeds (f<<1)|1),
tile-partitioned over N GPUs (interleaved 8-row slabs, replicated BVH), NCCL gather of the tonemapped
framebuffer after every displayed frame.  Strong scaling: the image is fixed, each rank renders 1/N of it.

    python tools/bench_tiles.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_tiles.py

Prints one JSON line on rank 0, including a sha256 of the final RGBA8 image: identical for every N
(per-pixel results depend only on (frameCounter, pixel)).
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="hall_260k")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--bounces", type=int, default=2)
    ap.add_argument("--slab", type=int, default=8)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from PIL import Image
    from minotert_b200 import capi, host, scenes
    from minotert_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = args.width, args.height
    pos, idx, alb, view = getattr(scenes, args.scene)()
    bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))
    ctx = capi.Context(local)
    ctx.upload_blue_noise(bn)
    ctx.upload_mesh(pos, idx, alb)
    ctx.build()
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    ctx.atmosphere(host.atmosphere_earth())
    ctx.sky_view(view["position"], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
    ctx.set_partition(rank, world, args.slab)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    amd = (16.0, 2.0, 1.0, 0.18, 0.18)

    side = torch.cuda.Stream(device=torch.device("cuda", local))
    rendered = [torch.cuda.Event(), torch.cuda.Event()]
    gathered = [torch.cuda.Event(), torch.cuda.Event()]

    def progressive():
        """8 displayed frames.  The gather of frame f runs on a side stream under the rendering of frame f+1 (the
        library double-buffers the RGBA8 framebuffer); frame f+2 waits for it before reusing the buffer."""
        full = None
        for f in range(1, args.frames + 1):
            b = f & 1
            pc, sc = host.camera_constants(cam, cam, f)
            ctx.primary_rays(w, h, pc)
            ctx.secondary_rays(sc, args.spp, args.bounces, capi.SECONDARY_ACCUMULATE if f > 1 else 0)
            if f > 2:
                stream.wait_event(gathered[b])
            ctx.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)
            ptr, nbytes = ctx.buffer(capi.BUF_LDR)
            ldr = D.device_tensor(ptr, nbytes, torch.uint8, f"cuda:{local}").view(-1, w, 4)
            if world > 1:
                rendered[b].record(stream)
                side.wait_event(rendered[b])
                with torch.cuda.stream(side):
                    full = D.gather_tiles(ldr, h, args.slab, dst=0)
                    gathered[b].record(side)
            else:
                full = ldr
        if world > 1:
            stream.wait_event(gathered[args.frames & 1])  # the last gather is part of the timed region
        return full

    progressive()  # warm-up: allocations, NCCL communicators
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    full = progressive()
    e1.record(stream)
    torch.cuda.synchronize()
    rays = int(ctx.stats().total_rays)  # device-side running sum over the timed frames
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    cnt = torch.tensor([float(rays)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        img = full.cpu().numpy()
        print(json.dumps({"workload": f"progressive {w}x{h} to {args.frames * args.spp} spp ({args.frames} frames x {args.spp} spp), "
                                      f"{args.bounces} bounces, {args.scene} ({idx.shape[0]} triangles), tile-partitioned",
                          "n_gpus": world, "ms_total": ms.item(), "ms_per_displayed_frame": ms.item() / args.frames,
                          "Mrays_per_s": cnt.item() / ms.item() / 1e3, "scaling": "strong",
                          "exchange": "NCCL gather of RGBA8 slabs to rank 0 after every frame, overlapped with the next frame" if world > 1 else "none",
                          "gather_bytes_per_frame": int(w * h * 4 * (world - 1) / world),
                          "image_sha256": hashlib.sha256(img.tobytes()).hexdigest()}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
