#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + secondary) of the MinoteRT hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                      # our arm
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                 # CPU arm (oracle on host cores)

A step is one frame of BASELINE.json configs[1]: the ~260k-triangle procedural scene at 1920x1080,
1 spp, 2 bounces, blue-noise-rotated sampling: primaryRays -> secondaryRays -> tonemap, the reference's
per-frame call order (src/gfx/renderer.ixx:56-62).  Rays are counted as the reference's structure
implies: pixels x 1 primary + every secondary ray for which a traversal was issued.

value : device-timed (CUDA events on the context's stream), scene/BVH/LUTs resident in HBM, one frame at a time
        (the pass the per-kernel times and the traversal kernel's roofline are measured in).
pipelined : the same steps device-timed with 3 frames in flight (frame contexts sharing one BVH).
e2e   : wall clock through the host modules' Renderer::draw(camera) (C++20 modules -> C ABI) with the reference's
        3 frames in flight, camera PODs coming from host memory and the RGBA8 framebuffer read back into pinned
        host memory every step.
N > 1 : sample-set partition with a replicated BVH: rank r renders frame (step*N + r + 1), the fp32
        accumulators are summed onto rank 0 with NCCL (reduce) and rank 0 tonemaps.  Weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene generator, width, height, spp, bounces)
    "hall_260k_1080p": ("hall_260k", 1920, 1080, 1, 2),      # BASELINE.json configs[1]
    "scene_1m_1080p": ("scene_1m", 1920, 1080, 1, 1),        # north_star target: 1M tris, primary + one bounce
    "scene_10m_4k": ("scene_10m", 3840, 2160, 4, 3),         # configs[2]
    "cornell_512": ("cornell", 512, 512, 1, 1),              # configs[0]
}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~2 ms by a background thread DURING the timed
    region (nvidia-smi -lms is too coarse for a region of a few tens of milliseconds; same counters)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_sm = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML, 2 ms period, timed region only"}


def make_scene(name):
    from minotert_b200 import scenes
    return getattr(scenes, name)()


# ----------------------------------------------------------------------------- CPU arm

def oracle_sample(workload, budget_s=15.0):
    """Times the CPU oracle (all host threads) on a bounded row strip of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    gen, w, h, spp, bounces = WORKLOADS[workload]
    pos, idx, alb, view = make_scene(gen)
    cam = O.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    atmo = O.earth()
    trans, multi, skyv = O.sky_luts(atmo, cam.position[:])
    bn = O.load_blue_noise()
    scene = O.Scene(pos, idx, alb)  # oracle's own binary BVH (setup, untimed)
    cores = O.lib().orc_num_threads()

    def run(rows, frame):
        pc, sc = O.constants(cam, frame=frame)
        t0 = time.perf_counter()
        acc, vis, rays = scene.render(w, h, pc, sc, bn, atmo, trans, skyv, spp, bounces, rows=rows)
        O.tonemap("amd", O.resolve(acc)[rows[0]:rows[1]])
        return time.perf_counter() - t0, rays[0] + rays[1]

    mid = h // 2
    dt, rays = run((mid - 8, mid + 8), 1)  # calibration strip (also warms the caches)
    nrows = int(max(16, min(h, 16 * budget_s / max(dt, 1e-6))))
    y0 = max(0, mid - nrows // 2)
    rows = (y0, min(h, y0 + nrows))
    return run, rows, cores, (w, h, spp, bounces)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    steps, warmup = args.steps, args.warmup
    run, rows, cores, (w, h, spp, bounces) = oracle_sample(wl, budget_s=max(2.0, 120.0 / max(1, steps + warmup)))
    for i in range(warmup):
        run(rows, i + 1)
    t_total, rays_total = 0.0, 0
    for i in range(steps):
        dt, rays = run(rows, warmup + i + 1)
        t_total += dt
        rays_total += rays
    value = rays_total / t_total / 1e6
    sample = f"rows [{rows[0]},{rows[1]}) of {w}x{h} ({rows[1] - rows[0]} rows) per step, {spp} spp, {bounces} bounces"
    line = {"impl": "reference", "metric": "Mrays/s (primary+secondary)", "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_total / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "resolution": [w, h], "spp": spp, "bounces": bounces,
                       "sampling": "PCG + blue-noise Cranley-Patterson rotation", "tonemap": "amd",
                       "note": "CPU oracle (port of the reference's GLSL path + binary BVH) on all host threads; the reference's "
                               "Vulkan renderer cannot run here (no lavapipe/glslc, MSVC-only host)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from minotert_b200 import capi, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    gen, w, h, spp, bounces = WORKLOADS[wl]
    if args.spp:
        spp = args.spp
    if args.bounces is not None:
        bounces = args.bounces
    pos, idx, alb, view = make_scene(gen)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from PIL import Image
    bn = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))

    # host modules: Cuda::Provider + Renderer::Provider, scene upload + BVH build (setup, untimed)
    # frames in flight: Renderer::draw rotates through that many frame contexts, as the reference does
    # (renderer.ixx:36); frame context 0 owns the scene and is the one the device-timed sequential pass runs on.
    # N > 1: the e2e loop rotates the frame contexts by hand (NCCL reduce of each frame's accumulator in between)
    in_flight = max(1, min(3, args.frames_in_flight))
    r = host.Renderer(w, h, bn, device=local, frames_in_flight=in_flight)
    r.set_option("builder", 1 if args.builder == "ploc" else 0)
    r.set_option("ploc_radius", args.ploc_radius)
    r.set_option("sort_rays", 1 if args.sort_rays else 0)
    r.set_option("trace_timing", 0 if args.no_trace_timing else 1)
    for kv in args.opt:  # A/B experiments: any mrt_set_option switch
        name, value = kv.split("=")
        r.set_option(name, int(value))
    r.set_mesh(pos, idx, alb)
    r.configure(samples=spp, bounces=bounces, accumulate=False, tonemap="amd", exposure=1.0)
    ctx = r.context()
    frame_ctxs = [r.context(i) if i else ctx for i in range(in_flight)]
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    for _ in range(in_flight):
        r.draw(cam)  # builds the atmosphere LUTs + sky view, allocates every frame buffer (of every frame context)
    for c in frame_ctxs:
        c.sync()
    ctx.build()  # second, warm build: ms_build without the first-launch module loading
    for c in frame_ctxs[1:]:
        c.share_scene(ctx)  # the rebuild made the borrowing frame contexts stale
    build_stats = ctx.stats()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    npix = w * h
    amd = (16.0, 2.0, 1.0, 0.18, 0.18)
    accum_ptr, accum_bytes = ctx.buffer(capi.BUF_ACCUM)

    class _Wrap:  # zero-copy torch view of the context's fp32 accumulator (for the NCCL reduce)
        __cuda_array_interface__ = {"shape": (npix * 4,), "typestr": "<f4", "data": (accum_ptr, False), "version": 2}
    accum_t = torch.as_tensor(_Wrap(), device=f"cuda:{local}") if world > 1 else None

    def frame(frame_no, tonemap=True):
        pc, sc = host.camera_constants(cam, cam, frame_no)
        ctx.primary_rays(w, h, pc)
        ctx.secondary_rays(sc, spp, bounces, 0)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
        if tonemap and rank == 0:
            ctx.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)

    def flush_l2():
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    # The sequential pass below runs one frame at a time on frame context 0: its traversal grids fill every SM.
    # (With frames in flight the host modules cap them at 3 CTAs per SM so that frames co-run; restored afterwards.)
    user_ctas = [kv for kv in args.opt if kv.startswith("trace_ctas_per_sm=")]
    if in_flight > 1 and not user_ctas:
        ctx.set_option("trace_ctas_per_sm", 0)

    # per-ray visit counts for the algorithmic-bytes figure (untimed, counted pass)
    ctx.set_option("count_visits", 1)
    frame(1)
    st = ctx.stats()
    nodes_per_ray = st.node_visits / max(1, st.primary_rays + st.secondary_rays)
    tris_per_ray = st.tri_tests / max(1, st.primary_rays + st.secondary_rays)
    ctx.set_option("count_visits", 0)

    for i in range(args.warmup):
        frame(i * world + rank + 1)
    ctx.sync()
    ctx.stats_reset()

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world == 1:
        for i in range(args.steps):
            flush_l2()                                   # outside the event pair of the step
            ev[i][0].record(stream)
            frame((args.warmup + i) * world + rank + 1)  # no host sync inside the loop: frames are issued back to back
            ev[i][1].record(stream)
    else:
        # Sample-set mode, software-pipelined by one step: the NCCL reduce of step i-1's accumulator runs on a side
        # stream while step i's primary pass (which does not touch the accumulator) runs on the main stream; rank 0
        # tonemaps step i-1 once its reduce has landed, then step i's secondary pass restarts the accumulator.
        # Every timed step still holds one primary pass, one secondary pass, one reduce and one tonemap; the
        # reduce is forked at the step's start event, so it never runs in the untimed L2-flush gap, and the
        # last step's reduce + tonemap are not overlapped with anything.
        side = torch.cuda.Stream(device=torch.device("cuda", local))
        red_done = torch.cuda.Event()
        for i in range(args.steps):
            flush_l2()
            ev[i][0].record(stream)
            if i > 0:
                side.wait_event(ev[i][0])
                with torch.cuda.stream(side):
                    dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
                    red_done.record(side)
            pc, sc = host.camera_constants(cam, cam, (args.warmup + i) * world + rank + 1)
            ctx.primary_rays(w, h, pc)
            if i > 0:
                stream.wait_event(red_done)
                if rank == 0:
                    ctx.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)
            ctx.secondary_rays(sc, spp, bounces, 0)
            if i == args.steps - 1:
                with torch.cuda.stream(stream):
                    dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    ctx.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)
            ev[i][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    st = ctx.stats()  # running totals since stats_reset: rays (device-side sum), traversal launches and their event times
    rays_total = int(st.total_rays)
    trace_ms, trace_launches = st.ms_trace, st.trace_launches
    trace_rays = rays_total - npix * args.steps
    primary_ms = st.ms_primary * args.steps          # last frame's primary pass (identical work every frame)
    launches = ctx.stats().kernel_launches
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- frames in flight, device-timed: the same K steps issued round-robin over the frame contexts (each its own
    # stream and frame buffers, one shared BVH), so the drain phase of one frame's persistent traversal launches is
    # filled by the next frame's kernels.  One start event (all streams idle), one end event per stream, max taken.
    # The L2 flush of every step sits on that step's stream INSIDE the timed region.
    pipelined = None
    if in_flight > 1 and not user_ctas:
        ctx.set_option("trace_ctas_per_sm", 3)
    if world == 1 and in_flight > 1:
        streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local)) for c in frame_ctxs]

        def pframe(c, frame_no):
            pc, sc = host.camera_constants(cam, cam, frame_no)
            c.primary_rays(w, h, pc)
            c.secondary_rays(sc, spp, bounces, 0)
            c.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)

        def prun(flush):
            for i in range(in_flight * 2):
                pframe(frame_ctxs[i % in_flight], i + 1)
            for c in frame_ctxs:
                c.set_option("trace_timing", 0)
                c.stats_reset()
            torch.cuda.synchronize()
            start = torch.cuda.Event(enable_timing=True)
            start.record(streams[0])
            for s_ in streams[1:]:
                s_.wait_event(start)
            for i in range(args.steps):
                k = i % in_flight
                if flush:
                    with torch.cuda.stream(streams[k]):
                        flush_buf.zero_()
                pframe(frame_ctxs[k], args.warmup + i + 1)
            ends = []
            for s_ in streams:
                e = torch.cuda.Event(enable_timing=True)
                e.record(s_)
                ends.append(e)
            torch.cuda.synchronize()
            ms = max(start.elapsed_time(e) for e in ends)
            rays = sum(int(c.stats().total_rays) for c in frame_ctxs)
            return rays / (ms * 1e-3) / 1e6, ms / args.steps

        v_flush, ms_flush = prun(True)
        v_noflush, ms_noflush = prun(False)
        for c in frame_ctxs:
            c.set_option("trace_timing", 0 if args.no_trace_timing else 1)
        pipelined = {"frames_in_flight": in_flight, "value": v_flush, "unit": "Mrays/s", "ms_per_step": ms_flush,
                     "value_without_l2_flush": v_noflush, "ms_per_step_without_l2_flush": ms_noflush,
                     "note": "same K steps round-robin over the frame contexts (one stream each, shared BVH); one start event, "
                             "max over the streams' end events; the 256 MiB L2 flush of each step is inside the timed region"}

    # ---- the stage after the path tracer in Renderer::draw: bilateral denoiser (reference defaults), timed on its own
    # (not part of a "step": the metric counts rays; BASELINE's configs do not name the denoiser)
    den_ms = None
    if world == 1:
        dm = []
        for i in range(8):
            ctx.denoise_bilateral(capi.BILATERAL_DEFAULT, cam.nearPlane, i + 1)
            dm.append(ctx.stats().ms_denoise)
        den_ms = float(np.median(dm[3:]))

    # ---- e2e: Renderer::draw(camera) + framebuffer readback into pinned host memory, wall clock.
    # One frame in flight, like the reference's swapchain (renderer.ixx:36): the D2H copy of frame i overlaps
    # the rendering of frame i+1 (double-buffered framebuffer), every frame's result still reaches the host.
    nfb = in_flight + 1
    fbs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(nfb)]
    fb = fbs[0]
    fb_ptrs = [C.c_void_p(t.data_ptr()) for t in fbs]
    for _ in range(2 * in_flight):
        r.draw(cam)
        r.read_framebuffer_into(fb_ptrs[0], fb.numel())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r.stats_reset()  # zeroes the device-side running ray totals (every frame context)
    e2e_rays, t0 = 0, time.perf_counter()
    if world > 1:  # zero-copy views of every frame context's accumulator + its stream
        def accum_view(c):
            ptr, _ = c.buffer(capi.BUF_ACCUM)

            class W:
                __cuda_array_interface__ = {"shape": (npix * 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}
            return torch.as_tensor(W(), device=f"cuda:{local}")
        fc_accum = [accum_view(c) for c in frame_ctxs]
        fc_stream = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local)) for c in frame_ctxs]
        frame0 = r.frame_count()
    for i in range(args.steps):
        if world > 1:
            k = i % in_flight
            c = frame_ctxs[k]
            pc, sc = host.camera_constants(cam, cam, (frame0 + i) * world + rank + 1)
            c.primary_rays(w, h, pc)
            c.secondary_rays(sc, spp, bounces, 0)
            with torch.cuda.stream(fc_stream[k]):
                dist.reduce(fc_accum[k], dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                c.tonemap("amd", 1.0, amd, capi.BUF_ACCUM)
                c.readback_async(capi.BUF_LDR, fb_ptrs[i % nfb], fb.numel())
                for back in range(in_flight):  # all but the newest in_flight frames have landed in host memory
                    frame_ctxs[(k - back) % in_flight].readback_wait(1)
        else:
            r.draw(cam)                                              # host: camera -> constants -> sky view -> primary -> secondary -> tonemap
            r.read_framebuffer_async(fb_ptrs[i % nfb], fb.numel())   # D2H of this frame on its context's copy stream
            r.wait_framebuffer(in_flight)                            # all but the newest in_flight frames have landed in host memory
    if world > 1 and rank != 0:
        for c in frame_ctxs:
            c.sync()
    elif world > 1:
        for c in frame_ctxs:
            c.readback_wait(0)
    else:
        r.wait_framebuffer(0)
    e2e_s = time.perf_counter() - t0
    e2e_rays = int(r.stats().total_rays)  # device-side running sum over exactly the e2e frames (all frame contexts)

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    counts = torch.tensor([rays_total, e2e_rays, launches], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = times.tolist()
    rays_all, e2e_rays_all, launches_all = counts.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_ray = 32 + 16 + 80.0 * nodes_per_ray + 48.0 * tris_per_ray
        achieved = (trace_rays * bytes_per_ray) / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "trace_traffic.json")) as f:
                traffic = json.load(f).get(wl)
        except Exception:
            pass
        line = {
            "metric": "Mrays/s (primary+secondary)", "value": rays_all / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "triangles": int(idx.shape[0]), "resolution": [w, h], "spp": spp, "bounces": bounces,
                       "sampling": "PCG + blue-noise Cranley-Patterson rotation", "tonemap": "amd", "denoise": "none (timed separately: kernels.denoise_bilateral_ms)",
                       "l2": "flushed between steps (256 MiB memset)", "bvh_bytes": int(build_stats.bvh_bytes),
                       "wide_nodes": int(build_stats.num_wide_nodes), "bvh_build_ms": build_stats.ms_build, "builder": args.builder, "sort_rays": bool(args.sort_rays),
                       "stack_overflows": int(ctx.stats().stack_overflows),
                       "sah_node_cost": build_stats.sah_node_cost, "sah_tri_cost": build_stats.sah_tri_cost,
                       "parallelism": "1 GPU" if world == 1 else f"sample sets over {world} GPUs, replicated BVH, NCCL reduce of the fp32 accumulator overlapped with the next step's primary pass"},
            "clocks": clocks,
            "e2e": {"value": e2e_rays_all / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 44 + 324 + 272 + 36, "d2h_bytes_per_step": int(fb.numel()),
                    "ms_per_step": e2e_ms / args.steps, "frames_in_flight": in_flight},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "k_trace (secondary-ray BVH traversal)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "launches": trace_launches, "avg_launch_ms": trace_ms / max(1, trace_launches),
                         "share_of_step": trace_ms / max(dev_ms if world == 1 else dev_ms, 1e-9),
                         "note": ("algorithmic bytes; the BVH of this config fits in L2, so frac > DRAM utilisation" if build_stats.bvh_bytes < (100 << 20)
                                  else "algorithmic bytes; the BVH exceeds L2 (HBM-resident)")},
            "kernels": {"primary_ms_per_step": primary_ms / args.steps, "trace_ms_per_step": trace_ms / args.steps,
                        "denoise_bilateral_ms": den_ms},
            "pipelined": pipelined,
        }
        if not args.no_cpu_baseline and world == 1:
            run, rows, cores, _ = oracle_sample(wl, budget_s=15.0)
            dt_sum, rays_sum, nframes = 0.0, 0, 0
            while dt_sum < 10.0 and nframes < 64:  # about 10 s of CPU work on this box
                dt, rays = run(rows, args.warmup + 1 + nframes)
                dt_sum += dt
                rays_sum += rays
                nframes += 1
            line["cpu_baseline"] = {"value": rays_sum / dt_sum / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "sample": f"rows [{rows[0]},{rows[1]}) of {w}x{h}, {nframes} frame(s), {dt_sum:.1f} s; CPU oracle with its own "
                                              "binary BVH (stands in for Mesa lavapipe, which is not installable offline)"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hall_260k_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ploc-radius", type=int, default=6)
    ap.add_argument("--no-trace-timing", action="store_true", help="A/B: drop the per-launch CUDA events (roofline fields become 0)")
    ap.add_argument("--sort-rays", action="store_true", help="bin each bounce's ray queue by direction octant before tracing")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="extra mrt_set_option switches (A/B experiments)")
    ap.add_argument("--frames-in-flight", type=int, default=3, help="frame contexts of the e2e / pipelined measurements at N = 1 (reference: 3)")
    ap.add_argument("--builder", default="ploc", choices=["ploc", "lbvh"], help="binary hierarchy under the 8-wide BVH")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
