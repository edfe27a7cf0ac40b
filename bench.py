#!/usr/bin/env python
"""bench.py -- Mrays/s (primary + secondary) of the MinoteRT ray-tracing hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                      # our arm, one GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # our arm, N GPUs
    python bench.py --impl reference ...                                 # CPU arm on the box's host cores

N = 1.  A step is one frame of BASELINE.json configs[1]: the ~260k-triangle procedural scene at 1920x1080, 1 spp,
2 bounces, blue-noise-rotated sampling: primaryRays -> secondaryRays -> tonemap, the reference's per-frame call order
(src/gfx/renderer.ixx:56-62).  Rays are counted as the reference's structure implies: pixels x 1 primary + every
secondary ray for which a traversal was issued.
  value     device-timed (CUDA events on the context's stream), scene/BVH/LUTs resident in HBM, one frame at a time
            (the pass the per-kernel times and the traversal kernel's roofline are measured in); L2 flushed between steps
  pipelined the same steps device-timed with 3 frames in flight (frame contexts sharing one BVH)
  e2e       wall clock through the host modules' Renderer::draw(camera) (C++20 modules -> C ABI) with the reference's
            3 frames in flight, camera PODs from host memory in, RGBA8 framebuffer read back into pinned host memory
  configs   the same {value, ms_per_step, e2e, roofline} block for the other configurations the targets are stated on:
            the 1 M-triangle target scene (primary + 1 bounce, 1080p), BASELINE config 3 (10.4 M triangles, 4K,
            4 spp x 3 bounces), config 1 (Cornell-class, 512^2), the reference's own sphere frame (960x540, 8x8; CPU
            side = oracle/_ref, the reference's shaders compiled as C++), config 4 at N = 1 (the strong-scaling base), and
            config 5 (1 M animated triangles: upload + refit every frame, rebuild every 10th; end to end only)
N > 1.  BASELINE config 4: progressive 4K accumulation, tile-partitioned over the N GPUs (interleaved 8-row slabs,
replicated BVH), the RGBA8 framebuffer gathered onto rank 0 with NCCL every displayed frame -- all through the C ABI's
mrt_group_* (NCCL is called from libminotert.so; torch.distributed only hands round the ncclUniqueId and the timings).
A step is one displayed frame: 8 spp x 3 bounces at 3840x2160 on the 10.4 M-triangle scene, accumulated onto the previous
steps (8 steps = 64 spp).  Total work is fixed as N grows: STRONG scaling; `configs[tiles_4k_progressive]` of the N = 1
line is its single-GPU base.
"""
import argparse
import ctypes as C
import json
import os
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
    "tiles_4k_progressive": ("scene_10m", 3840, 2160, 8, 3), # configs[3]: one displayed frame of the 64-spp accumulation
    "spheres_960x540": (None, 960, 540, 8, 8),               # the reference's own frame (scene.glsl, main.cpp:24)
}
EXTRA_CONFIGS = ["scene_1m_1080p", "scene_10m_4k", "tiles_4k_progressive", "cornell_512", "spheres_960x540", "animated_1m_1080p"]
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
AMD = (16.0, 2.0, 1.0, 0.18, 0.18)
METRIC = "Mrays/s (primary+secondary)"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_counters(workload):
    """Counters of the dominant kernel captured under ncu for this workload (profiles/kernel_counters.json, committed with
    the ncu summaries they come from): DRAM bytes per launch, L1/L2 hit rates, lanes per warp, issue-slot utilisation."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def workload_config(wl, triangles, n_gpus):
    """The workload-defining part of the JSON line; both arms print exactly this dict."""
    gen, w, h, spp, bounces = WORKLOADS[wl]
    cfg = {"workload": wl, "scene": gen or "reference spheres (src/gpu/scene.glsl)", "triangles": triangles,
           "resolution": [w, h], "spp": spp, "bounces": bounces,
           "sampling": "PCG + blue-noise Cranley-Patterson rotation", "tonemap": "amd", "denoise": "none",
           "l2": "flushed between steps (256 MiB memset)" if wl != "tiles_4k_progressive" else "working set (BVH 0.65 GB + 4K frame buffers) larger than L2",
           "partition": "whole image per frame" if wl != "tiles_4k_progressive" else f"interleaved 8-row slabs over {n_gpus} rank(s), progressive accumulation, framebuffer gathered on rank 0 every step"}
    return cfg


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
        return self

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML, 2 ms period, timed regions only"}


_SCENES = {}


def make_scene(name):
    from minotert_b200 import scenes
    if name not in _SCENES:
        _SCENES.clear()  # one big scene in host memory at a time
        _SCENES[name] = getattr(scenes, name)()
    return _SCENES[name]


def blue_noise():
    from PIL import Image
    return np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA"), np.uint8))


# ----------------------------------------------------------------------------- CPU arm

def tests_path():
    p = os.path.join(ROOT, "tests")
    if p not in sys.path:
        sys.path.insert(0, p)


def oracle_sample(workload, budget_s=15.0):
    """Times the CPU oracle (the port; all host threads) on a bounded row strip of a triangle workload."""
    tests_path()
    import oracle_lib as O
    O.lib().orc_set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
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
    return run, rows, cores, int(idx.shape[0])


def ref_spheres_sample():
    """The reference's OWN shaders (oracle/_ref: src/gpu/*.comp compiled as C++) on the reference's own frame, all host
    threads: sky view LUT + primaryRay.comp + secondaryRays.comp (8 spp x 8 bounces) + tonemap/amd.comp per step."""
    tests_path()
    import oracle_lib as O
    import ref_lib as R
    R.lib().ref_set_num_threads(os.cpu_count() or 1)
    O.lib().orc_set_num_threads(os.cpu_count() or 1)
    _, w, h, spp, bounces = WORKLOADS["spheres_960x540"]
    cam = O.default_camera(w, h)
    atmo = R.earth()
    trans, multi, view = R.sky_luts(atmo, cam.position[:])
    bn = O.load_blue_noise()
    sp = O.spheres_array()

    def run(frame):
        pc, sc = R.constants(cam, frame=frame)
        t0 = time.perf_counter()
        vis, depth, normal, motion = R.primary(w, h, pc)
        c16 = R.secondary(w, h, sc, vis, depth, normal, bn, atmo, trans, view)
        R.tonemap("amd", c16)
        dt = time.perf_counter() - t0
        # ray count of this frame: the restatement counts them (bit-identical image, tests/test_ref_pins_oracle.py); untimed
        _, _, rays = O.secondary_spheres(w, h, sc, sp, vis, depth, normal, bn, atmo, trans, view)
        return dt, w * h + rays
    return run, R.lib().ref_num_threads()


def cpu_baseline_for(workload, frame0, budget_s=10.0):
    if workload == "spheres_960x540":
        run, cores = ref_spheres_sample()
        run(1)
        dt_sum, rays_sum, n = 0.0, 0, 0
        while dt_sum < min(budget_s, 4.0) and n < 16:
            dt, rays = run(frame0 + n)
            dt_sum, rays_sum, n = dt_sum + dt, rays_sum + rays, n + 1
        return {"value": rays_sum / dt_sum / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                "sample": f"{n} full frame(s) of 960x540, 8 spp x 8 bounces, {dt_sum:.1f} s: the reference's own GLSL shaders compiled as "
                          "C++ (oracle/_ref), OpenMP over image rows"}
    run, rows, cores, _ = oracle_sample(workload, budget_s=budget_s)
    _, w, h, _, _ = WORKLOADS[workload]
    dt_sum, rays_sum, n = 0.0, 0, 0
    while dt_sum < budget_s and n < 64:
        dt, rays = run(rows, frame0 + n)
        dt_sum, rays_sum, n = dt_sum + dt, rays_sum + rays, n + 1
    return {"value": rays_sum / dt_sum / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"rows [{rows[0]},{rows[1]}) of {w}x{h}, {n} frame(s), {dt_sum:.1f} s; CPU oracle with its own "
                      "binary BVH (stands in for Mesa lavapipe, which is not installable offline; the reference has no triangle path)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    wl = args.workload or ("hall_260k_1080p" if world == 1 else "tiles_4k_progressive")
    steps, warmup = args.steps, args.warmup
    if wl == "spheres_960x540":
        run1, cores = ref_spheres_sample()
        run = lambda rows, frame: run1(frame)
        rows, tris, kind = (0, 540), 0, "reference"
    else:
        run, rows, cores, tris = oracle_sample(wl, budget_s=max(2.0, 100.0 / max(1, steps + warmup)))
        kind = "port"
    for i in range(warmup):
        run(rows, i + 1)
    times, rays_total = [], 0
    for i in range(steps):
        dt, rays = run(rows, warmup + i + 1)
        times.append(dt)
        rays_total += rays
    t_total = sum(times)
    value = rays_total / t_total / 1e6
    _, w, h, spp, bounces = WORKLOADS[wl]
    sample = f"rows [{rows[0]},{rows[1]}) of {w}x{h} ({rows[1] - rows[0]} rows) per step, {spp} spp, {bounces} bounces"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_total / steps,
            "ms_per_step_median": 1e3 * float(np.median(times)), "ms_per_step_min": 1e3 * min(times),
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, tris, world),
            "note": "CPU arm: the reference's Vulkan renderer cannot run here (no lavapipe/glslc, MSVC-only host); triangle workloads run the "
                    "CPU oracle (port of the reference's GLSL path + its own binary BVH) on all host threads, the sphere frame runs the "
                    "reference's own shaders compiled as C++ (oracle/_ref)",
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if world == 1 and not args.no_extra_configs and wl != "spheres_960x540":
        try:
            line["configs"] = [{"workload": "spheres_960x540", "config": workload_config("spheres_960x540", 0, 1),
                                "cpu_baseline": cpu_baseline_for("spheres_960x540", 1)}]
            line["configs"][0]["value"] = line["configs"][0]["cpu_baseline"]["value"]
        except Exception as e:  # oracle/_ref not built on this box
            line["configs"] = [{"workload": "spheres_960x540", "unavailable": str(e)[:200]}]
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm, one GPU: frames of a workload

def roofline_block(wl, trace_rays, trace_ms, trace_launches, nodes_per_ray, tris_per_ray, step_ms, bvh_bytes, kernel, all_rays_visits=None):
    """nodes_per_ray / tris_per_ray: visit counts of the rays THIS kernel traces (bounce rays; primary rays, traced by
    k_mesh_primary, visit fewer nodes -- the average over all rays of a frame is reported beside it as *_all_rays)."""
    peak, peak_src = measured_peak()
    bytes_per_ray = 32 + 16 + 80.0 * nodes_per_ray + 48.0 * tris_per_ray
    achieved = (trace_rays * bytes_per_ray) / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
    ncu = ncu_counters(wl) or {}
    traffic = ncu.get("dram_bytes_per_launch")
    avg_ms = trace_ms / max(1, trace_launches)
    out = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": traffic, "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray,
           "tris_per_ray": tris_per_ray, "nodes_per_ray_all_rays": (all_rays_visits or (None, None))[0],
           "tris_per_ray_all_rays": (all_rays_visits or (None, None))[1], "launches": trace_launches, "avg_launch_ms": avg_ms,
           "share_of_step": trace_ms / max(step_ms, 1e-9),
           # what the kernel really moves through DRAM (ncu, cold caches, per launch) over its live launch time: the honest
           # memory-utilisation figure; and the issue-side counters that bound it on L2-resident scenes
           "dram_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if (traffic and avg_ms > 0) else None,
           "issue_frac": ncu.get("issue_slot_utilisation"), "lanes_per_warp": ncu.get("lanes_per_warp"),
           "l1_hit_rate": ncu.get("l1_hit_rate"), "l2_hit_rate": ncu.get("l2_hit_rate"), "ncu_source": ncu.get("source"),
           "note": ("algorithmic bytes (visit counts x node/triangle sizes); the BVH of this config fits in L2, so frac > DRAM utilisation (dram_frac)"
                    if bvh_bytes < (100 << 20) else "algorithmic bytes; the BVH exceeds L2 (HBM-resident)")}
    return out


def measure_frames(args, wl, steps, warmup, local, detail):
    """Frames of a triangle workload (or the reference's sphere frame) on one GPU -> the JSON block."""
    import torch
    from minotert_b200 import capi, host
    gen, w, h, spp, bounces = WORKLOADS[wl]
    spheres = gen is None
    if args.spp and detail:
        spp = args.spp
    if args.bounces is not None and detail:
        bounces = args.bounces
    bn = blue_noise()
    in_flight = max(1, min(3, args.frames_in_flight))
    r = host.Renderer(w, h, bn, device=local, frames_in_flight=in_flight)
    try:
        r.set_option("builder", 1 if args.builder == "ploc" else 0)
        r.set_option("ploc_radius", args.ploc_radius)
        r.set_option("trace_timing", 0 if args.no_trace_timing else 1)
        for kv in args.opt:  # A/B experiments: any mrt_set_option switch
            name, value = kv.split("=")
            r.set_option(name, int(value))
        if spheres:
            tests_path()
            import oracle_lib as O
            r.set_spheres(O.REFERENCE_SPHERES)
            cam = host.default_camera(w, h)
            ntris = 0
        else:
            pos, idx, alb, view = make_scene(gen)
            r.set_mesh(pos, idx, alb)
            cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
            ntris = int(idx.shape[0])
        r.configure(samples=spp, bounces=bounces, accumulate=False, tonemap="amd", exposure=1.0)
        if args.secondary_flags and not spheres:  # A/B of the sky extensions: the same flags through Renderer::draw
            r.set_sky_extensions(bool(args.secondary_flags & capi.SECONDARY_NEE_SUN), bool(args.secondary_flags & capi.SECONDARY_SKY_AT_HIT))
        ctx = r.context()
        frame_ctxs = [r.context(i) if i else ctx for i in range(in_flight)]
        for _ in range(in_flight):
            r.draw(cam)  # builds the atmosphere LUTs + sky view, allocates every frame buffer (of every frame context)
        for c in frame_ctxs:
            c.sync()
        if not spheres:
            ctx.build()  # second, warm build: ms_build without the first-launch module loading
            for c in frame_ctxs[1:]:
                c.share_scene(ctx)  # the rebuild made the borrowing frame contexts stale
        build_stats = ctx.stats()
        stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
        npix = w * h
        src = capi.BUF_COLOR if spheres else capi.BUF_ACCUM

        def frame(c, frame_no):
            pc, sc = host.camera_constants(cam, cam, frame_no)
            c.primary_rays(w, h, pc)
            c.secondary_rays(sc, spp, bounces, args.secondary_flags)
            c.tonemap("amd", 1.0, AMD, src)

        # The sequential pass runs one frame at a time on frame context 0: its traversal grids fill every SM.
        # (With frames in flight the host modules cap them at 3 CTAs per SM so that frames co-run; restored afterwards.)
        user_ctas = [kv for kv in args.opt if kv.startswith("trace_ctas_per_sm=")]
        if in_flight > 1 and not user_ctas:
            ctx.set_option("trace_ctas_per_sm", 0)

        nodes_per_ray = tris_per_ray = 0.0
        bounce_visits = (0.0, 0.0)
        if not spheres:  # per-ray visit counts for the algorithmic-bytes figure (untimed, counted pass)
            ctx.set_option("count_visits", 1)
            frame(ctx, 1)
            st = ctx.stats()
            nodes_per_ray = st.node_visits / max(1, st.primary_rays + st.secondary_rays)
            tris_per_ray = st.tri_tests / max(1, st.primary_rays + st.secondary_rays)
            # the bounce-wave kernel's own rays (what its roofline is quoted on); primary rays visit fewer nodes
            bounce_visits = (st.secondary_node_visits / max(1, st.secondary_rays), st.secondary_tri_tests / max(1, st.secondary_rays))
            ctx.set_option("count_visits", 0)
        for i in range(warmup):
            frame(ctx, i + 1)
        ctx.sync()
        ctx.stats_reset()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sampler = ClockSampler(local).start() if detail else None
        for i in range(steps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()                        # L2 flush, outside the event pair of the step
            ev[i][0].record(stream)
            frame(ctx, warmup + i + 1)                   # no host sync inside the loop: frames are issued back to back
            ev[i][1].record(stream)
        torch.cuda.synchronize()
        st = ctx.stats()  # running totals since stats_reset: rays (device-side sum), traversal launches and their event times
        if spheres:
            rays_total = (npix + int(st.secondary_rays)) * steps   # same frame constants structure every step: count of the last frame
            # (the sphere kernels keep the secondary-ray count of the last call only; frames differ by their seed, not their structure)
        else:
            rays_total = int(st.total_rays)
        trace_ms, trace_launches = st.ms_trace, st.trace_launches
        trace_rays = rays_total - npix * steps
        primary_ms = st.ms_primary
        secondary_ms = st.ms_secondary
        launches = st.kernel_launches
        step_ms = [a.elapsed_time(b) for a, b in ev]
        dev_ms = sum(step_ms)

        # ---- frames in flight, device-timed (detail only): the same K steps issued round-robin over the frame contexts
        pipelined = None
        if in_flight > 1 and not user_ctas:
            ctx.set_option("trace_ctas_per_sm", 3)
        if detail and in_flight > 1:
            streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local)) for c in frame_ctxs]

            def prun(flush):
                for i in range(in_flight * 2):
                    frame(frame_ctxs[i % in_flight], i + 1)
                for c in frame_ctxs:
                    c.set_option("trace_timing", 0)
                    c.stats_reset()
                torch.cuda.synchronize()
                start = torch.cuda.Event(enable_timing=True)
                start.record(streams[0])
                for s_ in streams[1:]:
                    s_.wait_event(start)
                for i in range(steps):
                    k = i % in_flight
                    if flush:
                        with torch.cuda.stream(streams[k]):
                            flush_buf.zero_()
                    frame(frame_ctxs[k], warmup + i + 1)
                ends = []
                for s_ in streams:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(s_)
                    ends.append(e)
                torch.cuda.synchronize()
                ms = max(start.elapsed_time(e) for e in ends)
                rays = sum(int(c.stats().total_rays) for c in frame_ctxs)
                return rays / (ms * 1e-3) / 1e6, ms / steps

            v_flush, ms_flush = prun(True)
            v_noflush, ms_noflush = prun(False)
            for c in frame_ctxs:
                c.set_option("trace_timing", 0 if args.no_trace_timing else 1)
            pipelined = {"frames_in_flight": in_flight, "value": v_flush, "unit": "Mrays/s", "ms_per_step": ms_flush,
                         "value_without_l2_flush": v_noflush, "ms_per_step_without_l2_flush": ms_noflush,
                         "note": "same K steps round-robin over the frame contexts (one stream each, shared BVH); one start event, "
                                 "max over the streams' end events; the 256 MiB L2 flush of each step is inside the timed region"}

        # ---- the stage after the path tracer in Renderer::draw: bilateral denoiser (reference defaults), timed on its own
        den_ms = None
        if detail:
            dm = []
            for i in range(8):
                ctx.denoise_bilateral(capi.BILATERAL_DEFAULT, cam.nearPlane, i + 1)
                dm.append(ctx.stats().ms_denoise)
            den_ms = float(np.median(dm[3:]))

        # ---- e2e: Renderer::draw(camera) + framebuffer readback into pinned host memory, wall clock.
        nfb = in_flight + 1
        fbs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(nfb)]
        fb_ptrs = [C.c_void_p(t.data_ptr()) for t in fbs]
        nbytes = fbs[0].numel()
        for _ in range(2 * in_flight):
            r.draw(cam)
            r.read_framebuffer_into(fb_ptrs[0], nbytes)
        torch.cuda.synchronize()
        r.stats_reset()  # zeroes the device-side running ray totals (every frame context)
        e2e_steps = steps
        sampler2 = ClockSampler(local).start() if detail else None
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            r.draw(cam)                                      # host: camera -> constants -> sky view -> primary -> secondary -> tonemap
            r.read_framebuffer_async(fb_ptrs[i % nfb], nbytes)   # D2H of this frame on its context's copy stream
            r.wait_framebuffer(in_flight)                    # all but the newest in_flight frames have landed in host memory
        r.wait_framebuffer(0)
        e2e_s = time.perf_counter() - t0
        e2e_rays = rays_total / steps * e2e_steps if spheres else int(r.stats().total_rays)
        clocks = None
        if detail:
            c1, c2 = sampler.stop(), sampler2.stop()
            clocks = {"sm_mhz": c1["sm_mhz"], "sm_max_mhz": c1["sm_max_mhz"], "reasons": sorted(set(c1["reasons"]) | set(c2["reasons"])),
                      "samples": c1["samples"] + c2["samples"], "sm_mhz_e2e": c2["sm_mhz"], "source": c1["source"]}

        out = {"workload": wl, "value": rays_total / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": dev_ms / steps,
               "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_min": float(min(step_ms)), "steps": steps,
               "config": workload_config(wl, ntris, 1),
               "details": {"bvh_bytes": int(build_stats.bvh_bytes), "wide_nodes": int(build_stats.num_wide_nodes),
                           "bvh_build_ms": build_stats.ms_build, "builder": args.builder, "stack_overflows": int(ctx.stats().stack_overflows),
                           "sah_node_cost": build_stats.sah_node_cost, "sah_tri_cost": build_stats.sah_tri_cost,
                           "frames_in_flight_e2e": in_flight, "options": args.opt, "secondary_flags": args.secondary_flags},
               "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 44 + 324 + 272 + 36,
                       "d2h_bytes_per_step": int(nbytes), "ms_per_step": 1e3 * e2e_s / e2e_steps, "frames_in_flight": in_flight},
               "gpu_launches": int(launches),
               "kernels": {"primary_ms_per_step": primary_ms, "secondary_ms_per_step": secondary_ms,
                           "trace_ms_per_step": trace_ms / steps, "denoise_bilateral_ms": den_ms},
               "pipelined": pipelined, "clocks": clocks}
        if spheres:
            # the reference's own frame is ALU/SFU-bound (42 B/px for up to 65 rays/px): reported against issue
            # slots (ncu), not against HBM
            ncu = ncu_counters(wl) or {}
            out["roofline"] = {"bound": "issue", "kernel": "k_spheres_secondary", "achieved": None, "peak": None, "unit": "GB/s", "frac": None,
                               "traffic": ncu.get("dram_bytes_per_launch"), "issue_frac": ncu.get("issue_slot_utilisation"),
                               "note": "ALU/SFU-bound kernel: 8 samples x <= 9 vertices x 5 sphere tests in registers per pixel"}
        else:
            out["roofline"] = roofline_block(wl, trace_rays, trace_ms, trace_launches, bounce_visits[0], bounce_visits[1], dev_ms, build_stats.bvh_bytes,
                                             "k_trace (secondary-ray BVH traversal, persistent warps)", (nodes_per_ray, tris_per_ray))
        return out
    finally:
        r.close()
        import gc
        gc.collect()
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------- GPU arm: config 4, tile-partitioned

def measure_tiles(args, steps, warmup, world, rank, local):
    """BASELINE config 4 through mrt_group_*: every rank renders its slabs of each 4K frame, tonemaps them, the RGBA8
    framebuffer is gathered on rank 0 inside the timed region."""
    import torch
    import torch.distributed as dist
    from minotert_b200 import capi, host
    wl = "tiles_4k_progressive"
    gen, w, h, spp, bounces = WORKLOADS[wl]
    dev = torch.device("cuda", local)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)     # plumbing only: the 128-byte ncclUniqueId
        g = capi.Group.rank_of(local, rank, world, bytes(uid.cpu().numpy().tobytes()))
    else:
        g = capi.Group([local], transport="p2p")
    try:
        ctx = g.contexts[0]
        pos, idx, alb, view = make_scene(gen)
        cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
        ctx.upload_blue_noise(blue_noise())
        for kv in args.opt:
            name, value = kv.split("=")
            ctx.set_option(name, int(value))
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        ctx.atmosphere(host.atmosphere_earth())
        ctx.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
        g.set_tiles(8)
        build_stats = ctx.stats()
        # Frames in flight (the reference keeps 3, renderer.ixx:36): consecutive progressive frames render on their own
        # frame contexts (shared BVH) and are added to the rank's accumulator in frame order (MRT_SECONDARY_FRAME_SUM +
        # mrt_accum_commit inside mrt_group_render), so one frame's traversal drains are filled by the next frame's kernels.
        in_flight = max(1, min(8, args.tile_frames_in_flight))
        fcs = g.set_frames_in_flight(in_flight)[0]
        for fc in fcs:
            if fc is ctx:
                continue
            fc.upload_blue_noise(blue_noise())
            for kv in args.opt:
                name, value = kv.split("=")
                fc.set_option(name, int(value))
            fc.share_scene(ctx)
            fc.atmosphere(host.atmosphere_earth())
            fc.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
        every = fcs if ctx in fcs else fcs + [ctx]
        stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
        fbs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(in_flight)] if rank == 0 else None

        def frame(i, first):
            pc, sc = host.camera_constants(cam, cam, i + 1)
            g.render(w, h, pc, sc, spp, bounces, capi.SECONDARY_FRAME_SUM | (0 if first else capi.SECONDARY_ACCUMULATE))
            g.tonemap("amd", 1.0, AMD, capi.BUF_ACCUM)
            g.gather(capi.BUF_LDR, 0)

        def stats_sum():
            per = [c.stats() for c in every]
            return per, sum(int(x.total_rays) for x in per), sum(int(x.kernel_launches) for x in per)

        fcs[0].set_option("count_visits", 1)
        frame(0, True)
        g.sync()
        st = fcs[0].stats()
        nodes_per_ray = st.node_visits / max(1, st.primary_rays + st.secondary_rays)
        tris_per_ray = st.tri_tests / max(1, st.primary_rays + st.secondary_rays)
        bounce_visits = (st.secondary_node_visits / max(1, st.secondary_rays), st.secondary_tri_tests / max(1, st.secondary_rays))
        fcs[0].set_option("count_visits", 0)
        for i in range(1, max(warmup, in_flight) + 1):
            frame(i, False)
        g.sync()
        for c in every:
            c.stats_reset()
        sampler = ClockSampler(local).start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            frame(i, i == 0)
        e1.record(stream)
        ends = [e1]
        if rank == 0:  # the gathered framebuffer of the last step lands on the exchange stream
            _, _, cs = g.result()
            e2 = torch.cuda.Event(enable_timing=True)
            e2.record(torch.cuda.ExternalStream(cs, device=dev))
            ends.append(e2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dev_ms = max(e0.elapsed_time(e) for e in ends)
        per, rays, launches = stats_sum()
        overflows = sum(int(x.stack_overflows) for x in per)
        clocks = sampler.stop()
        local_rows = len(fcs[0].partition_rows(h))

        # kernel times and the roofline of the traversal kernel: one frame at a time on ONE frame context with its full
        # traversal grid (with frames in flight the launches of different frames overlap and inflate each other's events)
        kc = fcs[0]
        kc.set_option("trace_ctas_per_sm", 0)
        g.sync()
        kc.stats_reset()
        ksteps = 3
        for i in range(ksteps):
            pc, sc = host.camera_constants(cam, cam, i + 1)
            kc.primary_rays(w, h, pc)
            kc.secondary_rays(sc, spp, bounces, capi.SECONDARY_FRAME_SUM)
        kc.sync()
        st = kc.stats()
        trace_ms, trace_launches, trace_rays = st.ms_trace, st.trace_launches, int(st.total_rays) - w * local_rows * ksteps
        kc.set_option("trace_ctas_per_sm", 0 if in_flight == 1 else (8 + in_flight - 1) // in_flight)
        frame(0, True)  # restart the accumulation for the e2e loop below

        # ---- e2e: camera constants from the host in, rank 0 reads the gathered framebuffer back every step; wall clock
        g.sync()
        for c in every:
            c.stats_reset()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            frame(i, i == 0)
            if rank == 0:  # every displayed frame is read back to pinned host memory; the host stays in_flight - 1 frames ahead
                fb = fbs[i % in_flight]
                g.readback_async(C.c_void_p(fb.data_ptr()), fb.numel())
                g.readback_wait(in_flight - 1)
        g.sync()
        if world > 1:
            dist.barrier()
        e2e_s = time.perf_counter() - t0
        e2e_rays = stats_sum()[1]
        sha = None
        if rank == 0:
            import hashlib
            sha = hashlib.sha256(fb.numpy().tobytes()).hexdigest()[:16]

        t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        cnt = torch.tensor([rays, e2e_rays, launches, trace_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = t.tolist()
        rays_all, e2e_rays_all, launches_all, trace_ms_all = cnt.tolist()
        npix_local = w * local_rows
        out = {"workload": wl, "value": rays_all / (dev_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": dev_ms / steps, "steps": steps,
               "ms_per_displayed_frame": dev_ms / steps, "n_gpus": world,
               "config": workload_config(wl, int(idx.shape[0]), world),
               "details": {"bvh_bytes": int(build_stats.bvh_bytes), "bvh_build_ms": build_stats.ms_build, "slab_rows": 8,
                           "gather_bytes_per_step": w * h * 4, "gather": "ncclSend/ncclRecv (grouped) from libminotert.so + scatter kernel" if world > 1 else "local copy + scatter kernel",
                           "framebuffer_sha256_16": sha, "stack_overflows": overflows, "options": args.opt,
                           "final_spp": spp * steps, "frames_in_flight": in_flight,
                           "accumulation": "per-frame sums (MRT_SECONDARY_FRAME_SUM) committed in frame order: same bits for every N and every number of frames in flight"},
               "e2e": {"value": e2e_rays_all / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": (324 + 272) * world,
                       "d2h_bytes_per_step": w * h * 4, "ms_per_step": e2e_ms / steps},
               "gpu_launches": int(launches_all), "clocks": clocks,
               "kernels": {"primary_ms_per_step": st.ms_primary, "secondary_ms_per_step": st.ms_secondary, "trace_ms_per_step": trace_ms / ksteps,
                           "note": "one frame at a time on one frame context (untimed extra pass)"},
               "roofline": roofline_block(wl, trace_rays, trace_ms, trace_launches, bounce_visits[0], bounce_visits[1], (st.ms_primary + st.ms_secondary) * ksteps,
                                          build_stats.bvh_bytes, "k_trace (secondary-ray BVH traversal, persistent warps), rank 0's launches",
                                          (nodes_per_ray, tris_per_ray))}
        return out
    finally:
        g.close()
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def measure_animated():
    """BASELINE.json configs[4]: ~1 M animated triangles, a new vertex array every frame (H2D from pinned memory), BVH refit
    per frame and a full rebuild every 10th, 1080p, 1 spp, one bounce, scripted freecam, framebuffer read back every frame --
    end to end through Renderer::updateMesh + draw (tools/bench_animated.py).  3 frames in flight, option async_update."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_animated
    d = bench_animated.run_animated(frames=200, rebuild_every=10, frames_in_flight=3, async_update=1)
    return {"workload": "animated_1m_1080p", "value": d["Mrays_per_s_e2e"], "ms_per_step": d["ms_per_frame"], "steps": d["frames"],
            "config": {"workload": d["workload"], "mode": d["mode"], "rebuild_every": d["rebuild_every"], "fps": d["fps"]},
            "e2e": {"value": d["Mrays_per_s_e2e"], "unit": METRIC, "h2d_bytes_per_step": d["h2d_bytes_per_frame"],
                    "d2h_bytes_per_step": d["d2h_bytes_per_frame"]},
            "note": "timed on the host around the whole loop (uploads, refits / rebuilds, frames, readbacks): value == e2e"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        blk = measure_tiles(args, args.steps, args.warmup, world, rank, local)
        # the single-GPU base of this strong-scaling point, measured in the same job: rank 0 renders the same workload
        # alone (same steps, same frames in flight) while the other ranks wait
        base = None
        if rank == 0 and not args.no_single_gpu_base:
            try:
                b1 = measure_tiles(args, args.steps, args.warmup, 1, 0, local)
                base = {"n_gpus": 1, "value": b1["value"], "ms_per_step": b1["ms_per_step"], "e2e": b1["e2e"]["value"],
                        "framebuffer_sha256_16": b1["details"]["framebuffer_sha256_16"],
                        "speedup": b1["ms_per_step"] / blk["ms_per_step"],
                        "same_framebuffer": b1["details"]["framebuffer_sha256_16"] == blk["details"]["framebuffer_sha256_16"]}
            except Exception as e:
                base = {"error": str(e)[:300]}
        dist.barrier()
        if rank == 0:
            line = {"metric": METRIC, "value": blk["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic"}
            line.update({k: blk[k] for k in ("config", "details", "clocks", "e2e", "gpu_launches", "roofline", "kernels", "ms_per_displayed_frame")})
            line["cpu_baseline"] = None
            line["single_gpu_base"] = base
            line["note"] = ("strong scaling of BASELINE config 4 (fixed 4K / 8 spp per displayed frame); its single-GPU base is "
                            "`single_gpu_base` (the same workload on rank 0's GPU alone, measured in this job; also "
                            "configs[tiles_4k_progressive] of the N = 1 line), not the N = 1 line's `value` (config 2)")
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return

    wl = args.workload or "hall_260k_1080p"
    if wl == "tiles_4k_progressive":
        blk = measure_tiles(args, args.steps, args.warmup, 1, 0, local)
    else:
        blk = measure_frames(args, wl, args.steps, args.warmup, local, detail=True)
    line = {"metric": METRIC, "value": blk["value"], "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    line.update({k: v for k, v in blk.items() if k not in ("workload", "value", "unit", "ms_per_step", "steps", "n_gpus")})
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_for(wl if wl != "tiles_4k_progressive" else "scene_10m_4k", args.warmup + 1)
    else:
        line["cpu_baseline"] = None
    if not args.no_extra_configs and not args.workload:
        line["configs"] = []
        for extra in EXTRA_CONFIGS:
            try:
                if extra == "animated_1m_1080p":
                    b = measure_animated()
                elif extra == "tiles_4k_progressive":
                    b = measure_tiles(args, 8, 3, 1, 0, local)
                else:
                    b = measure_frames(args, extra, 10 if extra != "scene_10m_4k" else 5, 3, local, detail=False)
                for k in ("pipelined", "clocks", "unit"):
                    b.pop(k, None)
                if extra == "spheres_960x540" and not args.no_cpu_baseline:
                    try:
                        b["cpu_baseline"] = cpu_baseline_for(extra, 1)
                    except Exception as e:
                        b["cpu_baseline"] = {"unavailable": str(e)[:200]}
                line["configs"].append(b)
            except Exception as e:  # an extra config must never cost the headline line
                line["configs"].append({"workload": extra, "error": str(e)[:300]})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="default: BASELINE config 2 at N = 1 (plus the `configs` array), config 4 at N > 1")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the `configs` array (1 M / 10 M / Cornell / spheres / tiles)")
    ap.add_argument("--ploc-radius", type=int, default=6)
    ap.add_argument("--no-trace-timing", action="store_true", help="A/B: drop the per-launch CUDA events (roofline fields become 0)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="extra mrt_set_option switches (A/B experiments)")
    ap.add_argument("--frames-in-flight", type=int, default=3, help="frame contexts of the e2e / pipelined measurements at N = 1 (reference: 3)")
    ap.add_argument("--secondary-flags", type=int, default=0, help="extra mrt_secondary_rays flags for A/B runs (8: sun sampling, 16: sky at hit); the headline uses 0")
    ap.add_argument("--no-single-gpu-base", action="store_true", help="N > 1: skip rank 0's single-GPU run of the same workload")
    ap.add_argument("--tile-frames-in-flight", type=int, default=6, help="frame contexts per rank of the tile-partitioned progressive workload (config 4)")
    ap.add_argument("--builder", default="ploc", choices=["ploc", "lbvh"], help="binary hierarchy under the 8-wide BVH")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
