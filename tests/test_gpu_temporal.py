"""GPU parity of the temporal reprojection stage (minotert_b200/csrc/temporal.cu) through the
C ABI vs the CPU oracle.  The stage is tested in isolation, like the denoiser: the oracle runs on the SAME fp32
accumulator / visibility / motion images the GPU rendered (read back), its history is its own previous output.
Only + - x / floor on fp32 without contraction on both sides, so the bar is bit-exact."""
import ctypes as C

import numpy as np
import pytest

from minotert_b200 import capi, host, scenes
from test_gpu_spheres import as_capi, setup_sky

pytestmark = pytest.mark.gpu


def move(oracle, cam, k):
    """a freecam-like path: yaw a little, drift sideways and up"""
    c = oracle.Camera.from_buffer_copy(bytes(cam))
    oracle.lib().orc_camera_rotate(C.byref(c), 6.0 * k, -2.0 * k)
    d = (C.c_float * 3)(0.00004 * k, 0.0, 0.00002 * k)
    oracle.lib().orc_camera_shift(C.byref(c), d)
    return c


def render(ctx, oracle, cam, prev, f, w, h, spp, bounces):
    pc, sc = oracle.constants(cam, prev=prev, frame=f)
    ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces)


def check_against_oracle(ctx, oracle, hist, max_history, reset=False):
    acc, vis, mo = ctx.readback(capi.BUF_ACCUM), ctx.readback(capi.BUF_VISIBILITY), ctx.readback(capi.BUF_MOTION)
    ctx.temporal_accumulate(max_history, reset=reset)
    got, cnt = ctx.readback(capi.BUF_TEMPORAL), ctx.readback(capi.BUF_TEMPORAL_COUNT)
    want = oracle.temporal_accumulate(acc, vis, mo, None if reset else hist, max_history)
    assert np.array_equal(got.view(np.uint32), want[0].view(np.uint32)), f"{(got != want[0]).any(-1).sum()} pixels differ"
    assert np.array_equal(cnt, want[1])
    return want, vis


@pytest.mark.parametrize("scene", ["small_terrain", "spheres"])
def test_moving_camera_bit_exact(gpu_ctx, oracle, sky_inputs, blue_noise, scene):
    atmo = sky_inputs[0]
    if scene == "spheres":
        w, h = 240, 135
        cam0 = oracle.default_camera(w, h)
        gpu_ctx.set_spheres(oracle.REFERENCE_SPHERES)
        spp, bounces = 2, 3
    else:
        pos, idx, alb, view = scenes.small_terrain()
        w, h = 200, 120
        cam0 = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
        gpu_ctx.upload_mesh(pos, idx, alb)
        gpu_ctx.build()
        spp, bounces = 1, 2
    gpu_ctx.upload_blue_noise(blue_noise)
    with pytest.raises(capi.MinoteError):  # nothing rendered yet
        gpu_ctx.temporal_accumulate()
    hist, prev = None, cam0
    counts = []
    for f in range(1, 7):
        cam = move(oracle, cam0, f - 1)
        setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
        render(gpu_ctx, oracle, cam, prev, f, w, h, spp, bounces)
        hist, vis = check_against_oracle(gpu_ctx, oracle, hist, 8.0)
        hit = vis != capi.MISS_ID
        assert hit.mean() > 0.2
        counts.append(float(hist[1][hit].mean()))
        assert np.all(hist[1][~hit] == 1.0)
        prev = cam
    # the history survives the motion: the mean history length of the hit pixels keeps growing
    assert counts[0] == 1.0 and all(b > a for a, b in zip(counts, counts[1:])) and counts[-1] > 3.0, counts
    # tonemap straight from the temporal image == the oracle's tonemapper on it
    gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_TEMPORAL)
    ldr = gpu_ctx.readback(capi.BUF_LDR)
    o_ldr = oracle.tonemap("amd", hist[0])
    d = np.abs(ldr.astype(int) - o_ldr.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    # an explicit reset starts over; so does a new image size
    render(gpu_ctx, oracle, cam, cam, 7, w, h, spp, bounces)
    check_against_oracle(gpu_ctx, oracle, hist, 8.0, reset=True)
    assert np.all(gpu_ctx.readback(capi.BUF_TEMPORAL_COUNT) == 1.0)
    cam_small = oracle.Camera.from_buffer_copy(bytes(cam))
    cam_small.viewport[:] = [w // 2, h // 2]
    render(gpu_ctx, oracle, cam_small, cam_small, 8, w // 2, h // 2, spp, bounces)
    gpu_ctx.temporal_accumulate(8.0)
    assert np.all(gpu_ctx.readback(capi.BUF_TEMPORAL_COUNT) == 1.0)


def test_static_camera_converges_like_progressive_accumulation(gpu_ctx, oracle, sky_inputs, blue_noise):
    """With no motion and an uncapped history the stage is a running mean over the frames: after k frames it equals
    the k-frame progressive accumulator (row n7) up to fp32 rounding of the incremental mean."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.cornell()
    w, h = 128, 128
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    total = np.zeros((h, w, 3), np.float64)
    k = 6
    for f in range(1, k + 1):
        render(gpu_ctx, oracle, cam, cam, f, w, h, 2, 2)
        total += oracle.resolve(gpu_ctx.readback(capi.BUF_ACCUM))[..., :3]
        gpu_ctx.temporal_accumulate(1024.0)
    got = gpu_ctx.readback(capi.BUF_TEMPORAL)
    cnt = gpu_ctx.readback(capi.BUF_TEMPORAL_COUNT)
    hit = gpu_ctx.readback(capi.BUF_VISIBILITY) != capi.MISS_ID
    assert np.all(cnt[hit] == k) and np.all(cnt[~hit] == 1.0)
    mean = (total / k).astype(np.float32)
    assert np.allclose(got[..., :3][hit], mean[hit], rtol=2e-5, atol=1e-6)


def test_moving_camera_reduces_the_error(gpu_ctx, oracle, sky_inputs, blue_noise):
    """What the stage is for: under camera motion at 1 spp the reprojected history is closer to a 64-spp render of
    the same frame than the 1-spp frame alone.  Metric: mean absolute error on the hit pixels with radiance clipped
    to [0, 2] -- a squared error is decided by the handful of pixels whose path found the 1.2e5-nit sun disc, in the
    64-spp image as much as in the others.  Measured on the oracle: 0.0356 -> 0.0117 with ~9 frames of history."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 200, 120
    cam0 = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    prev = cam0
    nframes = 10
    for f in range(1, nframes + 1):
        cam = move(oracle, cam0, 0.5 * (f - 1))
        setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
        render(gpu_ctx, oracle, cam, prev, f, w, h, 1, 2)
        gpu_ctx.temporal_accumulate(32.0)
        prev = cam
    single = oracle.resolve(gpu_ctx.readback(capi.BUF_ACCUM))[..., :3]
    temporal = gpu_ctx.readback(capi.BUF_TEMPORAL)[..., :3].copy()
    hit = gpu_ctx.readback(capi.BUF_VISIBILITY) != capi.MISS_ID
    # ground truth for the last frame's camera: 64 spp from other seeds
    for i, f in enumerate(range(101, 109)):
        pc, sc = oracle.constants(cam, frame=f)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 8, 2, capi.SECONDARY_ACCUMULATE if i else 0)
    truth = oracle.resolve(gpu_ctx.readback(capi.BUF_ACCUM))[..., :3]

    def mae(a):
        return float(np.mean(np.abs(np.clip(a[hit], 0, 2) - np.clip(truth[hit], 0, 2))))

    assert mae(temporal) < 0.5 * mae(single), (mae(temporal), mae(single))


def test_errors_and_renderer_mode(oracle, blue_noise):
    """MRT_ERR_STATE under a partition; Renderer::draw with the temporal stage on == the same calls made by hand
    (and it keeps recording into one frame context, where the history lives, also with 3 frames in flight)."""
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 160, 96
    ctx = capi.Context(0)
    try:
        ctx.upload_blue_noise(blue_noise)
        ctx.atmosphere(host.atmosphere_earth())
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        ctx.set_partition(0, 2, 8)
        cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
        ctx.sky_view(list(cam.position), oracle.SUN_DIRECTION, oracle.SUN_ILLUMINANCE)
        pc, sc = host.camera_constants(cam, cam, 1)
        ctx.primary_rays(w, h, pc)
        ctx.secondary_rays(sc, 1, 1)
        with pytest.raises(capi.MinoteError):
            ctx.temporal_accumulate()
        with pytest.raises(capi.MinoteError):
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_TEMPORAL)
    finally:
        ctx.close()

    def cameras():
        cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
        for f in range(1, 6):
            if f > 1:
                host.load().minote_camera_rotate(C.byref(cam), 5.0, -1.0)
            yield f, host.Camera.from_buffer_copy(bytes(cam))

    r = host.Renderer(w, h, blue_noise, frames_in_flight=3)
    try:
        r.set_mesh(pos, idx, alb)
        r.configure(samples=1, bounces=2, accumulate=False, tonemap="amd", exposure=1.0)
        r.set_temporal(True, 16.0)
        auto = []
        for f, cam in cameras():
            r.draw(cam)
            auto.append(r.read_framebuffer().copy())
        cnt = r.context(-1).readback(capi.BUF_TEMPORAL_COUNT)
        assert cnt.max() == 5.0  # five frames of history in ONE frame context
    finally:
        r.close()

    ctx = capi.Context(0)
    try:
        ctx.upload_blue_noise(blue_noise)
        ctx.atmosphere(host.atmosphere_earth())
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        manual, prev = [], None
        for f, cam in cameras():
            pc, sc = host.camera_constants(cam, prev if prev is not None else cam, f)
            ctx.sky_view(list(cam.position), oracle.SUN_DIRECTION, oracle.SUN_ILLUMINANCE)
            ctx.primary_rays(w, h, pc)
            ctx.secondary_rays(sc, 1, 2)
            ctx.temporal_accumulate(16.0)
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_TEMPORAL)
            manual.append(ctx.readback(capi.BUF_LDR).copy())
            prev = cam
    finally:
        ctx.close()
    for f in range(5):
        assert np.array_equal(auto[f], manual[f]), f"frame {f + 1}"
    assert not np.array_equal(auto[0], auto[4])
