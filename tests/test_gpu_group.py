"""Multi-GPU groups through the C ABI only (mrt_group_*): no torch.distributed anywhere.

Single-GPU box: N contexts of one device in a P2P-transport group (the N-rank tile logic, the gather and our scatter
kernel); >= 2 GPUs: the NCCL transport (ncclSend/ncclRecv gather, ncclReduce for sample sets).
Bar: the gathered image == the 1-context image, bit for bit (ids, fp32 accumulator, RGBA8 framebuffer)."""
import numpy as np
import pytest

from minotert_b200 import capi, scenes
from test_gpu_spheres import as_capi, setup_sky

pytestmark = pytest.mark.gpu


def same(a, b):
    """Bit-for-bit equality that lets the reference-faithful NaN pixels through: skyViewLutParamsToUv takes
    acos(dot(dir, up)) unclamped (skyAccess.glsl:101,109), so a bounce ray within ~1e-4 rad of the zenith whose dot
    product rounds to 1 + 1 ulp yields NaN -- a few pixels in 10^9 rays, identical in every partition."""
    return np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def device_count():
    import torch
    return torch.cuda.device_count()


def prepare(ctxs, oracle, atmo, cam, blue_noise, mesh, share=True):
    pos, idx, alb = mesh
    for i, c in enumerate(ctxs):
        setup_sky(c, oracle, atmo, cam.position[:])
        c.upload_blue_noise(blue_noise)
        if i == 0 or not share or c.device != ctxs[0].device:
            c.upload_mesh(pos, idx, alb)
            c.build()
        else:
            c.share_scene(ctxs[0])   # contexts of one device render one copy of the triangles + BVH


def progressive(render_frame, tonemap_gather, frames):
    for f in range(frames):
        render_frame(f + 1, capi.SECONDARY_ACCUMULATE if f else 0)
        tonemap_gather()


def single_context_image(oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames):
    ctx = capi.Context(0)
    try:
        prepare([ctx], oracle, atmo, cam, blue_noise, mesh)
        for f in range(frames):
            pc, sc = oracle.constants(cam, frame=f + 1)
            ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
            ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, capi.SECONDARY_ACCUMULATE if f else 0)
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        return ctx.readback(capi.BUF_LDR).copy(), ctx.readback(capi.BUF_ACCUM).copy(), ctx.readback(capi.BUF_VISIBILITY).copy()
    finally:
        ctx.close()


def group_image(g, oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames, slab):
    prepare(g.contexts, oracle, atmo, cam, blue_noise, mesh)
    g.set_tiles(slab)
    for f in range(frames):
        pc, sc = oracle.constants(cam, frame=f + 1)
        g.render(w, h, as_capi(pc, capi.PrimaryConstants), as_capi(sc, capi.SecondaryConstants), spp, bounces,
                 capi.SECONDARY_ACCUMULATE if f else 0)
        g.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        g.gather(capi.BUF_LDR, 0)          # one gather per displayed frame, asynchronous
    ldr = g.readback().copy()
    g.gather(capi.BUF_ACCUM, 0)
    acc = g.readback().copy()
    g.gather(capi.BUF_VISIBILITY, 0)
    vis = g.readback().copy()
    for c in g.contexts:
        assert c.stats().stack_overflows == 0
    return ldr, acc, vis


@pytest.mark.parametrize("nranks,slab", [(2, 8), (3, 5), (4, 8)])
def test_group_tiles_p2p_equals_single_context(oracle, sky_inputs, blue_noise, nranks, slab):
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 167, 101   # odd width: 2-byte row copies for R16F; height not a multiple of the slab
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    want = single_context_image(oracle, atmo, cam, blue_noise, (pos, idx, alb), w, h, 2, 2, 3)
    g = capi.Group([0] * nranks, transport="p2p")
    try:
        got = group_image(g, oracle, atmo, cam, blue_noise, (pos, idx, alb), w, h, 2, 2, 3, slab)
        g.gather(capi.BUF_DEPTH, 0)
        depth = g.readback()
        assert depth.shape == (h, w)
    finally:
        g.close()
    for a, b, name in zip(got, want, ("LDR", "accumulator", "visibility")):
        assert same(a, b), f"{name}: {nranks}-rank gathered image differs from the 1-context image"


def test_config4_4k_progressive_tiles_equals_single_context(oracle, sky_inputs, blue_noise):
    """BASELINE config 4 at full size: 3840x2160, 8 frames x 8 spp accumulated to 64 spp, tile-partitioned over
    N in {2, 4, 8} ranks (simulated as N contexts of this GPU sharing one BVH, P2P transport), framebuffer gathered
    every displayed frame: final framebuffer and fp32 accumulator == the 1-context render, bit for bit."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.hall_260k()
    w, h, spp, bounces, frames = 3840, 2160, 8, 2, 8
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    want = single_context_image(oracle, atmo, cam, blue_noise, (pos, idx, alb), w, h, spp, bounces, frames)
    assert np.all(want[1][..., 3] == spp * frames)
    assert np.isnan(want[1]).any(-1).mean() < 1e-5
    for nranks in (2, 4, 8):
        g = capi.Group([0] * nranks, transport="p2p")
        try:
            got = group_image(g, oracle, atmo, cam, blue_noise, (pos, idx, alb), w, h, spp, bounces, frames, 8)
        finally:
            g.close()
        for a, b, name in zip(got, want, ("LDR", "accumulator", "visibility")):
            assert same(a, b), f"config 4, {nranks} ranks: {name} differs from the 1-context image"


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs (NCCL refuses two ranks on one device)")
def test_group_nccl_two_gpus_tiles_and_sample_sets(oracle, sky_inputs, blue_noise):
    """Two ranks on two GPUs, driven through the C ABI alone: tile mode == 1-GPU image bit for bit; sample-set mode:
    ncclReduce of the accumulators == the sum of the two frames rendered one after the other on one GPU."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 320, 180
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    mesh = (pos, idx, alb)
    want = single_context_image(oracle, atmo, cam, blue_noise, mesh, w, h, 2, 2, 3)
    g = capi.Group([0, 1], transport="nccl")
    try:
        got = group_image(g, oracle, atmo, cam, blue_noise, mesh, w, h, 2, 2, 3, 8)
        for a, b, name in zip(got, want, ("LDR", "accumulator", "visibility")):
            assert same(a, b), f"NCCL tiles: {name} differs"
        # sample sets: rank r renders frame 1 + r of the whole image
        for c in g.contexts:
            c.set_partition(0, 1, 8)
        pc, sc = oracle.constants(cam, frame=1)
        g.render(w, h, as_capi(pc, capi.PrimaryConstants), as_capi(sc, capi.SecondaryConstants), 2, 2, 0, frame_stride=1)
        g.reduce(0)
        g.sync()
        summed = g.contexts[0].readback(capi.BUF_ACCUM).copy()
    finally:
        g.close()
    ctx = capi.Context(0)
    try:
        prepare([ctx], oracle, atmo, cam, blue_noise, mesh)
        parts = []
        for f in (1, 2):
            pc, sc = oracle.constants(cam, frame=f)
            ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
            ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 2, 2, 0)
            parts.append(ctx.readback(capi.BUF_ACCUM).copy())
    finally:
        ctx.close()
    assert same(summed, parts[0] + parts[1])   # two addends: fp32 addition is commutative, the sum is exact-order-free
    assert np.all(summed[..., 3] == 4.0)


# ---- frames in flight for progressive tile mode (MRT_SECONDARY_FRAME_SUM + mrt_accum_commit) ----

def frame_sum_single_context(oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames):
    """One context, one frame at a time, in the frame-sum semantic: every frame's samples are summed from zero into the
    per-frame buffer and the frame is then added to the accumulator -- the expected bits for every N and every K."""
    ctx = capi.Context(0)
    try:
        prepare([ctx], oracle, atmo, cam, blue_noise, mesh)
        for f in range(frames):
            pc, sc = oracle.constants(cam, frame=f + 1)
            ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
            ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, capi.SECONDARY_FRAME_SUM)
            ctx.accum_commit(None, capi.SECONDARY_ACCUMULATE if f else 0)
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        return ctx.readback(capi.BUF_LDR).copy(), ctx.readback(capi.BUF_ACCUM).copy()
    finally:
        ctx.close()


def group_frames_in_flight_image(nranks, k, oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames, slab):
    g = capi.Group([0] * nranks, transport="p2p")
    try:
        prepare(g.contexts, oracle, atmo, cam, blue_noise, mesh)
        g.set_tiles(slab)
        slots = g.set_frames_in_flight(k)
        for i, fcs in enumerate(slots):
            for fc in fcs:
                if fc is g.contexts[i]:
                    continue
                setup_sky(fc, oracle, atmo, cam.position[:])
                fc.upload_blue_noise(blue_noise)
                fc.share_scene(g.contexts[0])
        for f in range(frames):
            pc, sc = oracle.constants(cam, frame=f + 1)
            g.render(w, h, as_capi(pc, capi.PrimaryConstants), as_capi(sc, capi.SecondaryConstants), spp, bounces,
                     capi.SECONDARY_FRAME_SUM | (capi.SECONDARY_ACCUMULATE if f else 0))
            g.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
            g.gather(capi.BUF_LDR, 0)
        ldr = g.readback().copy()
        g.gather(capi.BUF_ACCUM, 0)
        acc = g.readback().copy()
        for fcs in slots:
            for fc in fcs:
                assert fc.stats().stack_overflows == 0
        return ldr, acc
    finally:
        g.close()


@pytest.mark.parametrize("nranks,k,slab", [(1, 1, 8), (1, 3, 8), (2, 2, 8), (3, 3, 5), (4, 2, 8), (2, 6, 8)])
def test_frames_in_flight_tiles_equal_the_sequential_frame_sum_render(oracle, sky_inputs, blue_noise, nranks, k, slab):
    """Progressive tile mode with K frame contexts per rank (frames overlap on the GPU) == one context rendering the
    frames one after the other, bit for bit, for every number of ranks and every K."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h, spp, bounces, frames = 167, 101, 2, 2, 7
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    mesh = (pos, idx, alb)
    want = frame_sum_single_context(oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames)
    got = group_frames_in_flight_image(nranks, k, oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames, slab)
    for a, b, name in zip(got, want, ("LDR", "accumulator")):
        assert same(a, b), f"{nranks} ranks x {k} frames in flight: {name} differs from the sequential frame-sum render"
    assert np.all(got[1][..., 3] == spp * frames)


def test_frame_sum_semantic_is_the_plain_accumulation_up_to_rounding(oracle, sky_inputs, blue_noise):
    """The frame-sum accumulator adds the same samples as the plain progressive accumulator, grouped per frame: equal up
    to fp32 summation order (a few ulp of the sum), identical sample counts, framebuffers within one code value."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h, spp, bounces, frames = 160, 90, 4, 2, 4
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    mesh = (pos, idx, alb)
    ldr_a, acc_a, _ = single_context_image(oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames)
    ldr_b, acc_b = frame_sum_single_context(oracle, atmo, cam, blue_noise, mesh, w, h, spp, bounces, frames)
    assert np.array_equal(acc_a[..., 3], acc_b[..., 3])
    ok = np.isfinite(acc_a).all(-1) & np.isfinite(acc_b).all(-1)
    assert np.allclose(acc_a[ok], acc_b[ok], rtol=1e-5, atol=1e-6)
    assert (np.abs(ldr_a.astype(int) - ldr_b.astype(int)).max(-1) <= 1).all()


def test_frame_sum_call_order_errors(oracle, sky_inputs, blue_noise):
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    cam = oracle.make_camera(64, 36, view["position"], view["yaw_deg"], view["pitch_deg"])
    ctx = capi.Context(0)
    try:
        prepare([ctx], oracle, atmo, cam, blue_noise, (pos, idx, alb))
        with pytest.raises(capi.MinoteError):   # nothing rendered with FRAME_SUM yet
            ctx.accum_commit(None, 0)
        pc, sc = oracle.constants(cam, frame=1)
        ctx.primary_rays(64, 36, as_capi(pc, capi.PrimaryConstants))
        ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 1, 1, capi.SECONDARY_FRAME_SUM)
        with pytest.raises(capi.MinoteError):   # the accumulator does not exist until the frame is committed
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        ctx.accum_commit(None, 0)
        ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        with pytest.raises(capi.MinoteError):   # a frame is committed once
            ctx.accum_commit(None, 0)
    finally:
        ctx.close()
