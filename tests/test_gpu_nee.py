"""SURVEY 8f-4 on the GPU: MRT_SECONDARY_NEE_SUN (shadow rays into the sun's disc, any-hit traversal) and
MRT_SECONDARY_SKY_AT_HIT against the oracle's contract (orc_render_tris_ext; CPU-side checks of the contract itself are in
test_oracle_nee.py).  Bar: radiance PSNR >= 50 dB at equal spp like the plain path (the RNG stream is the same bits; bounce
and sun directions differ in the last ulp through libdevice's sincos), equal sample counts, ray counts within 0.2 %."""
import numpy as np
import pytest

from minotert_b200 import capi, scenes
from test_gpu_spheres import as_capi, setup_sky
from test_gpu_mesh import camera_for

pytestmark = pytest.mark.gpu

EXT = {"nee": (capi.SECONDARY_NEE_SUN, 1), "nee+at_hit": (capi.SECONDARY_NEE_SUN | capi.SECONDARY_SKY_AT_HIT, 3),
       "at_hit": (capi.SECONDARY_SKY_AT_HIT, 2)}


def prepare(ctx, oracle, atmo, cam, blue_noise, mesh):
    setup_sky(ctx, oracle, atmo, cam.position[:])
    ctx.upload_blue_noise(blue_noise)
    ctx.upload_mesh(*mesh)
    ctx.build()


@pytest.mark.parametrize("name", ["nee", "nee+at_hit", "at_hit"])
def test_sky_extensions_match_the_oracle(gpu_ctx, oracle, sky_inputs, blue_noise, name):
    gflags, oext = EXT[name]
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h, spp, bounces = 200, 120, 4, 2
    cam = camera_for(oracle, view, w, h)
    pc, sc = oracle.constants(cam, frame=3)
    trans, multi, view_lut = oracle.sky_luts(atmo, cam.position[:])
    acc, vis, rays = oracle.Scene(pos, idx, alb).render(w, h, pc, sc, blue_noise, atmo, trans, view_lut, spp, bounces, ext=oext)
    prepare(gpu_ctx, oracle, atmo, cam, blue_noise, (pos, idx, alb))
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, gflags)
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    st = gpu_ctx.stats()
    assert st.stack_overflows == 0
    assert np.all(gacc[..., 3] == spp)
    assert np.array_equal(gpu_ctx.readback(capi.BUF_VISIBILITY), vis)
    assert abs(int(st.secondary_rays) - rays[1]) <= 2e-3 * rays[1] + 2, (st.secondary_rays, rays[1])
    if gflags & capi.SECONDARY_NEE_SUN:   # shadow rays are on top of the bounce rays
        plain = oracle.Scene(pos, idx, alb).render(w, h, pc, sc, blue_noise, atmo, trans, view_lut, spp, bounces)[2]
        assert rays[1] > 1.3 * plain[1]
    p = oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(acc)[..., :3], 16.0)
    assert p >= 50.0, f"{name}: PSNR {p:.1f} dB"
    # and the sun light is really there: the lit image is brighter than the plain one
    if gflags & capi.SECONDARY_NEE_SUN:
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, 0)
        plain_gpu = gpu_ctx.readback(capi.BUF_ACCUM)
        hit = vis != oracle.NONE_ID
        assert np.median(gacc[hit][:, :3].sum(-1)) > 1.5 * np.median(plain_gpu[hit][:, :3].sum(-1))


def test_nee_image_does_not_depend_on_partition_bands_or_kernel_option(oracle, sky_inputs, blue_noise):
    """Shadow rays ride the wavefront: tile partitions (2 ranks, gathered), pixel bands and the path-kernel option (which
    falls back to the wavefront for the sky extensions) give the single-context image bit for bit."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h, spp, bounces = 167, 101, 2, 2
    cam = camera_for(oracle, view, w, h)
    flags = capi.SECONDARY_NEE_SUN | capi.SECONDARY_SKY_AT_HIT

    def single(opts):
        ctx = capi.Context(0)
        try:
            for k, v in opts:
                ctx.set_option(k, v)
            prepare(ctx, oracle, atmo, cam, blue_noise, (pos, idx, alb))
            for f in range(2):
                pc, sc = oracle.constants(cam, frame=f + 1)
                ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
                ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, flags | (capi.SECONDARY_ACCUMULATE if f else 0))
            return ctx.readback(capi.BUF_ACCUM).copy()
        finally:
            ctx.close()

    want = single([])
    assert np.array_equal(single([("path_kernel", 1)]), want, equal_nan=True)
    assert np.array_equal(single([("bands", 2)]), want, equal_nan=True)
    g = capi.Group([0, 0], transport="p2p")
    try:
        for i, c in enumerate(g.contexts):
            setup_sky(c, oracle, atmo, cam.position[:])
            c.upload_blue_noise(blue_noise)
            if i == 0:
                c.upload_mesh(pos, idx, alb)
                c.build()
            else:
                c.share_scene(g.contexts[0])
        g.set_tiles(8)
        for f in range(2):
            pc, sc = oracle.constants(cam, frame=f + 1)
            g.render(w, h, as_capi(pc, capi.PrimaryConstants), as_capi(sc, capi.SecondaryConstants), spp, bounces,
                     flags | (capi.SECONDARY_ACCUMULATE if f else 0))
        g.gather(capi.BUF_ACCUM, 0)
        got = g.readback().copy()
    finally:
        g.close()
    assert np.array_equal(got, want, equal_nan=True)


def test_sky_extension_flags_are_for_triangle_scenes(gpu_ctx, oracle, sky_inputs, blue_noise):
    atmo = sky_inputs[0]
    cam = oracle.default_camera(64, 36)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.set_spheres(oracle.REFERENCE_SPHERES)
    pc, sc = oracle.constants(cam, frame=1)
    gpu_ctx.primary_rays(64, 36, as_capi(pc, capi.PrimaryConstants))
    with pytest.raises(capi.MinoteError):
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 1, 1, capi.SECONDARY_NEE_SUN)


def test_renderer_sky_extensions_reach_the_context(oracle, blue_noise):
    """Pathtracer::sunSampling / skyAtHit through Renderer::draw == the same flags on a bare context (RGBA8 framebuffer)."""
    from minotert_b200 import host
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 192, 108
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    r = host.Renderer(w, h, blue_noise, device=0, frames_in_flight=1)
    try:
        r.set_mesh(pos, idx, alb)
        r.configure(samples=2, bounces=2, accumulate=True)
        r.set_sky_extensions(sun_sampling=True, sky_at_hit=True)
        r.draw(cam)
        fb = r.read_framebuffer().copy()
        r.set_sky_extensions(False, False)
        r.configure(samples=2, bounces=2, accumulate=False)
        r.draw(cam)
        plain = r.read_framebuffer().copy()
    finally:
        r.close()
    assert fb.astype(int).sum() > 1.1 * plain.astype(int).sum()      # the sun lights the terrain
    ctx = capi.Context(0)
    try:
        ctx.upload_blue_noise(blue_noise)
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        ctx.atmosphere(host.atmosphere_earth())
        ctx.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
        pc, sc = host.camera_constants(cam, cam, 1)
        ctx.primary_rays(w, h, pc)
        ctx.secondary_rays(sc, 2, 2, capi.SECONDARY_NEE_SUN | capi.SECONDARY_SKY_AT_HIT)
        ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        want = ctx.readback(capi.BUF_LDR)
        assert np.array_equal(fb, want)
    finally:
        ctx.close()


def test_aerial_perspective_volume_and_render_match_the_oracle(gpu_ctx, oracle, sky_inputs, blue_noise):
    """mrt_sky_aerial_perspective vs orc_gen_aerial_perspective (<= 2 fp16 steps on every froxel channel: libdevice vs
    glibc exp/sqrt inside a 2..64-step ray march), and a frame with MRT_SECONDARY_AERIAL on a wall 30 km away vs the
    oracle's ORC_EXT_AERIAL (PSNR >= 50 dB; the plain frame differs from it by far more)."""
    from test_oracle_nee import far_wall_scene
    atmo = sky_inputs[0]
    w, h, spp = 96, 54, 2
    pos, idx, alb = far_wall_scene()
    cam = oracle.make_camera(w, h, (0.0, 0.0, 0.1), 90.0, 0.0, vfov_deg=30.0)
    pc, sc = oracle.constants(cam, frame=1)
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    vol = oracle.aerial_perspective(atmo, trans, multi, pc, cam.position[:])
    prepare(gpu_ctx, oracle, atmo, cam, blue_noise, (pos, idx, alb))
    gpc = as_capi(pc, capi.PrimaryConstants)
    with pytest.raises(capi.MinoteError):    # the flag needs the volume
        gpu_ctx.primary_rays(w, h, gpc)
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, 1, capi.SECONDARY_AERIAL)
    gpu_ctx.sky_aerial_perspective(gpc, cam.position[:])
    gvol = gpu_ctx.readback(capi.BUF_AERIAL)
    assert gvol.shape == vol.shape
    a, b = oracle.f16_to_f32(gvol).astype(np.float64), oracle.f16_to_f32(vol).astype(np.float64)
    step = np.maximum(np.abs(b), 6.1e-5) * 2.0 ** -10     # one fp16 step at the value (normal range)
    assert (np.abs(a - b) <= 2.0 * step + 1e-7).mean() >= 0.999, float((np.abs(a - b) / step).max())
    osc = oracle.Scene(pos, idx, alb)
    hazy, vis, _ = osc.render(w, h, pc, sc, blue_noise, atmo, trans, view, spp, 1, use_bvh=False, ext=oracle.EXT_AERIAL, aerial=vol)
    plain, _, _ = osc.render(w, h, pc, sc, blue_noise, atmo, trans, view, spp, 1, use_bvh=False)
    gpu_ctx.primary_rays(w, h, gpc)
    gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, 1, capi.SECONDARY_AERIAL)
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.array_equal(gpu_ctx.readback(capi.BUF_VISIBILITY), vis)
    p = oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(hazy)[..., :3], 16.0)
    assert p >= 50.0, f"PSNR {p:.1f} dB"
    assert oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(plain)[..., :3], 16.0) < 40.0
