"""GPU parity of the bilateral denoiser (src/gpu/denoise/bilateral.comp,
src/gfx/modules/denoiser.ixx) through the C ABI vs the CPU oracle.

The filter is tested in isolation: the oracle runs on the SAME RGBA16F colour / R16F depth / RGBA16F normal
images the GPU rendered (read back), so differences of the path tracer itself (a PSNR-level tolerance) cannot
leak into this comparison.  Bar for the RGBA8 output: at most 1 code value apart on >= 99.9 % of the pixels and
never more than 2 -- the kernel evaluates exp and the depth division on the SFU (ex2.approx / rcp.approx) where
the oracle calls glibc; everything else is the same fp32 operation sequence."""
import numpy as np
import pytest

from minotert_b200 import capi, host, scenes
from test_gpu_spheres import as_capi, setup_sky
from test_gpu_mesh import camera_for

pytestmark = pytest.mark.gpu


def check_denoise(gpu_ctx, oracle, params, near, frame):
    col = gpu_ctx.readback(capi.BUF_COLOR)
    dep = gpu_ctx.readback(capi.BUF_DEPTH)
    nor = gpu_ctx.readback(capi.BUF_NORMAL)
    gpu_ctx.denoise_bilateral(params, near, frame)
    got = gpu_ctx.readback(capi.BUF_DENOISED)
    want = oracle.denoise_bilateral(col, dep, nor, params, near, frame)
    diff = np.abs(got.astype(int) - want.astype(int)).max(-1)
    assert diff.max() <= 2, f"max diff {diff.max()} at {np.argwhere(diff == diff.max())[:4]}"
    assert (diff <= 1).mean() >= 0.999, f"{(diff > 1).mean():.5f} of the pixels differ by more than 1"
    assert np.all(got[..., 3] == 255)
    return got, want, (diff > 0).mean()


def render_spheres(gpu_ctx, oracle, sky_inputs, blue_noise, w, h, spp, bounces, frame=1):
    atmo = sky_inputs[0]
    cam = oracle.default_camera(w, h)
    pc, sc = oracle.constants(cam, frame=frame)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.set_spheres(oracle.REFERENCE_SPHERES)
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces)
    return cam


def test_reference_scene_default_params(gpu_ctx, oracle, sky_inputs, blue_noise):
    """The reference's own frame: 960x540, 8 spp x 8 bounces, bilateral defaults (sigma 5, kSigma 2, threshold
    0.12: 325 taps), then the AMD tonemapper on the denoised RGBA8 image as Renderer_impl::draw chains them."""
    w, h = 960, 540
    cam = render_spheres(gpu_ctx, oracle, sky_inputs, blue_noise, w, h, 8, 8)
    got, want, frac = check_denoise(gpu_ctx, oracle, oracle.BILATERAL_DEFAULT, cam.nearPlane, 1)
    assert frac < 0.02
    # the sun disc overflows RGBA16F to +inf (1.2e5 nits); the filter spreads it as saturated white, never as NaN -> 0
    col = oracle.f16_to_f32(gpu_ctx.readback(capi.BUF_COLOR))
    assert np.isinf(col[..., :3]).any()
    assert np.array_equal(got[..., :3] == 255, want[..., :3] == 255)
    gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_DENOISED)
    ldr = gpu_ctx.readback(capi.BUF_LDR)
    o_ldr = oracle.tonemap("amd", got)   # the GPU's denoised image through the oracle's tonemapper
    d = np.abs(ldr.astype(int) - o_ldr.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    st = gpu_ctx.stats()
    assert st.ms_denoise > 0.0


@pytest.mark.parametrize("size,params", [((67, 35), (5.0, 2.0, 0.12)),      # ragged, smaller than one CTA tile row
                                          ((200, 113), (3.3, 1.7, 0.3)),     # radius round(5.61) = 6
                                          ((97, 61), (10.0, 3.0, 0.1)),      # radius 30: 2 821 taps, > 48 KB of shared memory
                                          ((70, 40), (16.0, 2.0, 0.2)),      # radius 32, the largest: the tile leaves no room for the tap weights
                                          ((33, 9), (1.0, 1.0, 1.0)),        # radius 1
                                          ((64, 64), (1.0, 0.4, 0.5))])      # radius 0: a single tap
def test_sizes_and_parameters(gpu_ctx, oracle, sky_inputs, blue_noise, size, params):
    w, h = size
    cam = render_spheres(gpu_ctx, oracle, sky_inputs, blue_noise, w, h, 2, 3, frame=7)
    check_denoise(gpu_ctx, oracle, params, cam.nearPlane, 7)


def test_triangle_scene_progressive(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Triangle path: the filter reads the RGBA16F resolve of the fp32 accumulator (row a12 format) and the
    G-buffer written by the mesh primary pass; depth steps and creases everywhere."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 320, 180
    cam = camera_for(oracle, view, w, h)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    for frame in (1, 2):
        pc, sc = oracle.constants(cam, frame=frame)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 2, 2, capi.SECONDARY_ACCUMULATE if frame > 1 else 0)
    acc = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.all(acc[..., 3] == 4.0)
    got, want, _ = check_denoise(gpu_ctx, oracle, oracle.BILATERAL_DEFAULT, cam.nearPlane, 2)
    # the filter smooths: less pixel-to-pixel variation than the 4 spp input on the geometry
    src = np.clip(oracle.f16_to_f32(gpu_ctx.readback(capi.BUF_COLOR))[..., :3], 0, 1)
    geo = gpu_ctx.readback(capi.BUF_VISIBILITY) != capi.MISS_ID
    m = geo[:, 1:] & geo[:, :-1]
    tv_src = np.abs(np.diff(src, axis=1)).sum(-1)[m].mean()
    tv_out = np.abs(np.diff(got[..., :3].astype(np.float32) / 255.0, axis=1)).sum(-1)[m].mean()
    assert tv_out < 0.8 * tv_src


def test_renderer_draw_applies_the_reference_default_chain(oracle, sky_inputs, blue_noise):
    """Renderer_impl::draw with the reference's defaults (renderer.ixx:56-62,140-141,183-187): sky -> primary ->
    secondary -> Denoiser::bilateral(defaults) -> Tonemapper::amd(exposure 1).  The C++ Renderer is used exactly
    as constructed (no configure call)."""
    w, h = 240, 135
    r = host.Renderer(w, h, blue_noise, device=0)
    try:
        r.set_spheres(oracle.REFERENCE_SPHERES)
        cam = host.default_camera(w, h)
        r.draw(cam)
        fb = r.read_framebuffer()
        ctx = r.context()
        col, dep, nor = (ctx.readback(b) for b in (capi.BUF_COLOR, capi.BUF_DEPTH, capi.BUF_NORMAL))
        den = oracle.denoise_bilateral(col, dep, nor, oracle.BILATERAL_DEFAULT, 0.001, 1)
        g_den = ctx.readback(capi.BUF_DENOISED)
        d = np.abs(g_den.astype(int) - den.astype(int)).max(-1)
        assert d.max() <= 2 and (d <= 1).mean() >= 0.999
        # one code value of the denoised image moves the sRGB-encoded output by up to 13 codes near black (slope
        # 12.92 of srgbEncode's linear toe), so the tonemap stage is checked on the GPU's own denoised image
        want = oracle.tonemap("amd", g_den)
        assert np.abs(fb.astype(int) - want.astype(int)).max() <= 1
        # with the denoiser switched off the frame is the path tracer's image through the tonemapper
        r.configure(samples=8, bounces=8, denoise="none")
        r.draw(cam)
        fb2 = r.read_framebuffer()
        assert np.any(fb2 != fb)
    finally:
        r.close()


def test_errors(gpu_ctx, oracle, sky_inputs, blue_noise):
    with pytest.raises(capi.MinoteError, match="before primary"):
        gpu_ctx.denoise_bilateral()
    cam = render_spheres(gpu_ctx, oracle, sky_inputs, blue_noise, 64, 48, 1, 1)
    with pytest.raises(capi.MinoteError, match="before mrt_denoise_bilateral"):
        gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_DENOISED)
    with pytest.raises(capi.MinoteError, match="<= 32"):
        gpu_ctx.denoise_bilateral((20.0, 2.0, 0.12))
    with pytest.raises(capi.MinoteError, match="threshold"):
        gpu_ctx.denoise_bilateral((5.0, 2.0, 0.0))
    gpu_ctx.denoise_bilateral()
    # a new secondary pass invalidates the denoised image
    _, sc = oracle.constants(cam)
    gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 1, 1)
    with pytest.raises(capi.MinoteError):
        gpu_ctx.readback(capi.BUF_DENOISED)
    # tile-partitioned contexts hold only their own rows
    gpu_ctx.set_partition(0, 2, 8)
    pc, sc = oracle.constants(cam)
    gpu_ctx.primary_rays(64, 48, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), 1, 1)
    with pytest.raises(capi.MinoteError, match="whole image"):
        gpu_ctx.denoise_bilateral()
