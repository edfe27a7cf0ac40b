"""Known-answer tests for the oracle's bilateral denoiser (src/gpu/denoise/bilateral.comp:23-76,
src/gfx/modules/denoiser.ixx:20-97).  The reference ships no fixtures for it (parity unpinned), so the oracle is
pinned by (1) answers that follow from the formulas alone and (2) an independent numpy restatement written from
the shader text, tap by tap, in fp32."""
import ctypes as C

import numpy as np
import pytest


def f16(a):
    return np.ascontiguousarray(np.asarray(a, np.float16)).view(np.uint16)


def flat_inputs(w, h, rgb=(0.5, 0.25, 0.125), depth=0.01, normal=(0.0, 0.0, 1.0)):
    col = np.zeros((h, w, 4), np.float16)
    col[..., :3] = rgb
    col[..., 3] = 1.0
    dep = np.full((h, w), depth, np.float16)
    nor = np.zeros((h, w, 4), np.float16)
    nor[..., :3] = normal
    return col, dep, nor


def noise_of(oracle, x, y, frame):
    """bilateral.comp:71-72"""
    seed = C.c_uint32((x * 709 + y * 1153 + frame * 1361) & 0xFFFFFFFF)
    return np.float32(oracle.lib().orc_random_float(C.byref(seed))) * np.float32(0.005)


def unorm8(v):
    return int(np.rint(np.clip(np.float32(v), 0, 1) * np.float32(255)))


def test_tap_count_of_default_disc():
    """radius = round(2 * 5) = 10; column d.x holds floor(2*sqrt(100 - d.x^2)) + 1 taps (bilateral.comp:43-47)."""
    r = np.float32(10)
    n = 0
    for dx in range(-10, 11):
        pt = np.sqrt(r * r - np.float32(dx) * np.float32(dx), dtype=np.float32)
        dy = -pt
        while dy <= pt:
            n += 1
            dy = np.float32(dy + np.float32(1))
    assert n == sum(int(np.floor(2 * np.sqrt(100 - dx * dx))) + 1 for dx in range(-10, 11)) == 325


def test_flat_image_is_identity_plus_dither(oracle):
    """Every tap sees the centre's colour, so the weighted mean is that colour; what is left is the
    per-pixel dither noise*0.005 (same value on r, g, b) and the RGBA8 store."""
    w, h, frame = 40, 28, 3
    col, dep, nor = flat_inputs(w, h)
    out = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor), frame=frame)
    for (x, y) in [(0, 0), (39, 27), (17, 9), (5, 20)]:
        nz = noise_of(oracle, x, y, frame)
        want = [unorm8(np.float32(c) + nz) for c in (0.5, 0.25, 0.125)] + [255]
        assert np.all(np.abs(out[y, x].astype(int) - np.array(want)) <= 1), (x, y, out[y, x], want)
    assert np.all(out[..., 3] == 255)
    # the dither depends on the frame counter
    out2 = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor), frame=frame + 1)
    assert np.any(out2 != out)


def test_negative_depth_passes_centre_through(oracle):
    """bilateral.comp:36: centrZ < 0 returns the centre texel unfiltered."""
    w, h = 32, 24
    rng = np.random.default_rng(5)
    col = np.ones((h, w, 4), np.float16)
    col[..., :3] = rng.random((h, w, 3)) * 0.9
    dep = np.full((h, w), -0.5, np.float16)
    nor = np.zeros((h, w, 4), np.float16)
    nor[..., 2] = 1
    out = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor), frame=1)
    for (x, y) in [(3, 4), (31, 23), (0, 12)]:
        nz = noise_of(oracle, x, y, 1)
        want = [unorm8(np.float32(col[y, x, c]) + nz) for c in range(3)]
        assert np.all(np.abs(out[y, x, :3].astype(int) - np.array(want)) <= 1)


def test_edge_is_preserved_across_a_depth_and_normal_step(oracle):
    """Same-surface taps weigh exp(1 / (2 * 0.12^2)) = 1.2e15 times more than taps across a depth step
    (dN - dZ^2 clamps to 0 there): each side of the edge keeps its own colour."""
    w, h = 48, 32
    col, dep, nor = flat_inputs(w, h, rgb=(0.8, 0.1, 0.1), depth=0.01)
    col[:, w // 2:, :3] = (0.1, 0.1, 0.8)
    dep[:, w // 2:] = 0.0005      # z_view = near / depth: 0.1 vs 2.0 -> dZ = 190
    nor[:, w // 2:, :3] = (1.0, 0.0, 0.0)
    out = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor), frame=1)
    left, right = out[:, : w // 2, :3].astype(int), out[:, w // 2:, :3].astype(int)
    assert np.all(np.abs(left - np.array([204, 26, 26])) <= 2)
    assert np.all(np.abs(right - np.array([26, 26, 204])) <= 2)


def test_output_is_clamped_to_unit_range(oracle):
    """RGBA8 unorm output (denoiser.ixx:56): HDR values saturate at 255 before the tonemapper."""
    col, dep, nor = flat_inputs(24, 16, rgb=(7.0, 1.0, 0.0))
    out = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor))
    assert np.all(out[..., 0] == 255) and np.all(out[..., 1] == 255) and np.all(out[..., 2] <= 2)


def numpy_bilateral(col, dep, nor, sigma, ksigma, threshold, near):
    """Independent restatement of smartDeNoise from the shader text (fp32, LinearClamp = explicit lerp)."""
    F = np.float32
    h, w = dep.shape
    col, dep, nor = col.astype(F), dep.astype(F), nor.astype(F)
    sx, sy = F(w), F(h)
    gx, gy = np.meshgrid(np.arange(w, dtype=F), np.arange(h, dtype=F))
    uvx, uvy = (gx + F(0.5)) / sx, (gy + F(0.5)) / sy

    def tex(img, u, v):
        x, y = u * sx - F(0.5), v * sy - F(0.5)
        x0f, y0f = np.floor(x), np.floor(y)
        # the oracle's sampler rule for this stage: 8 fractional weight bits, zero-weight texels not read
        fx, fy = np.rint((x - x0f).astype(F) * F(256)) / F(256), np.rint((y - y0f).astype(F) * F(256)) / F(256)
        x0, y0 = x0f.astype(int), y0f.astype(int)
        x1, y1 = np.clip(x0 + 1, 0, w - 1), np.clip(y0 + 1, 0, h - 1)
        x0, y0 = np.clip(x0, 0, w - 1), np.clip(y0, 0, h - 1)
        if img.ndim == 3:
            fx, fy = fx[..., None], fy[..., None]
        def mix(a, b, f):
            with np.errstate(invalid="ignore"):
                return np.where(f == 0, a, np.where(f == 1, b, a * (F(1) - f) + b * f)).astype(F)
        return mix(mix(img[y0, x0], img[y0, x1], fx), mix(img[y1, x0], img[y1, x1], fx), fy)

    radius = F(np.floor(F(ksigma) * F(sigma) + F(0.5)))
    radq = radius * radius
    inv_s = F(0.5) / (F(sigma) * F(sigma))
    inv_s_pi = F(0.31830988618379067) * inv_s
    inv_t = F(0.5) / (F(threshold) * F(threshold))
    inv_t_pi = F(0.3989422804014327) / F(threshold)
    c_px, c_z, c_n = tex(col, uvx, uvy), tex(dep, uvx, uvy), tex(nor, uvx, uvy)
    zbuf = np.zeros((h, w), F)
    abuf = np.zeros((h, w, 4), F)
    dx = -radius
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        while dx <= radius:
            pt = np.sqrt(radq - dx * dx, dtype=F)
            dy = -pt
            while dy <= pt:
                blur = F(np.exp(-(dx * dx + dy * dy) * inv_s, dtype=F) * inv_s_pi)
                u, v = (uvx + dx / sx).astype(F), (uvy + dy / sy).astype(F)
                w_px, w_z, w_n = tex(col, u, v), tex(dep, u, v), tex(nor, u, v)
                dz = (F(near) / w_z - F(near) / c_z).astype(F) * F(100)
                dn = (w_n[..., 0] * c_n[..., 0] + w_n[..., 1] * c_n[..., 1] + w_n[..., 2] * c_n[..., 2]).astype(F)
                e = np.nan_to_num(dn - dz * dz, nan=0.0, posinf=1.0, neginf=0.0)
                delta = (np.exp(np.clip(e, 0, 1).astype(F) * inv_t, dtype=F) * inv_t_pi * blur).astype(F)
                zbuf += delta
                abuf += delta[..., None] * w_px
                dy = F(dy + F(1))
            dx = F(dx + F(1))
    out = abuf / zbuf[..., None]
    return np.where((c_z < 0)[..., None], c_px, out)


@pytest.mark.parametrize("params", [(5.0, 2.0, 0.12), (2.3, 1.6, 0.4)])
def test_oracle_matches_numpy_restatement(oracle, params):
    w, h, frame, near = 37, 26, 2, 0.001
    rng = np.random.default_rng(11)
    col = np.ones((h, w, 4), np.float16)
    col[..., :3] = rng.random((h, w, 3)) * 1.2
    col[5:7, 20:23, :3] = np.inf   # a sun disc: fp16 overflow of its radiance (pathtracer.ixx:139-145 stores RGBA16F)
    # two tilted planes meeting in a crease, one sky column (depth 0), noisy normals
    zv = (0.5 + 0.01 * np.arange(w)[None, :] + 0.02 * np.arange(h)[:, None]).astype(np.float32)
    zv[:, w // 2:] += 0.004 * np.arange(w - w // 2)[None, :]
    dep = (near / zv).astype(np.float16)
    dep[:, -3:] = 0
    nor = np.zeros((h, w, 4), np.float16)
    n = rng.normal(size=(h, w, 3)) * 0.15 + np.array([0, 0.3, 1.0])
    nor[..., :3] = n / np.linalg.norm(n, axis=-1, keepdims=True)
    out = oracle.denoise_bilateral(f16(col), f16(dep), f16(nor), params, near, frame)
    want = numpy_bilateral(col[..., :], dep, nor[..., :3], *params, near)
    noise = np.zeros((h, w), np.float32)
    for y in range(h):
        for x in range(w):
            noise[y, x] = noise_of(oracle, x, y, frame)
    want8 = np.rint(np.clip(np.nan_to_num(want[..., :3] + noise[..., None], nan=0.0), 0, 1) * 255).astype(int)
    diff = np.abs(out[..., :3].astype(int) - want8)
    assert diff.max() <= 1, diff.max()          # np.exp vs glibc expf may differ in the last ulp
    assert (diff > 0).mean() < 0.01


def test_tonemap_reads_the_rgba8_image(oracle):
    """Renderer_impl::draw chains denoise -> tonemap (renderer.ixx:61-62): the tonemapper samples k/255."""
    rng = np.random.default_rng(3)
    img8 = rng.integers(0, 256, (9, 13, 4), dtype=np.uint8)
    a = oracle.tonemap("amd", img8)
    b = oracle.tonemap("amd", (img8.astype(np.float32) / np.float32(255.0)))
    assert np.array_equal(a, b)
