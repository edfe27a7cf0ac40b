"""Parity at the sizes the numbers are quoted on (VERDICT r1 "untested sizes"): the 1 M-triangle target scene at
1080p, BASELINE config 3 (10.4 M triangles, 4K, 4 spp x 3 bounces) and config 5 (1 M animated triangles, refit per
frame).  The oracle cannot render these in seconds, so the checks are the size-independent ones:
  * closest hits (primitive id AND t, bit for bit) of sampled primary pixels and of sampled bounce rays -- origins on
    the primary hits, offset along the normal, cosine-lobe directions, run through the same persistent traversal
    kernel as the bounce waves (mrt_trace_rays -> k_trace) -- against the GPU brute force over all triangles, and of
    a subset against the CPU oracle's brute force;
  * no traversal stack overflow; frame determinism; accumulator linearity / sample counts."""
import ctypes as C

import numpy as np
import pytest

from minotert_b200 import capi, scenes
from test_gpu_spheres import as_capi, setup_sky

pytestmark = pytest.mark.gpu


def sampled_rays(oracle, gpu_ctx, pos, idx, cam, w, h, n, seed):
    """n primary rays of random pixels + n bounce rays leaving their hit points."""
    rng = np.random.default_rng(seed)
    xs, ys = rng.integers(0, w, n), rng.integers(0, h, n)
    pc, _ = oracle.constants(cam, frame=1)
    o = np.zeros((n, 3), np.float32)
    d = np.zeros((n, 3), np.float32)
    oo, dd = (C.c_float * 3)(), (C.c_float * 3)()
    for k in range(n):
        oracle.lib().orc_ray_gen(C.byref(pc.invView), C.byref(pc.invProjection), int(xs[k]), int(ys[k]), w, h, oo, dd)
        o[k], d[k] = oo[:], dd[:]
    vis = gpu_ctx.readback(capi.BUF_VISIBILITY)[ys, xs]
    t = gpu_ctx.readback(capi.BUF_HIT_T)[ys, xs]
    hit = vis != capi.MISS_ID
    tri = idx[vis[hit]]
    p0, p1, p2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    flip = np.sum(nrm * d[hit], axis=1) > 0
    nrm[flip] = -nrm[flip]
    hp = o[hit] + d[hit] * t[hit, None]
    u = rng.normal(size=nrm.shape)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    bd = nrm + u
    bd /= np.maximum(np.linalg.norm(bd, axis=1, keepdims=True), 1e-12)
    bo = (hp + nrm * 1e-6).astype(np.float32)
    return (xs, ys, o, d, vis, t), (bo, bd.astype(np.float32))


def check_closest_hits(oracle, gpu_ctx, osc, o, d, n_oracle, what):
    ids, t = gpu_ctx.trace_rays(o, d)
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    assert gpu_ctx.stats().stack_overflows == 0
    assert np.array_equal(ids, ids_bf), f"{what}: {(ids != ids_bf).sum()} of {len(o)} hit ids differ from brute force"
    assert np.array_equal(t, t_bf), f"{what}: hit distances differ from brute force"
    for k in np.linspace(0, len(o) - 1, n_oracle).astype(int):
        i, tt, _, _ = osc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids[k] and (i == oracle.NONE_ID or np.float32(tt) == t[k]), f"{what}: ray {k} differs from the oracle"
    return ids, t


@pytest.mark.parametrize("scene,size,spp,bounces,n_oracle", [("scene_1m", (1920, 1080), 1, 1, 50),
                                                             ("scene_10m", (3840, 2160), 4, 3, 50)],
                         ids=["target_1m_1080p", "config3_10m_4k"])
def test_full_size_scene_parity(gpu_ctx, oracle, sky_inputs, blue_noise, scene, size, spp, bounces, n_oracle):
    atmo = sky_inputs[0]
    pos, idx, alb, view = getattr(scenes, scene)()
    w, h = size
    cam = oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    st = gpu_ctx.stats()
    assert st.num_triangles == idx.shape[0]

    def render(frame, flags=0):
        pc, sc = oracle.constants(cam, frame=frame)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces, flags)

    render(1)
    st = gpu_ctx.stats()
    assert st.stack_overflows == 0
    assert st.primary_rays == w * h and 0 < st.secondary_rays <= w * h * spp * bounces
    acc1 = gpu_ctx.readback(capi.BUF_ACCUM).copy()
    assert np.all(acc1[..., 3] == spp) and np.isnan(acc1).any(-1).mean() < 1e-5   # (reference-faithful NaN: acos of 1 + 1 ulp, skyAccess.glsl:101)
    (xs, ys, o, d, vis, t), (bo, bd) = sampled_rays(oracle, gpu_ctx, pos, idx, cam, w, h, 3000, 4)
    osc = oracle.Scene(pos, idx, alb)
    ids, tt = check_closest_hits(oracle, gpu_ctx, osc, o, d, n_oracle, "primary rays")
    assert np.array_equal(ids, vis) and np.array_equal(tt, t), "the primary pass and mrt_trace_rays disagree"
    assert len(bo) >= 1000
    check_closest_hits(oracle, gpu_ctx, osc, bo, bd, n_oracle, "bounce rays")
    # determinism + linearity at full size
    render(1)
    assert np.array_equal(gpu_ctx.readback(capi.BUF_ACCUM), acc1, equal_nan=True)
    render(2, capi.SECONDARY_ACCUMULATE)
    both = gpu_ctx.readback(capi.BUF_ACCUM).copy()
    render(2)
    second = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.all(both[..., 3] == 2 * spp)
    if spp == 1:   # one addend per pixel and frame: the sum is exact
        assert np.array_equal(both, acc1 + second, equal_nan=True)
    else:          # ((acc1 + c0) + c1) + ... vs acc1 + ((c0 + c1) + ...): fp32 addition order, a few ulps
        ok = np.isfinite(both) & np.isfinite(second) & np.isfinite(acc1)
        assert np.allclose(both[ok], (acc1 + second)[ok], rtol=2e-6, atol=1e-6)


def test_config5_animated_1m_refit_parity(gpu_ctx, oracle):
    """BASELINE config 5: 1 M triangles displaced per frame (time = frame / 60), REFIT on 9 frames out of 10 and a
    full rebuild on the 10th; sampled closest hits == brute force over the moved triangles on every frame."""
    pos, idx, alb, view = scenes.scene_1m()
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    rng = np.random.default_rng(8)
    lo, hi = pos.min(0), pos.max(0)
    osc = None
    for frame in range(1, 11):
        moved = scenes.animate(pos, frame / 60.0)
        gpu_ctx.update_positions(moved)
        gpu_ctx.build(capi.BUILD_FULL if frame == 10 else capi.BUILD_REFIT)
        n = 3000
        o = rng.uniform(lo - 0.001, hi + 0.001, (n, 3)).astype(np.float32)
        o[:, 2] = hi[2] + rng.uniform(0.0, 0.01, n)              # above the terrain, looking down and sideways
        d = rng.normal(size=(n, 3))
        d[:, 2] = -np.abs(d[:, 2])
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        ids, t = gpu_ctx.trace_rays(o, d)
        ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
        assert gpu_ctx.stats().stack_overflows == 0
        assert (ids != capi.MISS_ID).mean() > 0.3
        assert np.array_equal(ids, ids_bf) and np.array_equal(t, t_bf), f"frame {frame}: refit BVH differs from brute force"
        if frame in (1, 10):
            osc = oracle.Scene(moved, idx, alb)
            for k in range(0, n, 150):
                i, tt, _, _ = osc.closest_hit(o[k], d[k], use_bvh=False)
                assert i == ids[k] and (i == oracle.NONE_ID or np.float32(tt) == t[k])
