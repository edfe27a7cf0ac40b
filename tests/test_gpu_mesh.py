"""GPU parity, triangle path: BVH build + traversal + wavefront shading
through the C ABI vs the CPU oracle (brute force on small scenes, oracle BVH on larger ones)."""
import ctypes as C

import numpy as np
import pytest

from minotert_b200 import capi, scenes
from test_gpu_spheres import as_capi, setup_sky, ulp16_diff

pytestmark = pytest.mark.gpu


def camera_for(oracle, view, w, h):
    return oracle.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])


def random_rays(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


@pytest.mark.parametrize("builder", [0, 1], ids=["lbvh", "ploc"])
@pytest.mark.parametrize("maker", ["cornell", "small_terrain"])
def test_bvh_equals_brute_force_random_rays(gpu_ctx, oracle, maker, builder):
    pos, idx, alb, view = getattr(scenes, maker)()
    gpu_ctx.set_option("builder", builder)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    st = gpu_ctx.stats()
    assert st.num_triangles == idx.shape[0] and st.num_wide_nodes >= 1
    lo, hi = pos.min(0) - 0.001, pos.max(0) + 0.001
    o, d = random_rays(20000, lo, hi, 7)
    ids_bvh, t_bvh = gpu_ctx.trace_rays(o, d)
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    assert gpu_ctx.stats().stack_overflows == 0
    assert np.array_equal(ids_bvh, ids_bf), f"{(ids_bvh != ids_bf).sum()} of {len(o)} hit ids differ (GPU BVH vs GPU brute force)"
    assert np.array_equal(t_bvh, t_bf)
    # and against the CPU oracle's brute force (same fp32 test, no contraction): bit-exact
    sc = oracle.Scene(pos, idx, alb)
    for k in range(0, 20000, 97):
        i, t, _, _ = sc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids_bvh[k]
        if i != oracle.NONE_ID:
            assert np.float32(t) == t_bvh[k]


def test_axis_parallel_and_edge_rays(gpu_ctx, oracle):
    """Rays along axes, through shared edges and vertices of a tessellated quad: watertight, lowest id wins."""
    pos, idx, alb, _ = scenes.cornell()
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    sc = oracle.Scene(pos, idx, alb)
    # rays aimed exactly at mesh vertices (shared by up to 6 triangles) and edge midpoints of the floor grid
    verts = pos[:81]
    mids = 0.5 * (pos[idx[:128, 0]] + pos[idx[:128, 1]])
    targets = np.concatenate([verts, mids]).astype(np.float32)
    o = np.tile(np.array([[0.0, 0.0035, 0.1005]], np.float32), (targets.shape[0], 1))
    d = targets - o
    axis = np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1]], np.float32)
    o = np.concatenate([o, np.tile(np.array([[0.0003, 0.0041, 0.1001]], np.float32), (5, 1))])
    d = np.concatenate([d, axis])
    ids, t = gpu_ctx.trace_rays(o, d)
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    assert np.array_equal(ids, ids_bf) and np.array_equal(t, t_bf)
    assert (ids != capi.MISS_ID).all()
    for k in range(len(o)):
        i, tt, _, _ = sc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids[k] and np.float32(tt) == t[k]


def test_degenerate_and_tiny_meshes(gpu_ctx, oracle):
    # one triangle
    pos = np.array([[0, 1, 0], [1, 1, 0], [0, 1, 1]], np.float32)
    idx = np.array([[0, 1, 2]], np.uint32)
    alb = np.full((1, 3), 0.5, np.float32)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    ids, t = gpu_ctx.trace_rays([[0.2, 0, 0.2], [2, 0, 2]], [[0, 1, 0], [0, 1, 0]])
    assert ids[0] == 0 and t[0] == 1.0 and ids[1] == capi.MISS_ID
    # duplicates (equal t => lowest id), a zero-area triangle and 2..9 triangles
    for n in range(2, 10):
        p = np.concatenate([pos] * n + [np.array([[5, 5, 5], [5, 5, 5], [6, 6, 6]], np.float32)])
        i = np.arange(3 * (n + 1), dtype=np.uint32).reshape(-1, 3)
        a = np.full((n + 1, 3), 0.5, np.float32)
        gpu_ctx.upload_mesh(p, i, a)
        gpu_ctx.build()
        ids, t = gpu_ctx.trace_rays([[0.2, 0, 0.2]], [[0, 1, 0]])
        assert ids[0] == 0 and t[0] == 1.0
    # empty mesh: everything misses
    gpu_ctx.upload_mesh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), np.zeros((0, 3), np.float32))
    gpu_ctx.build()
    ids, t = gpu_ctx.trace_rays([[0, 0, 0]], [[0, 1, 0]])
    assert ids[0] == capi.MISS_ID
    with pytest.raises(capi.MinoteError, match="out of range"):
        gpu_ctx.upload_mesh(pos, np.array([[0, 1, 3]], np.uint32), alb)


def test_primary_gbuffer_cornell_config1(gpu_ctx, oracle):
    """BASELINE config 1 geometry: 512x512 primary pass, hit ids bit-exact vs brute-force oracle."""
    pos, idx, alb, view = scenes.cornell()
    w = h = 512
    cam = camera_for(oracle, view, w, h)
    pc, _ = oracle.constants(cam)
    sc = oracle.Scene(pos, idx, alb)
    vis, depth, normal, motion, t = sc.primary(w, h, pc, use_bvh=False)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    g_vis = gpu_ctx.readback(capi.BUF_VISIBILITY)
    mism = (g_vis != vis).mean()
    assert mism == 0.0, f"{mism:.2e} of primary hit ids differ"
    assert np.array_equal(gpu_ctx.readback(capi.BUF_HIT_T), t)
    assert ulp16_diff(gpu_ctx.readback(capi.BUF_DEPTH), depth).max() <= 1
    assert ulp16_diff(gpu_ctx.readback(capi.BUF_NORMAL)[..., :3], normal[..., :3]).max() <= 1


def render_both(gpu_ctx, oracle, sky_inputs, blue_noise, scene, w, h, spp, bounces, frame=1, use_bvh=True):
    atmo, trans, multi, view_lut = sky_inputs
    pos, idx, alb, view = scene
    cam = camera_for(oracle, view, w, h)
    pc, scn = oracle.constants(cam, frame=frame)
    # sky LUTs for this camera from the oracle and from the GPU
    trans, multi, view_lut = oracle.sky_luts(atmo, cam.position[:])
    osc = oracle.Scene(pos, idx, alb)
    acc, vis, rays = osc.render(w, h, pc, scn, blue_noise, atmo, trans, view_lut, spp, bounces, use_bvh=use_bvh)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), spp, bounces)
    return acc, vis, rays, pc, scn


def test_render_cornell_config1(gpu_ctx, oracle, sky_inputs, blue_noise):
    """BASELINE config 1: 512x512, 1 spp, 1 bounce; oracle = brute force."""
    w = h = 512
    acc, vis, rays, _, _ = render_both(gpu_ctx, oracle, sky_inputs, blue_noise, scenes.cornell(), w, h, 1, 1, use_bvh=False)
    g_vis = gpu_ctx.readback(capi.BUF_VISIBILITY)
    assert (g_vis != vis).mean() <= 1e-3
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    st = gpu_ctx.stats()
    assert st.stack_overflows == 0
    assert st.primary_rays == rays[0]
    assert abs(int(st.secondary_rays) - rays[1]) <= 1e-3 * rays[1] + 2
    p = oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(acc)[..., :3], 16.0)
    assert p >= 50.0, f"PSNR {p:.1f} dB"
    gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
    g_ldr = gpu_ctx.readback(capi.BUF_LDR)
    o_ldr = oracle.tonemap("amd", oracle.resolve(acc))
    assert (np.abs(g_ldr.astype(int) - o_ldr.astype(int)).max(-1) <= 1).mean() >= 0.998


def test_render_multi_sample_multi_bounce(gpu_ctx, oracle, sky_inputs, blue_noise):
    """4 spp x 3 bounces: the PCG state threads through samples exactly as secondaryRays.comp:131-132."""
    w, h = 200, 120
    acc, vis, rays, _, _ = render_both(gpu_ctx, oracle, sky_inputs, blue_noise, scenes.small_terrain(), w, h, 4, 3)
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.all(gacc[..., 3] == 4.0)
    st = gpu_ctx.stats()
    assert abs(int(st.secondary_rays) - rays[1]) <= 2e-3 * rays[1] + 2
    p = oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(acc)[..., :3], 16.0)
    assert p >= 50.0, f"PSNR {p:.1f} dB"
    # RGBA16F colour image the reference interface returns
    g16 = gpu_ctx.readback(capi.BUF_COLOR)
    assert oracle.psnr(oracle.f16_to_f32(g16)[..., :3], oracle.resolve(acc)[..., :3], 16.0) >= 50.0


def test_render_260k_config2_reduced_resolution(gpu_ctx, oracle, sky_inputs, blue_noise):
    """BASELINE config 2 scene (~260k tris), 1 spp, 2 bounces, at 480x270 so the oracle finishes in seconds."""
    w, h = 480, 270
    acc, vis, rays, _, _ = render_both(gpu_ctx, oracle, sky_inputs, blue_noise, scenes.hall_260k(), w, h, 1, 2)
    g_vis = gpu_ctx.readback(capi.BUF_VISIBILITY)
    mism = (g_vis != vis).mean()
    assert mism <= 1e-3, f"primary hit-id mismatch {mism:.2e}"
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    st = gpu_ctx.stats()
    assert st.stack_overflows == 0
    p = oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(acc)[..., :3], 16.0)
    assert p >= 50.0, f"PSNR {p:.1f} dB"


def test_ray_sort_does_not_change_the_image(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Row n5: reordering the bounce queues -- mode 1 binning by direction octant, mode 2 radix sort by (Morton cell of
    the ray origin, direction octant) -- reorders work only; every pixel owns its path, so the accumulator must be
    bit-identical with and without the sort stage."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 192, 128
    cam = camera_for(oracle, view, w, h)
    pc, scn = oracle.constants(cam, frame=5)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    gpu_ctx.set_option("path_kernel", 0)   # the sort stage sits between the waves of the wavefront
    out = []
    for sort in (0, 1, 2):
        gpu_ctx.set_option("sort_rays", sort)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 3, 3)
        out.append((gpu_ctx.readback(capi.BUF_ACCUM).copy(), int(gpu_ctx.stats().secondary_rays)))
    assert out[0][1] == out[1][1] == out[2][1] and out[0][1] > 0
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][0], out[2][0])
    # the flag on the call does the same as the option
    gpu_ctx.set_option("sort_rays", 0)
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 3, 3, capi.SECONDARY_SORT_RAYS)
    assert np.array_equal(gpu_ctx.readback(capi.BUF_ACCUM), out[0][0])


def test_fused_shade_does_not_change_the_image(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Option fused_shade (shade stage inside the traversal launch, hit record = ready flag): the accumulator is
    bit-identical to the two-kernel path, with and without the sort stage, also when toggled between frames and when
    the image size (and with it the hit-record buffer) changes."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    gpu_ctx.set_option("path_kernel", 0)   # fused_shade is an option of the wavefront
    try:
        for (w, h) in ((192, 128), (333, 211), (64, 48)):
            cam = camera_for(oracle, view, w, h)
            pc, scn = oracle.constants(cam, frame=3)
            setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
            out = {}
            for fused, flags in ((0, 0), (1, 0), (0, 0), (1, capi.SECONDARY_SORT_RAYS), (1, 0)):
                gpu_ctx.set_option("fused_shade", fused)
                gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
                gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 3, 3, flags)
                acc = gpu_ctx.readback(capi.BUF_ACCUM).copy()
                st = gpu_ctx.stats()
                assert st.stack_overflows == 0
                out.setdefault("ref", (acc, int(st.secondary_rays)))
                assert int(st.secondary_rays) == out["ref"][1]
                assert np.array_equal(acc, out["ref"][0]), f"fused={fused} flags={flags} {w}x{h}"
    finally:
        gpu_ctx.set_option("fused_shade", 0)


def test_path_kernel_equals_wavefront(oracle, sky_inputs, blue_noise):
    """Option path_kernel (the whole secondary pass as one persistent launch, a lane owns a pixel through all its
    samples and bounces) against the wavefront (trace + shade launches per wave, compacted queues): accumulator,
    ray count and framebuffer are bit-identical -- across image sizes, spp x bounces shapes incl. 0 bounces,
    progressive accumulation, a tile partition, and when toggled between frames."""
    atmo = sky_inputs[0]
    for maker, sizes in (("small_terrain", ((192, 128), (333, 211), (64, 48))), ("cornell", ((160, 160),))):
        pos, idx, alb, view = getattr(scenes, maker)()
        ctx = capi.Context(0)
        try:
            ctx.upload_blue_noise(blue_noise)
            ctx.upload_mesh(pos, idx, alb)
            ctx.build()
            for (w, h) in sizes:
                cam = camera_for(oracle, view, w, h)
                setup_sky(ctx, oracle, atmo, cam.position[:])
                for part in ((0, 1), (1, 3)):
                    ctx.set_partition(part[0], part[1], 8)
                    for spp, bounces in ((1, 1), (3, 3), (8, 2), (2, 0), (1, 8)):
                        out = []
                        for pk in (0, 1, 0, 1):
                            ctx.set_option("path_kernel", pk)
                            rays = 0
                            for frame in (3, 4):   # second frame accumulates onto the first
                                pc, scn = oracle.constants(cam, frame=frame)
                                ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
                                ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), spp, bounces,
                                                   capi.SECONDARY_ACCUMULATE if frame == 4 else 0)
                                st = ctx.stats()
                                assert st.stack_overflows == 0
                                rays += int(st.secondary_rays)
                            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
                            out.append((ctx.readback(capi.BUF_ACCUM).copy(), ctx.readback(capi.BUF_LDR).copy(), rays))
                        for o in out[1:]:
                            what = f"{maker} {w}x{h} part {part} {spp} spp x {bounces} bounces"
                            assert o[2] == out[0][2], what
                            assert np.array_equal(o[0], out[0][0], equal_nan=True) and np.array_equal(o[1], out[0][1]), what
                        assert np.all(out[0][0][..., 3] == 2 * spp)
        finally:
            ctx.close()


def test_bands_do_not_change_the_image(oracle, sky_inputs, blue_noise):
    """Option bands (pixel ranges rendered on their own streams so that one band's traversal drain overlaps another
    band's kernels): accumulator, ray count and framebuffer equal the one-band render bit for bit, also with
    progressive accumulation and under a tile partition."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    ctx = capi.Context(0)
    try:
        ctx.upload_blue_noise(blue_noise)
        ctx.upload_mesh(pos, idx, alb)
        ctx.build()
        for (w, h), part in (((640, 411), (0, 1)), ((1001, 517), (1, 2))):
            cam = camera_for(oracle, view, w, h)
            setup_sky(ctx, oracle, atmo, cam.position[:])
            ctx.set_partition(part[0], part[1], 8)
            out = []
            for bands in (1, 2, 3, 4, 8, 1):
                ctx.set_option("bands", bands)
                rays = 0
                for frame in (5, 6):
                    pc, scn = oracle.constants(cam, frame=frame)
                    ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
                    ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 3, 2, capi.SECONDARY_ACCUMULATE if frame == 6 else 0)
                    st = ctx.stats()
                    assert st.stack_overflows == 0
                    rays += int(st.secondary_rays)
                ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
                out.append((ctx.readback(capi.BUF_ACCUM).copy(), ctx.readback(capi.BUF_LDR).copy(), rays))
            for o in out[1:]:
                assert o[2] == out[0][2]
                assert np.array_equal(o[0], out[0][0], equal_nan=True) and np.array_equal(o[1], out[0][1])
    finally:
        ctx.close()


def test_primary_entry_list_does_not_change_the_gbuffer(gpu_ctx, oracle):
    """Option primary_entry (the top of the BVH walked once per 8x4 tile against the tile's frustum, rays start from
    the entry list): visibility ids, hit distances and the fp16 G-buffer are bit-identical to the walk from the root --
    for cameras outside and inside the geometry, looking along the axes (tiles that straddle an octant boundary fall
    back), with prime image sizes (padding lanes), far from the origin (widened entry boxes), and ids/t equal the
    brute force over all triangles on sampled pixels."""
    rng = np.random.default_rng(12)
    cases = []
    for maker in ("cornell", "small_terrain", "hall_260k"):
        pos, idx, alb, view = getattr(scenes, maker)()
        cams = [(view["position"], view["yaw_deg"], view["pitch_deg"]), (view["position"], 90.0, 0.0),
                (tuple(np.asarray(pos.mean(0)) + [0.0, 0.0, 0.0005]), 37.0, -25.0), (view["position"], 180.0, 89.0)]
        cases.append((pos, idx, alb, cams))
    pos, idx, alb, view = scenes.small_terrain()
    shift = np.array([1000.0, -2000.0, 500.0], np.float32)   # far from the origin: coordinates' ulp ~ the ray offset
    cases.append((pos + shift, idx, alb, [(tuple(np.asarray(view["position"]) + shift), view["yaw_deg"], view["pitch_deg"])]))
    for pos, idx, alb, cams in cases:
        gpu_ctx.upload_mesh(pos, idx, alb)
        gpu_ctx.build()
        for (cpos, yaw, pitch) in cams:
            for (w, h) in ((331, 197), (1280, 720)):
                cam = oracle.make_camera(w, h, cpos, yaw, pitch)
                pc, _ = oracle.constants(cam, frame=1)
                out = []
                for entry in (0, 1):
                    gpu_ctx.set_option("primary_entry", entry)
                    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
                    assert gpu_ctx.stats().stack_overflows == 0
                    out.append([gpu_ctx.readback(b).copy() for b in (capi.BUF_VISIBILITY, capi.BUF_HIT_T, capi.BUF_DEPTH,
                                                                      capi.BUF_NORMAL, capi.BUF_MOTION)])
                for a, b in zip(*out):
                    assert np.array_equal(a, b), f"primary_entry changes the G-buffer ({w}x{h}, yaw {yaw}, pitch {pitch})"
                # sampled pixels vs brute force
                xs, ys = rng.integers(0, w, 400), rng.integers(0, h, 400)
                o = np.zeros((400, 3), np.float32)
                d = np.zeros((400, 3), np.float32)
                oo, dd = (C.c_float * 3)(), (C.c_float * 3)()
                for k in range(400):
                    oracle.lib().orc_ray_gen(C.byref(pc.invView), C.byref(pc.invProjection), int(xs[k]), int(ys[k]), w, h, oo, dd)
                    o[k], d[k] = oo[:], dd[:]
                ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
                assert np.array_equal(out[1][0][ys, xs], ids_bf) and np.array_equal(out[1][1][ys, xs], t_bf)
    gpu_ctx.set_option("primary_entry", 0)


def test_checkpoint_resume_of_a_progressive_render(oracle, sky_inputs, blue_noise):
    """SURVEY 5 checkpoint/resume: dump the fp32 accumulator after 2 frames, restore it in a fresh context, render
    frames 3-4 on top: bit-identical to the uninterrupted 4-frame accumulation."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 200, 120
    cam = camera_for(oracle, view, w, h)

    def fresh():
        c = capi.Context(0)
        setup_sky(c, oracle, atmo, cam.position[:])
        c.upload_blue_noise(blue_noise)
        c.upload_mesh(pos, idx, alb)
        c.build()
        return c

    def frame(c, f, accumulate):
        pc, scn = oracle.constants(cam, frame=f)
        c.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        c.secondary_rays(as_capi(scn, capi.SecondaryConstants), 2, 2, capi.SECONDARY_ACCUMULATE if accumulate else 0)

    a = fresh()
    try:
        for f in (1, 2, 3, 4):
            frame(a, f, f > 1)
            if f == 2:
                dump = a.readback(capi.BUF_ACCUM).copy()
        want = a.readback(capi.BUF_ACCUM).copy()
    finally:
        a.close()
    b = fresh()
    try:
        with pytest.raises(capi.MinoteError):
            b.accum_restore(dump)                      # size unknown before the first primary pass
        pc, _ = oracle.constants(cam, frame=3)
        b.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        with pytest.raises(capi.MinoteError):
            b.accum_restore(dump[:10])                 # wrong size
        b.accum_restore(dump)
        for f in (3, 4):
            frame(b, f, True)
        got = b.readback(capi.BUF_ACCUM)
        assert np.array_equal(got, want, equal_nan=True) and np.all(got[..., 3] == 8.0)
    finally:
        b.close()


def test_async_readback_matches_blocking(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Pipelined framebuffer readback (double-buffered LDR): frame f's async copy equals its blocking readback even
    when frame f+1 has been issued in between."""
    import torch
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.cornell()
    w, h = 160, 96
    cam = camera_for(oracle, view, w, h)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    bufs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    blocking = []
    for f in (1, 2, 3):
        pc, scn = oracle.constants(cam, frame=f)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 1, 2)
        gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
        gpu_ctx.readback_async(capi.BUF_LDR, C.c_void_p(bufs[f & 1].data_ptr()), bufs[f & 1].numel())
        gpu_ctx.readback_wait(1)
        if f > 1:  # frame f-1 has landed while frame f may still be copying
            assert np.array_equal(bufs[(f - 1) & 1].numpy(), blocking[-1])
        blocking.append(gpu_ctx.readback(capi.BUF_LDR).copy())
    gpu_ctx.readback_wait(0)
    assert np.array_equal(bufs[3 & 1].numpy(), blocking[-1])
    assert not np.array_equal(blocking[0], blocking[1])


def test_progressive_accumulation(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Row n7: frames f = 1..3 with seeds (f<<1)|1 accumulate; equals the oracle's accumulated sum."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.cornell()
    w, h = 128, 128
    cam = camera_for(oracle, view, w, h)
    trans, multi, view_lut = oracle.sky_luts(atmo, cam.position[:])
    osc = oracle.Scene(pos, idx, alb)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    acc = None
    for f in (1, 2, 3):
        pc, scn = oracle.constants(cam, frame=f)
        acc, _, _ = osc.render(w, h, pc, scn, blue_noise, atmo, trans, view_lut, 2, 2, accum=acc)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 2, 2, capi.SECONDARY_ACCUMULATE if f > 1 else 0)
    gacc = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.all(gacc[..., 3] == 6.0)
    assert oracle.psnr(oracle.resolve(gacc)[..., :3], oracle.resolve(acc)[..., :3], 16.0) >= 50.0


def test_tile_partition_equals_single_context(oracle, sky_inputs, blue_noise):
    """Row 8e: rank-r-of-N contexts render disjoint row slabs; reassembled image == 1-context image, bit for bit."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 160, 101  # not a multiple of the slab height
    cam = camera_for(oracle, view, w, h)
    pc, scn = oracle.constants(cam)

    def render(rank, nranks):
        ctx = capi.Context(0)
        try:
            setup_sky(ctx, oracle, atmo, cam.position[:])
            ctx.upload_blue_noise(blue_noise)
            ctx.upload_mesh(pos, idx, alb)
            ctx.build()
            ctx.set_partition(rank, nranks, 8)
            ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
            ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 2, 2)
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)
            return ctx.partition_rows(h), ctx.readback(capi.BUF_ACCUM), ctx.readback(capi.BUF_LDR), ctx.readback(capi.BUF_VISIBILITY)
        finally:
            ctx.close()

    rows1, acc1, ldr1, vis1 = render(0, 1)
    assert np.array_equal(rows1, np.arange(h))
    for nranks in (2, 3):
        acc = np.zeros_like(acc1)
        ldr = np.zeros_like(ldr1)
        vis = np.zeros_like(vis1)
        seen = np.zeros(h, bool)
        for r in range(nranks):
            rows, a, l, v = render(r, nranks)
            assert not seen[rows].any()
            seen[rows] = True
            acc[rows], ldr[rows], vis[rows] = a, l, v
        assert seen.all()
        assert np.array_equal(vis, vis1) and np.array_equal(acc, acc1) and np.array_equal(ldr, ldr1)


@pytest.mark.parametrize("builder", [0, 1])
@pytest.mark.parametrize("maker", ["tiny", "soup", "small_terrain", "hall_260k"])
def test_wide_refit_gives_the_same_nodes(gpu_ctx, maker, builder):
    """Option wide_refit (default): node emission and MRT_BUILD_REFIT level by level on the wide tree with 8 lanes per
    node, vs round 1's path (binary boxes climbed with atomics, one thread per wide node): a slot's box is the exact
    union of its triangles either way and the quantiser is the same code, so nodes and leaf triangles are
    bit-identical -- after a full build, after a refit of animated vertices, and when a wide refit follows a build
    emitted the old way."""
    if maker == "tiny":
        pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5], [2, 0, 1], [2, 2, 1]], np.float32)
        idx = np.array([[0, 1, 2], [1, 3, 2], [3, 4, 5]], np.uint32)
        alb = np.full((3, 3), 0.5, np.float32)
    elif maker == "soup":  # sizes over six decades, far from the origin
        rng = np.random.default_rng(13)
        n = 20000
        c = np.array([1.0e4, -2.0e4, 5.0e3], np.float32) + rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        size = (10.0 ** rng.uniform(-4, 2, (n, 1))).astype(np.float32)
        pos = (c[:, None, :] + size[:, None, :] * rng.normal(size=(n, 3, 3)).astype(np.float32)).reshape(-1, 3).astype(np.float32)
        idx = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
        alb = np.full((n, 3), 0.5, np.float32)
    else:
        pos, idx, alb, _ = getattr(scenes, maker)()
    rng = np.random.default_rng(5)
    pos2 = (pos + rng.normal(scale=0.01 * float(np.abs(pos).max() + 1e-3), size=pos.shape)).astype(np.float32)
    gpu_ctx.set_option("builder", builder)
    gpu_ctx.upload_mesh(pos, idx, alb)

    def tree():
        return gpu_ctx.readback(capi.BUF_BVH_NODES).copy(), gpu_ctx.readback(capi.BUF_BVH_TRIS).copy()

    got = {}
    for wide in (0, 1):
        gpu_ctx.set_option("wide_refit", wide)
        gpu_ctx.update_positions(pos)
        gpu_ctx.build()
        got[wide, "full"] = tree()
        gpu_ctx.update_positions(pos2)
        gpu_ctx.build(capi.BUILD_REFIT)
        got[wide, "refit"] = tree()
    # a wide refit on top of a build emitted by the one-thread-per-node kernel
    gpu_ctx.set_option("wide_refit", 0)
    gpu_ctx.update_positions(pos)
    gpu_ctx.build()
    gpu_ctx.set_option("wide_refit", 1)
    gpu_ctx.update_positions(pos2)
    gpu_ctx.build(capi.BUILD_REFIT)
    got["mixed", "refit"] = tree()
    for what in ("full", "refit"):
        assert np.array_equal(got[1, what][0], got[0, what][0]), f"{what}: nodes differ"
        assert np.array_equal(got[1, what][1], got[0, what][1]), f"{what}: leaf triangles differ"
    assert np.array_equal(got["mixed", "refit"][0], got[0, "refit"][0]) and np.array_equal(got["mixed", "refit"][1], got[0, "refit"][1])
    assert not np.array_equal(got[0, "full"][0], got[0, "refit"][0])  # the animation is visible in the nodes


@pytest.mark.parametrize("maker", ["tiny", "soup", "hall_260k", "scene_1m"])
def test_fused_sort_gives_the_same_tree(gpu_ctx, maker):
    """Option fused_sort (default): the Morton sort's eight passes inside one cooperative launch (one CTA per tile of 8192
    keys, grid barriers between the phases) vs five launches per pass: the same stable permutation, hence the same nodes
    and leaf triangles bit for bit.  scene_1m (128 of at most 148 tiles) is close to the co-residency limit of the fused kernel."""
    if maker == "tiny":
        pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5], [2, 0, 1], [2, 2, 1]], np.float32)
        idx = np.array([[0, 1, 2], [1, 3, 2], [3, 4, 5]], np.uint32)
        alb = np.full((3, 3), 0.5, np.float32)
    elif maker == "soup":  # many equal keys: duplicates exercise the stability of the ranking
        rng = np.random.default_rng(14)
        n = 30000
        c = rng.integers(0, 40, (n, 3)).astype(np.float32)
        pos = (c[:, None, :] + 0.25 * rng.integers(0, 3, (n, 3, 3)).astype(np.float32)).reshape(-1, 3).astype(np.float32)
        idx = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
        alb = np.full((n, 3), 0.5, np.float32)
    else:
        pos, idx, alb, _ = getattr(scenes, maker)()
    gpu_ctx.upload_mesh(pos, idx, alb)
    trees = []
    for fused in (0, 1):
        gpu_ctx.set_option("fused_sort", fused)
        gpu_ctx.build()
        trees.append((gpu_ctx.readback(capi.BUF_BVH_NODES).copy(), gpu_ctx.readback(capi.BUF_BVH_TRIS).copy()))
    assert np.array_equal(trees[0][0], trees[1][0]) and np.array_equal(trees[0][1], trees[1][1])


def test_refit_matches_full_rebuild(gpu_ctx, oracle):
    """BASELINE config 5 mechanics: animate vertices, REFIT; closest hits equal a fresh FULL build and brute force."""
    pos, idx, alb, _ = scenes.small_terrain()
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    moved = scenes.animate(pos, 0.37)
    gpu_ctx.update_positions(moved)
    gpu_ctx.build(capi.BUILD_REFIT)
    o, d = random_rays(8000, moved.min(0) - 0.001, moved.max(0) + 0.001, 3)
    ids_refit, t_refit = gpu_ctx.trace_rays(o, d)
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    gpu_ctx.build(capi.BUILD_FULL)
    ids_full, t_full = gpu_ctx.trace_rays(o, d)
    assert np.array_equal(ids_refit, ids_bf) and np.array_equal(t_refit, t_bf)
    assert np.array_equal(ids_full, ids_bf) and np.array_equal(t_full, t_bf)


def test_config2_full_size_properties(gpu_ctx, oracle, sky_inputs, blue_noise):
    """BASELINE config 2 at FULL size (260 864 triangles, 1920x1080, 1 spp, 2 bounces), checked through
    size-independent properties: (a) primary visibility ids equal GPU brute force and the oracle's brute force on
    sampled pixels; (b) the image is independent of the hierarchy builder (LBVH vs PLOC: both are exact
    closest-hit structures) and of the ray-sort stage; (c) rendering the same frame twice is bit-identical;
    (d) accumulating two frames equals the sum of the two frames rendered separately."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.hall_260k()
    w, h = 1920, 1080
    cam = camera_for(oracle, view, w, h)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)

    def render(frame, flags=0):
        pc, scn = oracle.constants(cam, frame=frame)
        gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 1, 2, flags)
        return gpu_ctx.readback(capi.BUF_ACCUM).copy()

    images = {}
    for builder in (0, 1):
        gpu_ctx.set_option("builder", builder)
        gpu_ctx.upload_mesh(pos, idx, alb)
        gpu_ctx.build()
        images[builder] = render(1)
        assert gpu_ctx.stats().stack_overflows == 0
    assert np.array_equal(images[0], images[1]), "image depends on the BVH builder"
    vis = gpu_ctx.readback(capi.BUF_VISIBILITY)
    t = gpu_ctx.readback(capi.BUF_HIT_T)
    # (a) sampled pixels against brute force
    rng = np.random.default_rng(2)
    xs, ys = rng.integers(0, w, 3000), rng.integers(0, h, 3000)
    pc, _ = oracle.constants(cam, frame=1)
    o = np.zeros((3000, 3), np.float32)
    d = np.zeros((3000, 3), np.float32)
    oo, dd = (C.c_float * 3)(), (C.c_float * 3)()
    for k in range(3000):
        oracle.lib().orc_ray_gen(C.byref(pc.invView), C.byref(pc.invProjection), int(xs[k]), int(ys[k]), w, h, oo, dd)
        o[k], d[k] = oo[:], dd[:]
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    assert np.array_equal(vis[ys, xs], ids_bf), "primary visibility differs from brute force at full size"
    assert np.array_equal(t[ys, xs], t_bf)
    osc = oracle.Scene(pos, idx, alb)
    for k in range(0, 3000, 150):
        i, tt, _, _ = osc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids_bf[k] and (i == oracle.NONE_ID or np.float32(tt) == t_bf[k])
    # (b) sort stage, (c) determinism
    assert np.array_equal(render(1, capi.SECONDARY_SORT_RAYS), images[1])
    assert np.array_equal(render(1), images[1])
    # (d) linearity of the accumulator
    f2 = render(2)
    render(1)
    pc, scn = oracle.constants(cam, frame=2)
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 1, 2, capi.SECONDARY_ACCUMULATE)
    both = gpu_ctx.readback(capi.BUF_ACCUM)
    assert np.array_equal(both, images[1] + f2)
    assert np.all(both[..., 3] == 2.0)


@pytest.mark.parametrize("builder", [0, 1], ids=["lbvh", "ploc"])
def test_mixed_scale_soup_far_from_origin(gpu_ctx, oracle, builder):
    """Conservativeness of the quantised boxes under stress: 20 000 triangles whose sizes span six decades,
    centred 10 000 units from the origin (few mantissa bits left for the node grids), rays from inside and
    outside the cloud.  BVH result == GPU brute force (ids and t bit-exact), sampled rays == oracle brute force."""
    rng = np.random.default_rng(11 + builder)
    n = 20000
    centre = np.array([1.0e4, -2.0e4, 5.0e3], np.float32)
    c = centre + rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    size = (10.0 ** rng.uniform(-4, 2, (n, 1))).astype(np.float32)
    pos = (c[:, None, :] + size[:, None, :] * rng.normal(size=(n, 3, 3)).astype(np.float32)).reshape(-1, 3).astype(np.float32)
    idx = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    alb = np.full((n, 3), 0.5, np.float32)
    gpu_ctx.set_option("builder", builder)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    m = 30000
    o = (centre + rng.uniform(-80, 80, (m, 3))).astype(np.float32)
    d = rng.normal(size=(m, 3)).astype(np.float32)
    d[: m // 10, 0] = 0.0          # axis-parallel components
    d[m // 10: m // 5, 1:] = 0.0   # axis-aligned rays
    d[m // 5] = (1.0, 0.0, 0.0)
    ids, t = gpu_ctx.trace_rays(o, d)
    ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
    assert gpu_ctx.stats().stack_overflows == 0
    assert np.array_equal(ids, ids_bf)
    assert np.array_equal(t, t_bf)
    assert (ids != capi.MISS_ID).mean() > 0.3
    osc = oracle.Scene(pos, idx, alb)
    for k in range(0, m, 500):
        i, tt, _, _ = osc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids[k] and (i == oracle.NONE_ID or np.float32(tt) == t[k])
    gpu_ctx.set_option("builder", 1)


def _icosphere(level):
    g = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, g, 0), (1, g, 0), (-1, -g, 0), (1, -g, 0), (0, -1, g), (0, 1, g), (0, -1, -g), (0, 1, -g),
         (g, 0, -1), (g, 0, 1), (-g, 0, -1), (-g, 0, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                p = v[a] + v[b]
                v.append(p / np.linalg.norm(p))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v, np.float32), np.array(f, np.uint32)


def test_watertight_closed_mesh(gpu_ctx, oracle):
    """A closed icosphere (20 480 triangles, shared vertices) seen from inside: rays aimed exactly at every vertex and
    every edge midpoint, plus random ones, must all hit something -- no ray leaks through a shared edge or vertex --
    and agree with brute force."""
    pos, idx = _icosphere(5)
    pos = (pos * np.float32(3.0) + np.array([0.3, -0.2, 0.1], np.float32)).astype(np.float32)
    alb = np.full((idx.shape[0], 3), 0.5, np.float32)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    e = np.concatenate([idx[:, [0, 1]], idx[:, [1, 2]], idx[:, [2, 0]]])
    mids = (0.5 * (pos[e[:, 0]].astype(np.float64) + pos[e[:, 1]].astype(np.float64))).astype(np.float32)
    rng = np.random.default_rng(5)
    for origin in ([0.3, -0.2, 0.1], [1.1, 0.4, -0.9], [0.0, 0.0, 0.0]):
        org = np.array(origin, np.float32)
        targets = np.concatenate([pos, mids, org + rng.normal(size=(20000, 3)).astype(np.float32)])
        d = (targets - org).astype(np.float32)
        o = np.tile(org, (d.shape[0], 1))
        ids, t = gpu_ctx.trace_rays(o, d)
        assert (ids != capi.MISS_ID).all(), f"{(ids == capi.MISS_ID).sum()} rays leaked out of a closed mesh"
        ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
        assert np.array_equal(ids, ids_bf) and np.array_equal(t, t_bf)
    osc = oracle.Scene(pos, idx, alb)
    for k in range(0, pos.shape[0], 97):   # vertex rays: up to six triangles tie, the lowest primitive id wins
        i, tt, _, _ = osc.closest_hit(o[k], d[k], use_bvh=False)
        assert i == ids[k] and np.float32(tt) == t[k]


@pytest.mark.parametrize("builder", [0, 1], ids=["lbvh", "ploc"])
@pytest.mark.parametrize("maker", ["cornell", "small_terrain", "hall_260k", "tiny", "soup"])
def test_device_side_build_loops_give_the_same_tree(gpu_ctx, maker, builder):
    """PLOC rounds and collapse levels looped inside cooperative kernels (default) vs driven from the host with
    a readback per round: node ids come from the same scans, so nodes and leaf triangles are bit-identical.
    hall_260k runs the grid-wide rounds (> 2048 clusters), the small scenes only the single-CTA tail."""
    if maker == "tiny":
        pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5], [2, 0, 1], [2, 2, 1]], np.float32)
        idx = np.array([[0, 1, 2], [1, 3, 2], [3, 4, 5]], np.uint32)
        alb = np.full((3, 3), 0.5, np.float32)
    elif maker == "soup":  # sizes over six decades, far from the origin: ties and near-degenerate boxes
        rng = np.random.default_rng(12)
        n = 20000
        c = np.array([1.0e4, -2.0e4, 5.0e3], np.float32) + rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        size = (10.0 ** rng.uniform(-4, 2, (n, 1))).astype(np.float32)
        pos = (c[:, None, :] + size[:, None, :] * rng.normal(size=(n, 3, 3)).astype(np.float32)).reshape(-1, 3).astype(np.float32)
        idx = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
        alb = np.full((n, 3), 0.5, np.float32)
    else:
        pos, idx, alb, _ = getattr(scenes, maker)()
    gpu_ctx.set_option("builder", builder)
    gpu_ctx.upload_mesh(pos, idx, alb)
    trees = []
    for device_loop in (0, 1, 1):
        gpu_ctx.set_option("build_device_loop", device_loop)
        gpu_ctx.build()
        st = gpu_ctx.stats()
        trees.append((gpu_ctx.readback(capi.BUF_BVH_NODES).copy(), gpu_ctx.readback(capi.BUF_BVH_TRIS).copy(), st.num_wide_nodes))
    for nodes, tris, nn in trees[1:]:
        assert nn == trees[0][2]
        assert np.array_equal(nodes, trees[0][0])
        assert np.array_equal(tris, trees[0][1])
    assert trees[0][0].size == trees[0][2] * 20 and trees[0][1].size == idx.shape[0] * 12


def test_shared_scene_renders_the_same_frames(gpu_ctx, oracle, sky_inputs, blue_noise):
    """mrt_scene_share (frames in flight): a context that borrows another context's triangles + BVH produces the same
    hits and the same image bit for bit, may not build or update the borrowed scene, follows the owner's refit after
    sharing again, and can be destroyed before or go back to a scene of its own."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 160, 96
    cam = camera_for(oracle, view, w, h)
    owner = gpu_ctx
    owner.upload_blue_noise(blue_noise)
    setup_sky(owner, oracle, atmo, cam.position[:])
    owner.upload_mesh(pos, idx, alb)
    other = capi.Context(0)
    try:
        with pytest.raises(capi.MinoteError):  # nothing built yet
            other.share_scene(owner)
        owner.build()
        other.upload_blue_noise(blue_noise)
        setup_sky(other, oracle, atmo, cam.position[:])
        other.share_scene(owner)
        so, sb = owner.stats(), other.stats()
        assert sb.num_triangles == so.num_triangles and sb.num_wide_nodes == so.num_wide_nodes and sb.bvh_bytes == so.bvh_bytes

        def render(ctx, f):
            pc, scn = oracle.constants(cam, frame=f)
            ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
            ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 2, 2)
            ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_ACCUM)

        # both contexts in flight at once, different frames each, then compared with the owner alone
        for f in (1, 2):
            render(owner, f)
            render(other, f + 10)
        got = (other.readback(capi.BUF_ACCUM).copy(), other.readback(capi.BUF_LDR).copy(), other.readback(capi.BUF_VISIBILITY).copy())
        render(owner, 12)
        want = (owner.readback(capi.BUF_ACCUM), owner.readback(capi.BUF_LDR), owner.readback(capi.BUF_VISIBILITY))
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        o, d = random_rays(5000, pos.min(0) - 0.001, pos.max(0) + 0.001, 3)
        ids_a, t_a = owner.trace_rays(o, d)
        ids_b, t_b = other.trace_rays(o, d)
        ids_bf, _ = other.trace_rays(o, d, brute_force=True)
        assert np.array_equal(ids_a, ids_b) and np.array_equal(t_a, t_b) and np.array_equal(ids_b, ids_bf)
        # the borrower cannot change the scene
        for call in (lambda: other.build(), lambda: other.build(capi.BUILD_REFIT), lambda: other.update_positions(pos)):
            with pytest.raises(capi.MinoteError):
                call()
        # owner animates + refits while the borrower has a frame in flight: the library waits for that frame, the
        # borrower is stale until it shares again, then it sees the new geometry
        render(other, 6)
        pos2 = pos.copy()
        pos2[:, 1] += 0.0004 * np.sin(40.0 * pos[:, 0]).astype(np.float32)
        owner.update_positions(pos2)
        owner.build(capi.BUILD_REFIT)
        with pytest.raises(capi.MinoteError, match="mrt_scene_share again"):
            render(other, 7)
        with pytest.raises(capi.MinoteError):
            other.trace_rays(o, d)
        other.share_scene(owner)
        render(other, 5)
        render(owner, 5)
        assert np.array_equal(other.readback(capi.BUF_ACCUM), owner.readback(capi.BUF_ACCUM))
        ids_b2, t_b2 = other.trace_rays(o, d)
        ids_bf2, t_bf2 = other.trace_rays(o, d, brute_force=True)
        assert np.array_equal(ids_b2, ids_bf2) and np.array_equal(t_b2, t_bf2)
        assert not np.array_equal(t_b2, t_b)
        # a scene of its own ends the borrowing; the owner's arrays stay alive
        cpos, cidx, calb, _ = scenes.cornell()
        other.upload_mesh(cpos, cidx, calb)
        other.build()
        assert other.stats().num_triangles == cidx.shape[0]
        ids_a2, _ = owner.trace_rays(o, d)
        assert np.array_equal(ids_a2, ids_bf2)
        other.share_scene(owner)  # and back: frees its own copy
    finally:
        other.close()  # borrower destroyed first; the owner (fixture) still works afterwards
    ids_a3, _ = owner.trace_rays(o, d)
    assert np.array_equal(ids_a3, ids_bf2)


def test_shared_scene_owner_destroyed_first(oracle, sky_inputs, blue_noise):
    """Destroying the owner while a borrower still exists leaves the borrower without a scene (errors, no dangling
    device pointers); the borrower can take a scene of its own afterwards."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.cornell()
    w, h = 64, 48
    cam = camera_for(oracle, view, w, h)
    pc, scn = oracle.constants(cam)
    owner, other = capi.Context(0), capi.Context(0)
    try:
        owner.upload_mesh(pos, idx, alb)
        owner.build()
        other.upload_blue_noise(blue_noise)
        setup_sky(other, oracle, atmo, cam.position[:])
        other.share_scene(owner)
        other.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        other.secondary_rays(as_capi(scn, capi.SecondaryConstants), 1, 1)   # in flight while the owner goes away
        owner.close()
        with pytest.raises(capi.MinoteError):
            other.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        with pytest.raises(capi.MinoteError):
            other.build()
        other.upload_mesh(pos, idx, alb)
        other.build()
        other.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
        assert (other.readback(capi.BUF_VISIBILITY) != capi.MISS_ID).any()
    finally:
        other.close()
        owner.close()


def test_renderer_frames_in_flight(oracle, blue_noise):
    """Renderer::draw with 3 frame contexts (the reference's frames in flight, renderer.ixx:36): every frame's
    framebuffer equals the one the single-context renderer produces for the same frame counter, also across a mesh
    update (refit) between frames, with the async readbacks of up to 2 frames pending."""
    import torch
    from minotert_b200 import host
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 160, 96
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    pos2 = pos.copy()
    pos2[:, 1] += 0.0004 * np.sin(40.0 * pos[:, 0]).astype(np.float32)
    nframes = 7

    def run(in_flight):
        r = host.Renderer(w, h, blue_noise, frames_in_flight=in_flight)
        try:
            assert r.frames_in_flight() == in_flight
            r.set_mesh(pos, idx, alb)
            r.configure(samples=2, bounces=2, accumulate=False, tonemap="amd", exposure=1.0)
            r.stats_reset()  # starts the running ray totals of every frame context
            bufs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(nframes)]
            for f in range(nframes):
                if f == 4:
                    r.update_mesh(pos2, refit=True)
                r.draw(cam)
                assert r.frame_count() == f + 1
                r.read_framebuffer_async(C.c_void_p(bufs[f].data_ptr()), bufs[f].numel())
                r.wait_framebuffer(in_flight - 1)
            r.wait_framebuffer(0)
            st = r.stats()
            return [b.numpy().copy() for b in bufs], int(st.total_rays)
        finally:
            r.close()

    one, rays1 = run(1)
    three, rays3 = run(3)
    assert rays1 == rays3 and rays1 > nframes * w * h
    for f in range(nframes):
        assert np.array_equal(one[f], three[f]), f"frame {f + 1}"
    assert not np.array_equal(one[0], one[1])      # different seeds
    assert not np.array_equal(one[3], one[4])      # the refit is visible


def test_renderer_async_mesh_updates(oracle, blue_noise):
    """Option async_update (animated scenes, BASELINE config 5): Renderer::updateMesh(refit) returns without waiting for
    the GPU -- the vertex upload runs on its own stream, the refit is ordered behind the frames in flight and before the
    next ones with events, the borrowed scenes stay valid.  A new vertex array EVERY frame, 1 and 3 frames in flight,
    with a synchronous full rebuild in the middle: every framebuffer equals the synchronous renderer's."""
    import torch
    from minotert_b200 import host
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 160, 96
    cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    nframes = 9
    anim = [torch.from_numpy(np.ascontiguousarray(scenes.animate(pos, f / 10.0, amplitude=0.002))).pin_memory() for f in range(nframes)]

    def run(in_flight, async_update):
        r = host.Renderer(w, h, blue_noise, frames_in_flight=in_flight)
        try:
            r.set_mesh(pos, idx, alb)
            r.configure(samples=2, bounces=2, accumulate=False, tonemap="amd", exposure=1.0)
            r.set_option("async_update", async_update)
            bufs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(nframes)]
            for f in range(nframes):
                r.update_mesh(anim[f].numpy(), refit=(f != 5))
                r.draw(cam)
                r.read_framebuffer_async(C.c_void_p(bufs[f].data_ptr()), bufs[f].numel())
                r.wait_framebuffer(in_flight - 1)
            r.wait_framebuffer(0)
            assert r.stats().ms_build > 0.0
            return [b.numpy().copy() for b in bufs]
        finally:
            r.close()

    want = run(1, 0)
    assert not np.array_equal(want[0], want[1])
    for in_flight in (1, 3):
        got = run(in_flight, 1)
        for f in range(nframes):
            assert np.array_equal(got[f], want[f]), f"{in_flight} frame(s) in flight, frame {f + 1}"


def test_tonemap_of_color_without_materialising_it(gpu_ctx, oracle, sky_inputs, blue_noise):
    """Triangle path, tonemap source COLOR: while nobody has asked for the RGBA16F image the tonemapper reads the fp32
    accumulator and rounds its average through fp16 in registers; the framebuffer is bit-identical to resolving
    MRT_BUF_COLOR first, and to the oracle's tonemapper on that RGBA16F image within its 1-code-value bar."""
    atmo = sky_inputs[0]
    pos, idx, alb, view = scenes.small_terrain()
    w, h = 192, 128
    cam = camera_for(oracle, view, w, h)
    pc, scn = oracle.constants(cam, frame=2)
    setup_sky(gpu_ctx, oracle, atmo, cam.position[:])
    gpu_ctx.upload_blue_noise(blue_noise)
    gpu_ctx.upload_mesh(pos, idx, alb)
    gpu_ctx.build()
    gpu_ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    gpu_ctx.secondary_rays(as_capi(scn, capi.SecondaryConstants), 2, 2)
    launches0 = gpu_ctx.stats().kernel_launches
    gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_COLOR)
    assert gpu_ctx.stats().kernel_launches == launches0 + 1      # no resolve kernel
    direct = gpu_ctx.readback(capi.BUF_LDR).copy()
    col16 = gpu_ctx.readback(capi.BUF_COLOR)                      # materialises the RGBA16F image
    gpu_ctx.tonemap("amd", 1.0, oracle.AMD_DEFAULT, capi.BUF_COLOR)
    assert np.array_equal(gpu_ctx.readback(capi.BUF_LDR), direct)
    d = np.abs(direct.astype(int) - oracle.tonemap("amd", col16).astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01


@pytest.mark.parametrize("shift", [0.0, 10.0, 1000.0])
def test_axis_aligned_walls_far_from_origin(gpu_ctx, shift):
    """Flat leaf boxes and flat nodes (the Cornell walls are axis-aligned planes) at coordinates where one ulp is up to
    1 % of the scene: rays that start exactly on vertices, edges and wall planes, axis-parallel rays running inside
    a wall's plane, and random rays -- BVH closest hits == brute force, bit for bit.  (The node quantiser keeps its
    grid step >= 2 ulp of the coordinates and evaluates planes exactly; CPU proof: tests/test_oracle_wide_node.py.)"""
    pos, idx, alb, _ = scenes.cornell()
    pos = (pos.astype(np.float64) + shift).astype(np.float32)
    rng = np.random.default_rng(11)
    for builder in (0, 1):
        gpu_ctx.set_option("builder", builder)
        gpu_ctx.upload_mesh(pos, idx, alb)
        gpu_ctx.build()
        lo, hi = pos.min(0), pos.max(0)
        ext = (hi - lo).max()
        o1, d1 = random_rays(6000, lo - 0.1 * ext, hi + 0.1 * ext, 5)
        # origins exactly on mesh vertices / edge midpoints, random directions
        verts = pos[rng.integers(0, pos.shape[0], 3000)]
        tri = idx[rng.integers(0, idx.shape[0], 3000)]
        mids = (0.5 * (pos[tri[:, 0]].astype(np.float64) + pos[tri[:, 1]])).astype(np.float32)
        o2 = np.concatenate([verts, mids])
        d2 = rng.normal(size=o2.shape)
        d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
        # axis-parallel rays from vertices: they run inside the planes of the walls the vertex belongs to
        o3 = np.repeat(pos[rng.integers(0, pos.shape[0], 500)], 6, axis=0)
        d3 = np.tile(np.float32([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]), (500, 1))
        o = np.concatenate([o1, o2, o3]).astype(np.float32)
        d = np.concatenate([d1, d2, d3]).astype(np.float32)
        ids, t = gpu_ctx.trace_rays(o, d)
        ids_bf, t_bf = gpu_ctx.trace_rays(o, d, brute_force=True)
        assert gpu_ctx.stats().stack_overflows == 0
        bad = np.flatnonzero((ids != ids_bf) | (t != t_bf))
        assert bad.size == 0, f"builder {builder}, shift {shift}: {bad.size} rays differ, first {bad[:5]}: ids {ids[bad[:5]]} vs {ids_bf[bad[:5]]}, t {t[bad[:5]]} vs {t_bf[bad[:5]]}"
        assert (ids != capi.MISS_ID).mean() > 0.3
