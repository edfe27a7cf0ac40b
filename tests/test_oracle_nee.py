"""SURVEY 8f-4 extensions of the native path tracer, CPU side: the oracle's contract for the sun as a sampled light
(orc_render_tris_ext, ORC_EXT_NEE_SUN) and for the sky evaluated at the shaded point (ORC_EXT_SKY_AT_HIT).

The reference has no counterpart (secondaryRays.comp lights the scene only through escaping rays), so the contract is
checked against closed forms: the sample lies in the sun's disc and carries limb * Omega / pi; on an unoccluded horizontal
plane the direct term of the rendered image equals albedo * E * cos(theta_sun) * (2/3) * Omega / pi, where 2/3 is the disc
average of the limb-darkening law of getSunLuminance (sky.glsl:150-151); a blocker between the plane and the sun removes
exactly that term; ext = 0 is orc_render_tris bit for bit."""
import numpy as np

SUN = np.array([-0.435286462, 0.818654716, 0.374606609], np.float64)
SUN_RADIUS = 0.5 * 0.505 * 3.14159 / 180.0
M = 0.001


def test_sun_sample_lies_in_the_disc_and_carries_limb_times_solid_angle(oracle):
    one_minus_cos = 1.0 - np.cos(np.float32(SUN_RADIUS), dtype=np.float32)
    ws = []
    for u0 in np.linspace(0.0, 0.999, 40):
        for u1 in np.linspace(0.0, 0.97, 12):
            l, w = oracle.nee_sun_sample(u0, u1)
            assert abs(np.linalg.norm(l.astype(np.float64)) - 1.0) < 1e-6
            c = float(l.astype(np.float64) @ SUN)
            assert c >= np.cos(SUN_RADIUS) - 3e-7, (u0, u1, c)
            assert abs(c - (1.0 - u0 * float(one_minus_cos))) < 3e-7
            assert abs(w - np.sqrt(max(1.0 - u0, 1e-4)) * 2.0 * float(one_minus_cos)) < 1e-9
            ws.append(w)
    # disc average of the limb law: int_0^1 sqrt(1 - u) du = 2/3
    fine = [oracle.nee_sun_sample(u, 0.3)[1] for u in (np.arange(4000) + 0.5) / 4000]
    omega_over_pi = 2.0 * (1.0 - np.cos(SUN_RADIUS))
    assert abs(np.mean(fine) / omega_over_pi - 2.0 / 3.0) < 2e-3


def plane_scene(blocker):
    """A 40 m horizontal square at z = 0.0985 (normal +z) seen from 6 m above; optionally a 30 m square 1 km away towards
    the sun, facing it: it hides the whole disc from every visible point and 3e-4 sr of the sky."""
    z = 0.0985
    s = 20.0 * M
    pos = [[-s, -s, z], [s, -s, z], [s, s, z], [-s, s, z]]
    idx = [[0, 1, 2], [0, 2, 3]]
    if blocker:
        c = np.array([0.0, 0.0, z]) + SUN * 1.0
        t = np.cross([0.0, 0.0, 1.0], SUN); t /= np.linalg.norm(t)
        b = np.cross(SUN, t)
        r = 15.0 * M
        base = len(pos)
        pos += [list(c - t * r - b * r), list(c + t * r - b * r), list(c + t * r + b * r), list(c - t * r + b * r)]
        idx += [[base, base + 1, base + 2], [base, base + 2, base + 3]]
    pos = np.array(pos, np.float32)
    idx = np.array(idx, np.uint32)
    alb = np.tile(np.array([[0.5, 0.6, 0.7]], np.float32), (idx.shape[0], 1))
    return pos, idx, alb


def render(oracle, sky_inputs, blue_noise, scene, ext, spp=16, bounces=1, w=48, h=27):
    atmo = sky_inputs[0]
    pos, idx, alb = scene
    cam = oracle.make_camera(w, h, (0.0, 0.0, 0.0985 + 6.0 * M), 90.0, -89.0, vfov_deg=40.0)
    pc, sc = oracle.constants(cam, frame=1)
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    osc = oracle.Scene(pos, idx, alb)
    acc, vis, rays = osc.render(w, h, pc, sc, blue_noise, atmo, trans, view, spp, bounces, use_bvh=False, ext=ext)
    return acc, vis, rays, (atmo, trans, view, cam)


def test_ext_zero_is_the_plain_path_tracer(oracle, sky_inputs, blue_noise):
    a, _, ra, _ = render(oracle, sky_inputs, blue_noise, plane_scene(False), 0, spp=2, bounces=2)
    atmo = sky_inputs[0]
    pos, idx, alb = plane_scene(False)
    cam = oracle.make_camera(48, 27, (0.0, 0.0, 0.0985 + 6.0 * M), 90.0, -89.0, vfov_deg=40.0)
    pc, sc = oracle.constants(cam, frame=1)
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    import ctypes as C
    acc = np.zeros((27, 48, 4), np.float32)
    vis = np.zeros((27, 48), np.uint32)
    rays = (C.c_uint64 * 2)()
    osc = oracle.Scene(pos, idx, alb)
    oracle.lib().orc_render_tris(osc.h, 48, 27, C.byref(pc), C.byref(sc), oracle._p(blue_noise, C.c_uint8), blue_noise.shape[1],
                                 blue_noise.shape[0], C.byref(atmo), oracle._p(trans, C.c_uint16), oracle._p(view, C.c_uint32), 2, 2, 0,
                                 0, 27, oracle._p(acc, C.c_float), oracle._p(vis, C.c_uint32), rays)
    assert np.array_equal(a, acc) and (rays[0], rays[1]) == ra


def test_direct_sun_term_on_an_open_plane_matches_the_closed_form(oracle, sky_inputs, blue_noise):
    spp = 24
    lit, vis, rays_lit, (atmo, trans, view, cam) = render(oracle, sky_inputs, blue_noise, plane_scene(False), oracle.EXT_NEE_SUN, spp)
    dark, vis2, rays_dark, _ = render(oracle, sky_inputs, blue_noise, plane_scene(True), oracle.EXT_NEE_SUN, spp)
    assert np.all(vis == vis2) and np.all(vis < 2)          # every pixel sees the plane in both scenes
    # same seeds, same draws: the two renders differ by the direct term (and by the few bounce rays the blocker catches)
    direct = (lit[..., :3] - dark[..., :3]) / spp
    E = oracle.sun_centre_radiance(atmo, trans, view, cam.position[:]).astype(np.float64)
    assert np.all(E > 1000.0)
    omega_over_pi = 2.0 * (1.0 - np.cos(SUN_RADIUS))
    want = np.array([0.5, 0.6, 0.7]) * E * SUN[2] * (2.0 / 3.0) * omega_over_pi
    got = direct.reshape(-1, 3).mean(0)
    assert np.allclose(got, want, rtol=0.02), (got, want)
    assert direct.min() > -1e-3 * want.max()
    # shadow rays are counted: one per hit vertex that bounces (n.l > 0 everywhere on the plane), in both scenes
    npx = vis.size
    assert rays_lit[1] == 2 * spp * npx and rays_dark[1] == 2 * spp * npx
    # the blocked scene receives no sun at all: it equals the sky-view-only light of the open scene up to the blocker's
    # 3e-4 sr of sky
    assert abs(dark[..., :3].sum() / (lit[..., :3].sum() - direct.sum() * spp) - 1.0) < 5e-3


def test_nee_replaces_the_disc_in_escaping_bounce_rays_only(oracle, sky_inputs, blue_noise):
    """A pixel whose PRIMARY ray escapes still sees the sun's disc; per-pixel: with one bounce and the plane lit, the NEE
    image minus its direct term is the plain image minus the (rare, huge) disc hits of bounce rays."""
    spp = 8
    nee, vis, _, _ = render(oracle, sky_inputs, blue_noise, plane_scene(False), oracle.EXT_NEE_SUN, spp)
    plain, _, _, _ = render(oracle, sky_inputs, blue_noise, plane_scene(False), 0, spp)
    assert np.isfinite(nee).all()
    # the plain estimator has no direct term: almost every pixel is darker than its NEE counterpart (a pixel is brighter
    # only when one of its 8 bounce rays hit the 6e-5 sr disc: radiance ~1e5)
    brighter = (plain[..., :3].sum(-1) > nee[..., :3].sum(-1)).mean()
    assert brighter < 0.02


def test_sky_at_hit_moves_the_sky_position_by_metres_only(oracle, sky_inputs, blue_noise):
    a, _, ra, _ = render(oracle, sky_inputs, blue_noise, plane_scene(False), 0, spp=4, bounces=2)
    b, _, rb, _ = render(oracle, sky_inputs, blue_noise, plane_scene(False), oracle.EXT_SKY_AT_HIT, spp=4, bounces=2)
    assert ra == rb
    assert not np.array_equal(a, b)                       # 6 m lower: the sky-view parameterisation moves in the last digits
    assert np.allclose(a[..., :3], b[..., :3], rtol=5e-2, atol=1e-2)


# ---- aerial-perspective volume (ORC_EXT_AERIAL) ----

def far_wall_scene(distance_km=30.0):
    """A 40 km wall facing the camera `distance_km` away along +y (the default camera's view axis), plus the ground plane
    of plane_scene so that near pixels exist too."""
    z = 0.0985
    d, s = distance_km, 20.0
    pos = np.array([[-s, d, z - 12.0], [s, d, z - 12.0], [s, d, z + 12.0], [-s, d, z + 12.0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    alb = np.tile(np.array([[0.5, 0.5, 0.5]], np.float32), (2, 1))
    return pos, idx, alb


def test_aerial_volume_and_lookup(oracle, sky_inputs):
    atmo = sky_inputs[0]
    cam = oracle.make_camera(64, 36, (0.0, 0.0, 0.1), 90.0, 20.0)   # looking up: no froxel of the centre column is below ground
    pc, sc = oracle.constants(cam, frame=1)
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    vol = oracle.aerial_perspective(atmo, trans, multi, pc, cam.position[:])
    v = oracle.f16_to_f32(vol)
    assert np.isfinite(v).all() and (v >= 0).all() and (v[..., 3] <= 1.0).all()
    col = v[:, 16, 16]
    assert np.all(np.diff(col[:, 3]) >= -1e-3)          # opacity grows with depth along a ray that leaves the atmosphere
    assert col[-1, 3] > 0.05 and col[0, 3] < 2e-3
    assert np.all(np.diff(col[:, :3].sum(-1)) >= -1e-3)  # so does the in-scattered luminance
    # lookup: texel centres reproduce the texel; t -> 0 fades to nothing; clamp to edge beyond the last slice
    z = 10
    t_centre = ((z + 0.5) / 32.0) ** 2 * 32.0 * 4.0
    got = oracle.aerial_lookup(vol, (16 + 0.5) / 32, (16 + 0.5) / 32, t_centre)
    assert np.allclose(got, v[z, 16, 16], rtol=2e-3, atol=1e-6)
    assert np.all(oracle.aerial_lookup(vol, 0.5, 0.5, 0.0) == 0.0)
    near = oracle.aerial_lookup(vol, 0.5, 0.5, 0.01)
    assert 0.0 < near[3] < 1e-3
    assert np.allclose(oracle.aerial_lookup(vol, 0.5, 0.5, 500.0), oracle.aerial_lookup(vol, 0.5, 0.5, 127.9), rtol=0.05)


def test_aerial_perspective_dims_a_far_wall_and_adds_inscatter(oracle, sky_inputs, blue_noise):
    atmo = sky_inputs[0]
    w, h, spp = 48, 27, 4
    pos, idx, alb = far_wall_scene()
    cam = oracle.make_camera(w, h, (0.0, 0.0, 0.1), 90.0, 0.0, vfov_deg=30.0)
    pc, sc = oracle.constants(cam, frame=1)
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    vol = oracle.aerial_perspective(atmo, trans, multi, pc, cam.position[:])
    osc = oracle.Scene(pos, idx, alb)
    plain, vis, _ = osc.render(w, h, pc, sc, blue_noise, atmo, trans, view, spp, 1, use_bvh=False)
    hazy, vis2, _ = osc.render(w, h, pc, sc, blue_noise, atmo, trans, view, spp, 1, use_bvh=False, ext=oracle.EXT_AERIAL, aerial=vol)
    assert np.array_equal(vis, vis2) and (vis != oracle.NONE_ID).mean() > 0.9
    hit = vis != oracle.NONE_ID
    ys, xs = np.nonzero(hit)
    for y, x in list(zip(ys, xs))[::97]:
        # same RNG stream: per pixel hazy = plain * (1 - a) + spp * inscatter, with the pixel's own lookup
        o, d = oracle.ray_gen(pc, x, y, w, h)
        t = (30.0 - o[1]) / d[1]
        ap = oracle.aerial_lookup(vol, (x + 0.5) / w, (y + 0.5) / h, t)
        want = plain[y, x, :3] * (1.0 - ap[3]) + spp * ap[:3]
        assert np.allclose(hazy[y, x, :3], want, rtol=2e-4, atol=1e-5), (x, y)
        assert 0.2 < ap[3] < 0.6                     # 30 km of air: a third of the light is gone
    assert np.array_equal(hazy[~hit], plain[~hit])   # escaping primary rays are left alone
