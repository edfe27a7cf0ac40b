"""CPU tests of the oracle: known-answer tests derivable from the reference's formulas,
internal consistency (BVH == brute force), and the committed golden fixtures.
The reference ships no tests or golden vectors, so parity is "unpinned" by it; these KATs are what pins
the restatement."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz")


def test_pcg_known_answers(oracle):
    # src/gpu/random.glsl:22-27 evaluated by hand/independently in numpy below
    L = oracle.lib()
    s = C.c_uint32(3)
    assert [L.orc_pcg(C.byref(s)) for _ in range(4)] == [0x7F0EF6BC, 0xD3016361, 0x0C4A3341, 0x8AA33A8E]
    s = C.c_uint32(1)
    assert [L.orc_pcg(C.byref(s)) for _ in range(2)] == [0xA8BEEA3C, 0xE92A518A]

    def pcg_np(v):
        state = (v * 747796405 + 2891336453) & 0xFFFFFFFF
        word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
        return ((word >> 22) ^ word) & 0xFFFFFFFF
    v = 12345
    s = C.c_uint32(v)
    for _ in range(100):
        v = pcg_np(v)
        assert L.orc_pcg(C.byref(s)) == v
    s = C.c_uint32(7)
    f = L.orc_random_float(C.byref(s))
    assert f == (pcg_np(7) & 0xFFFFFF) / 16777216.0 and 0.0 <= f < 1.0


def test_ray_sphere_known_answer(oracle):
    # src/gpu/intersect.glsl:26-37 with Spheres[0] from src/gpu/scene.glsl:6 and the default camera ray
    sp = oracle.spheres_array()
    t = oracle.lib().orc_ray_sphere(oracle.f3((0, -0.001, 0.1)), oracle.f3((0, 1, 0)), C.byref(sp[0]))
    assert abs(t - 0.0022) < 1e-7
    # origin inside the sphere: the far root is never used => negative t (treated as a miss)
    t = oracle.lib().orc_ray_sphere(oracle.f3((0, 0.0017, 0.1)), oracle.f3((0, 1, 0)), C.byref(sp[0]))
    assert t < 0
    # no intersection
    assert oracle.lib().orc_ray_sphere(oracle.f3((0, -0.001, 0.2)), oracle.f3((0, 1, 0)), C.byref(sp[0])) == -1.0


def test_camera_matrices(oracle):
    L = oracle.lib()
    cam = oracle.default_camera(960, 540)
    pc, sc = oracle.constants(cam)
    P = pc.projection.numpy()
    h = np.float32(1.0) / np.tan(np.float32(0.5) * np.float32(oracle.deg(60)))
    assert np.isclose(P[1][1], h, rtol=1e-6) and np.isclose(P[0][0], h * 0.5625, rtol=1e-6)
    assert P[2][3] == 1.0 and P[3][2] == np.float32(0.001)
    assert np.count_nonzero(P) == 4
    V = pc.view.numpy()
    # yaw 90 deg, pitch 0: rows ~ (+x, +z, +y) with cos(pi/2 in fp32) = -4.37e-8 residue
    assert V[0][0] == 1.0 and V[1][2] == 1.0 and V[2][1] == 1.0
    assert abs(abs(V[0][2]) - 4.371139e-08) < 1e-12
    assert np.allclose(V[3][:3], [0.0, -0.1, 0.001], atol=1e-9)
    for M, Mi in ((pc.view, pc.invView), (pc.projection, pc.invProjection)):
        prod = oracle.Mat4()
        L.orc_mat_mul(C.byref(M), C.byref(Mi), C.byref(prod))
        assert np.abs(prod.numpy() - np.eye(4, dtype=np.float32)).max() < 1e-6
    assert list(sc.cameraPos) == pytest.approx([0.0, -0.001, 0.1])
    assert C.sizeof(pc) == 324 and C.sizeof(sc) == 272


def test_camera_controls(oracle):
    # src/gfx/camera.ixx:47-63
    L = oracle.lib()
    cam = oracle.default_camera()
    L.orc_camera_rotate(C.byref(cam), 256.0 * 2, 1e6)  # yaw 90 deg - 2 rad < 0 => wraps once by +360 deg
    assert cam.yaw == pytest.approx(oracle.deg(90) - 2.0 + oracle.deg(360), rel=1e-6)
    assert cam.pitch == pytest.approx(oracle.deg(89))
    cam = oracle.default_camera()
    before = list(cam.position)
    L.orc_camera_roam(C.byref(cam), oracle.f3((0.0, 0.0, 1.0)))  # forward = +y for yaw 90 deg
    assert cam.position[1] - before[1] == pytest.approx(8.0, rel=1e-6)
    assert abs(cam.position[0] - before[0]) < 1e-5 and abs(cam.position[2] - before[2]) < 1e-5


def test_storage_formats(oracle):
    L = oracle.lib()
    vals = np.concatenate([np.linspace(-70000, 70000, 2001), np.logspace(-9, 5, 500), [0.0, -0.0, 65504, 65519.9, 65520, 1e-8, 6e-8]]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([L.orc_f32_to_f16(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got, want)
    back = np.array([L.orc_f16_to_f32(int(b)) for b in range(0, 0x7C01)], np.float32)
    assert np.array_equal(back, np.arange(0, 0x7C01, dtype=np.uint16).view(np.float16).astype(np.float32))
    # B10G11R11: exact values survive, negatives clamp, inf saturates, rounding is to nearest
    rgb = (C.c_float * 3)(1.0, 0.5, 2.0)
    p = L.orc_pack_b10g11r11(rgb)
    out = (C.c_float * 3)()
    L.orc_unpack_b10g11r11(p, out)
    assert list(out) == [1.0, 0.5, 2.0]
    p = L.orc_pack_b10g11r11((C.c_float * 3)(-1.0, float("inf"), float("nan")))
    L.orc_unpack_b10g11r11(p, out)
    assert list(out) == [0.0, 65024.0, 0.0]
    p = L.orc_pack_b10g11r11((C.c_float * 3)(1.0 + 1 / 128 + 1e-6, 1.0 + 1 / 128 - 1e-6, 1.0 + 1 / 64 + 1e-6))
    L.orc_unpack_b10g11r11(p, out)
    assert list(out) == [1.0 + 1 / 64, 1.0, 1.0 + 1 / 32]
    assert [L.orc_unorm8(x) for x in (-1.0, 0.0, 0.5, 1.0, 2.0, float("nan"))] == [0, 0, 128, 255, 255, 0]


def test_blue_noise_fixture(blue_noise):
    # decoded RGBA8 of assets/blue_noise.png as loaded at src/gfx/renderer.ixx:101-110 (LCT_RGBA, 8)
    assert blue_noise.shape == (256, 256, 4)
    assert hashlib.sha256(blue_noise.tobytes()).hexdigest().startswith("cf9a7ffb")
    assert tuple(blue_noise[0, 0, :2]) == (2, 57)


def test_watertight_triangle(oracle):
    L = oracle.lib()
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    tri = [oracle.f3(p) for p in ((0, 1, 0), (1, 1, 0), (0, 1, 1))]
    assert L.orc_ray_triangle(oracle.f3((0.2, 0, 0.2)), oracle.f3((0, 1, 0)), *tri, C.byref(t), C.byref(u), C.byref(v)) == 1
    assert t.value == 1.0 and u.value == pytest.approx(0.2) and v.value == pytest.approx(0.2)
    # back face hits too; behind the origin does not
    assert L.orc_ray_triangle(oracle.f3((0.2, 2, 0.2)), oracle.f3((0, -1, 0)), *tri, C.byref(t), C.byref(u), C.byref(v)) == 1
    assert L.orc_ray_triangle(oracle.f3((0.2, 2, 0.2)), oracle.f3((0, 1, 0)), *tri, C.byref(t), C.byref(u), C.byref(v)) == 0
    # a ray through the shared edge / vertex of two triangles hits at least one of them (watertight)
    tri2 = [oracle.f3(p) for p in ((1, 1, 0), (1, 1, 1), (0, 1, 1))]
    rng = np.random.default_rng(0)
    for _ in range(2000):
        s = rng.uniform()
        target = np.array([1 - s, 1.0, s], np.float32)  # on the shared edge (1,1,0)-(0,1,1)
        o = rng.uniform(-1, 1, 3).astype(np.float32) * np.array([1, 0.5, 1], np.float32)
        d = target - o
        h1 = L.orc_ray_triangle(oracle.f3(o), oracle.f3(d), *tri, C.byref(t), C.byref(u), C.byref(v))
        h2 = L.orc_ray_triangle(oracle.f3(o), oracle.f3(d), *tri2, C.byref(t), C.byref(u), C.byref(v))
        assert h1 or h2


def test_oracle_bvh_equals_brute_force(oracle):
    from minotert_b200 import scenes
    pos, idx, alb, view = scenes.small_terrain()
    sc = oracle.Scene(pos, idx, alb)
    cam = oracle.make_camera(96, 64, view["position"], view["yaw_deg"], view["pitch_deg"])
    pc, _ = oracle.constants(cam)
    a = sc.primary(96, 64, pc, use_bvh=True)
    b = sc.primary(96, 64, pc, use_bvh=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    rng = np.random.default_rng(5)
    for _ in range(300):
        o = rng.uniform(pos.min(0), pos.max(0)).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        assert sc.closest_hit(o, d, True) == sc.closest_hit(o, d, False)


def test_sky_luts_sanity(oracle, sky_inputs):
    atmo, trans, multi, view = sky_inputs
    t = oracle.f16_to_f32(trans)
    assert t.shape == (64, 256, 4) and np.all(t[..., 3] == 1.0)
    assert np.all((t[..., :3] >= 0) & (t[..., :3] <= 1))
    # looking up from the ground: blue is attenuated more than red (Rayleigh)
    assert t[0, 0, 0] > t[0, 0, 1] > t[0, 0, 2] > 0.5
    m = oracle.f16_to_f32(multi)
    assert np.all(m[..., :3] >= 0) and m[16, 16, 2] > m[16, 16, 0]
    # sky colour straight up is blue-ish and finite; the sun direction adds the disc (> 1e4)
    out = (C.c_float * 3)()
    cam = oracle.default_camera()
    oracle.lib().orc_sky_color(C.byref(atmo), oracle._p(trans, C.c_uint16), oracle._p(view, C.c_uint32),
                               oracle.f3(cam.position[:]), oracle.f3((0.0, 0.01, 1.0)), out)
    assert out[2] > out[0] > 0 and out[2] < 10
    oracle.lib().orc_sky_color(C.byref(atmo), oracle._p(trans, C.c_uint16), oracle._p(view, C.c_uint32),
                               oracle.f3(cam.position[:]), oracle.f3(oracle.SUN_DIRECTION), out)
    assert min(out) > 1e4


def test_tonemap_pixel_properties(oracle):
    L = oracle.lib()
    out = (C.c_float * 3)()
    par = (C.c_float * 8)(*oracle.AMD_DEFAULT)
    # AMD curve maps midIn -> midOut (0.18 -> 0.18) before the sRGB encode: srgb(0.18) = 0.4614
    L.orc_tonemap_pixel(5, oracle.f3((0.18, 0.18, 0.18)), 1.0, par, out)
    assert out[0] == pytest.approx(0.46135, abs=2e-4) and out[0] == out[1] == out[2]
    # hdrMax maps to 1.0
    L.orc_tonemap_pixel(5, oracle.f3((16.0, 16.0, 16.0)), 1.0, par, out)
    assert out[0] == pytest.approx(1.0, abs=1e-4)
    # linear: pure sRGB encode
    L.orc_tonemap_pixel(0, oracle.f3((0.001, 0.5, 1.0)), 1.0, par, out)
    assert out[0] == pytest.approx(0.01292) and out[1] == pytest.approx(0.735357, abs=1e-5) and out[2] == pytest.approx(1.0, abs=1e-6)


def test_golden_fixtures_reproduce(oracle, blue_noise):
    """The oracle still produces, bit for bit, the fixtures committed in tests/golden/oracle_v1.npz."""
    from minotert_b200 import scenes
    g = np.load(GOLD)
    L = oracle.lib()
    s = C.c_uint32(3)
    assert np.array_equal(g["pcg_seed3"], np.array([L.orc_pcg(C.byref(s)) for _ in range(16)], np.uint32))
    cam = oracle.default_camera(96, 54)
    pc, sc = oracle.constants(cam, frame=1)
    assert bytes(pc) == g["primary_constants_96x54"].tobytes()
    assert bytes(sc) == g["secondary_constants_96x54"].tobytes()
    atmo = oracle.earth()
    assert bytes(atmo) == g["atmosphere_earth"].tobytes()
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    assert np.array_equal(trans, g["sky_transmittance"]) and np.array_equal(multi, g["sky_multiscattering"])
    assert np.array_equal(view, g["sky_view"])
    sp = oracle.spheres_array()
    vis, depth, normal, motion = oracle.primary_spheres(96, 54, pc, sp)
    assert np.array_equal(vis, g["spheres_vis"]) and np.array_equal(depth, g["spheres_depth"])
    assert np.array_equal(normal, g["spheres_normal"]) and np.array_equal(motion, g["spheres_motion"])
    c16, c32, rays = oracle.secondary_spheres(96, 54, sc, sp, vis, depth, normal, blue_noise, atmo, trans, view, 8, 8)
    assert np.array_equal(c16, g["spheres_color16"]) and rays == int(g["spheres_rays"][0])
    for mode, params in [("linear", (0.0,)), ("reinhard", (8.0,)), ("hable", (0.0,)), ("aces", (0.0,)),
                         ("uchimura", oracle.UCHIMURA_DEFAULT), ("amd", oracle.AMD_DEFAULT)]:
        assert np.array_equal(oracle.tonemap(mode, c16, 1.0, params), g["spheres_ldr_" + mode]), mode
    pos, idx, alb, v = scenes.cornell()
    sha = hashlib.sha256(pos.tobytes() + idx.tobytes() + alb.tobytes()).digest()
    assert sha == g["cornell_positions_sha"].tobytes(), "procedural scene generator changed"
    camc = oracle.make_camera(64, 64, v["position"], v["yaw_deg"], v["pitch_deg"])
    pcc, scc = oracle.constants(camc, frame=1)
    tr2, mu2, vw2 = oracle.sky_luts(atmo, camc.position[:])
    scn = oracle.Scene(pos, idx, alb)
    acc, cvis, crays = scn.render(64, 64, pcc, scc, blue_noise, atmo, tr2, vw2, 2, 2, use_bvh=True)
    assert np.array_equal(cvis, g["cornell_vis"]) and np.array_equal(acc, g["cornell_accum"])
    assert list(crays) == list(g["cornell_rays"])


def test_golden_fixtures_v2_reproduce(oracle):
    """The stages after the path tracer (bilateral denoiser, temporal reprojection): the oracle still produces, bit
    for bit, tests/golden/oracle_v2.npz (generated by tests/golden/make_golden_v2.py, whose compute() is reused here)."""
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_v2", os.path.join(here, "make_golden_v2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.compute()
    g = np.load(os.path.join(here, "oracle_v2.npz"))
    assert sorted(got) == sorted(g.files)
    for k in g.files:
        assert np.array_equal(got[k].view(np.uint8), g[k].view(np.uint8)), k
    # the fixture exercises the history: a third of the pixels (the sphere hits) carry more than one frame
    assert (g["temporal_count"] > 1).mean() > 0.3 and g["temporal_count"].max() > 3.0
