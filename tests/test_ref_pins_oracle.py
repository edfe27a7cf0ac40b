"""The pin of the CPU oracle: oracle/minote_oracle.c (the restatement every GPU parity test compares against) must
reproduce, BIT FOR BIT, what the reference's own GLSL shaders compute.

Two layers:
  * golden vectors tests/golden/ref_v1.npz, generated from oracle/_ref (the reference's src/gpu/*.comp|*.glsl compiled
    as C++, oracle/ref/) by tests/golden/make_golden_ref.py -- always run, need nothing but the committed file;
  * live runs of oracle/_ref next to the restatement on the reference's full default frame (960x540, 8 spp x 8 bounces),
    the three sky LUTs, all six tonemappers, the bilateral denoiser and a 64 k-direction skyColor sweep -- run where
    the library exists (built here from /root/reference; the prebuilt .so travels to the GPU box).
Both sides are CPU code on glibc libm without FMA contraction, and the shim fixes every implementation-defined Vulkan
behaviour the way the restatement documents, so equality is exact: ids/integers AND floats.  The only values excluded
are depth and motion of primary MISS pixels, which are undefined in the reference (primaryRay.comp:33-34 reads an
uninitialised t and Spheres[-1u]).
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_v1.npz")
TONEMAPS = [("linear", ()), ("reinhard", (16.0,)), ("hable", ()), ("aces", ()), ("uchimura", O.UCHIMURA_DEFAULT),
            ("amd", O.AMD_DEFAULT)]
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")


def pad8(par):
    return tuple(par) + (0.0,) * (8 - len(par))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def struct_from(cls, arr):
    return cls.from_buffer_copy(arr.tobytes())


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


# ------------------------------------------------------------------ oracle vs the committed golden vectors

def test_golden_units(gold):
    L = O.lib()
    s = C.c_uint32(3)
    assert [L.orc_pcg(C.byref(s)) for _ in range(64)] == gold["pcg_seed3"].tolist()
    s = C.c_uint32(99)
    f = np.array([L.orc_random_float(C.byref(s)) for _ in range(64)], np.float32)
    assert np.array_equal(bits(f), bits(gold["random_float_seed99"]))
    for (rx, ry), want in zip(gold["sphere_point_in"], gold["sphere_point_out"]):
        p = (C.c_float * 3)()
        L.orc_random_sphere_point(float(rx), float(ry), p)
        assert np.array_equal(bits(np.array(p[:], np.float32)), bits(want))
    # the scene the shaders carry == the scene the tests upload (src/gpu/scene.glsl:5-11)
    mine = np.array([list(c) + [r] + list(a) for c, r, a in O.REFERENCE_SPHERES], np.float32)
    assert np.array_equal(bits(mine), bits(gold["scene_spheres"]))
    sp = O.spheres_array()
    for o, d, ts in zip(gold["ray_sphere_o"], gold["ray_sphere_d"], gold["ray_sphere_t"]):
        got = np.array([L.orc_ray_sphere(O.f3(o), O.f3(d), C.byref(sp[k])) for k in range(len(sp))], np.float32)
        assert np.array_equal(bits(got), bits(ts))
    assert bytes(O.earth()) == gold["atmosphere_earth"].tobytes()


@pytest.mark.parametrize("pose", ["default", "moved"])
def test_golden_camera_constants(gold, pose):
    """Row a1: Camera::view/projection + look/perspective/inverse + the constant-block fill, restated in the oracle,
    against the blocks the reference's own host code produced."""
    raw = gold[pose + "_camera"].tobytes()
    n = C.sizeof(O.Camera)
    cam, prev = O.Camera.from_buffer_copy(raw[:n]), O.Camera.from_buffer_copy(raw[n:])
    frame = int(struct_from(O.PrimaryConstants, gold[pose + "_primary_constants"]).frameCounter)
    pc, sc = O.constants(cam, prev, frame=frame)
    assert bytes(pc) == gold[pose + "_primary_constants"].tobytes()
    assert bytes(sc) == gold[pose + "_secondary_constants"].tobytes()


@pytest.mark.parametrize("pose", ["default", "moved"])
def test_golden_frames(gold, pose, blue_noise):
    pc = struct_from(O.PrimaryConstants, gold[pose + "_primary_constants"])
    sc = struct_from(O.SecondaryConstants, gold[pose + "_secondary_constants"])
    h, w = gold[pose + "_vis"].shape
    atmo = O.earth()
    trans, multi, view = O.sky_luts(atmo, sc.cameraPos[:])
    assert np.array_equal(trans, gold["trans"]) and np.array_equal(multi, gold["multi"])
    assert np.array_equal(view, gold[pose + "_view"])
    sp = O.spheres_array()
    vis, depth, normal, motion = O.primary_spheres(w, h, pc, sp)
    hit = vis != O.NONE_ID
    assert 0.2 < hit.mean() < 0.9
    assert np.array_equal(vis, gold[pose + "_vis"])
    assert np.array_equal(normal[..., :3], gold[pose + "_normal"][..., :3])
    assert np.array_equal(depth[hit], gold[pose + "_depth"][hit])
    assert np.array_equal(motion[hit], gold[pose + "_motion"][hit])
    # the secondary pass reads the G-buffer the REFERENCE's primary pass wrote (miss pixels: only id + normal are read)
    c16, _, _ = O.secondary_spheres(w, h, sc, sp, gold[pose + "_vis"], gold[pose + "_depth"], gold[pose + "_normal"],
                                    blue_noise, atmo, gold["trans"], gold[pose + "_view"], 8, 8)
    assert np.array_equal(c16, gold[pose + "_color16"])
    for mode, par in TONEMAPS:
        assert np.array_equal(O.tonemap(mode, gold[pose + "_color16"], 1.0, pad8(par)), gold[pose + "_ldr_" + mode]), mode
    den = O.denoise_bilateral(gold[pose + "_color16"], gold[pose + "_depth"], gold[pose + "_normal"],
                              frame=int(sc.frameCounter))
    assert np.array_equal(den, gold[pose + "_denoised"])


def test_golden_sky_color_and_tonemap_sweep(gold):
    atmo = O.earth()
    sc = struct_from(O.SecondaryConstants, gold["default_secondary_constants"])
    got = O.sky_color(atmo, gold["trans"], gold["default_view"], sc.cameraPos[:], gold["sky_dirs"])
    o3 = (C.c_float * 3)()   # the single-direction entry point agrees with the batch one
    O.lib().orc_sky_color(C.byref(atmo), O._p(gold["trans"], C.c_uint16), O._p(gold["default_view"], C.c_uint32),
                          O.f3(sc.cameraPos[:]), O.f3(gold["sky_dirs"][17]), o3)
    assert np.array_equal(bits(np.array(o3[:], np.float32)), bits(got[17]))
    assert np.array_equal(bits(got), bits(gold["sky_colors"]))
    assert (gold["sky_colors"].max(1) > 1000).sum() > 100, "the sweep must cover the sun disc"
    for mode, par in TONEMAPS:
        assert np.array_equal(O.tonemap(mode, gold["hdr_sweep16"], 0.37, pad8(par)), gold["sweep_ldr_" + mode]), mode


# ------------------------------------------------------------------ oracle vs oracle/_ref run live

@needs_ref
def test_golden_file_is_what_ref_produces_today(gold):
    """The committed vectors are not stale: regenerate two of them from the live library."""
    pc = struct_from(O.PrimaryConstants, gold["default_primary_constants"])
    h, w = gold["default_vis"].shape
    vis, depth, normal, motion = R.primary(w, h, pc)
    assert np.array_equal(vis, gold["default_vis"]) and np.array_equal(normal, gold["default_normal"])
    s = C.c_uint32(3)
    assert [R.lib().ref_pcg(C.byref(s)) for _ in range(64)] == gold["pcg_seed3"].tolist()


@needs_ref
def test_host_matrices_and_camera_bit_exact():
    """2000 random cameras through the reference's own math.ixx/camera.ixx and through the restatement; plus the
    freecam operations rotate / shift / roam (camera.ixx:47-63) and the _deg literal."""
    rng = np.random.default_rng(0)
    L = R.lib()
    for d in np.arange(-400, 400, 0.5):
        assert L.ref_deg(d) == O.deg(d)
    for i in range(2000):
        w, h = int(rng.integers(16, 4000)), int(rng.integers(16, 2200))
        cam = O.make_camera(w, h, tuple(rng.normal(size=3) * [0.01, 0.01, 0.2]), float(rng.uniform(0, 360)),
                            float(rng.uniform(-89, 89)), float(rng.uniform(20, 100)), float(10 ** rng.uniform(-4, -1)))
        prev = O.make_camera(w, h, tuple(rng.normal(size=3) * 0.01), float(rng.uniform(0, 360)), float(rng.uniform(-89, 89)))
        if i == 0:
            cam = prev = O.default_camera()
        po, so = O.constants(cam, prev, frame=i + 1)
        pr, sr = R.constants(cam, prev, frame=i + 1)
        assert bytes(po) == bytes(pr) and bytes(so) == bytes(sr), i
        if i < 300:
            a, b = O.Camera.from_buffer_copy(bytes(cam)), O.Camera.from_buffer_copy(bytes(cam))
            hv = rng.normal(size=2) * 200
            O.lib().orc_camera_rotate(C.byref(a), float(hv[0]), float(hv[1]))
            L.ref_camera_rotate(C.byref(b), float(hv[0]), float(hv[1]))
            dv = O.f3(rng.normal(size=3) * 1e-3)
            O.lib().orc_camera_shift(C.byref(a), dv)
            L.ref_camera_shift(C.byref(b), dv)
            O.lib().orc_camera_roam(C.byref(a), dv)
            L.ref_camera_roam(C.byref(b), dv)
            assert bytes(a) == bytes(b), i
    assert bytes(O.earth()) == bytes(R.earth())


@needs_ref
def test_reference_default_frame_bit_exact(blue_noise):
    """src/main.cpp:24 window, src/app.ixx:20-32 camera, frame counter 1 (renderer.ixx:52), 8 spp x 8 bounces."""
    w, h = 960, 540
    cam = O.default_camera(w, h)
    pc, sc = O.constants(cam, frame=1)
    atmo = O.earth()
    lo, lr = O.sky_luts(atmo, cam.position[:]), R.sky_luts(atmo, cam.position[:])
    for a, b, name in zip(lo, lr, ("transmittance", "multi-scattering", "sky view")):
        assert np.array_equal(a, b), name
    sp = O.spheres_array()
    vo, vr = O.primary_spheres(w, h, pc, sp), R.primary(w, h, pc)
    hit = vo[0] != O.NONE_ID
    assert np.array_equal(vo[0], vr[0]) and 0.3 < hit.mean() < 0.7
    assert np.array_equal(vo[2][..., :3], vr[2][..., :3])
    assert np.array_equal(vo[1][hit], vr[1][hit]) and np.array_equal(vo[3][hit], vr[3][hit])
    co, _, rays = O.secondary_spheres(w, h, sc, sp, vr[0], vr[1], vr[2], blue_noise, atmo, lr[0], lr[2], 8, 8)
    cr = R.secondary(w, h, sc, vr[0], vr[1], vr[2], blue_noise, atmo, lr[0], lr[2])
    assert np.array_equal(co, cr)
    assert rays > 8 * hit.sum()
    for mode, par in TONEMAPS:
        assert np.array_equal(O.tonemap(mode, cr, 1.0, pad8(par)), R.tonemap(mode, cr, 1.0, pad8(par))), mode
    # the denoiser on a 240-row band (the full frame costs ~1 s per side; the band holds spheres, ground and sky)
    band = slice(150, 390)
    do = O.denoise_bilateral(cr[band], vr[1][band], vr[2][band], frame=1)
    dr = R.denoise_bilateral(cr[band], vr[1][band], vr[2][band], frame=1)
    assert np.array_equal(do, dr)


@needs_ref
def test_sky_color_sweep_bit_exact(sky_inputs):
    atmo, trans, multi, view = sky_inputs
    rng = np.random.default_rng(5)
    n = 40000
    d = rng.normal(size=(n, 3))
    sun = np.array(O.SUN_DIRECTION)
    k = 12000
    ang = np.deg2rad(rng.uniform(0.0, 0.6, k))
    phi = rng.uniform(0, 2 * np.pi, k)
    t1 = np.cross(sun, [0, 0, 1.0])
    t1 /= np.linalg.norm(t1)
    t2 = np.cross(sun, t1)
    rim = np.cos(ang)[:, None] * sun + np.sin(ang)[:, None] * (np.cos(phi)[:, None] * t1 + np.sin(phi)[:, None] * t2)
    hz = rng.normal(size=(k, 3))
    hz[:, 2] = rng.uniform(-0.02, 0.02, k)
    dirs = np.concatenate([d, rim, hz, [[1e-7, 0, 1], [0, 1e-6, -1]]])
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    pos = O.default_camera().position[:]
    ref = R.sky_color(atmo, trans, view, pos, dirs)
    got = O.sky_color(atmo, trans, view, pos, dirs)
    assert np.array_equal(bits(got), bits(ref))
    assert (ref.max(1) > 1000).sum() > 1000


@needs_ref
def test_rng_and_bounce_stream_bit_exact():
    """pcg -> randomFloat -> randomSpherePoint chained exactly as secondaryRays.comp:60-72 does, 10^5 draws."""
    LO, LR = O.lib(), R.lib()
    so, sr = C.c_uint32((5 << 1) | 1), C.c_uint32((5 << 1) | 1)
    rot = (np.float32(57 / 255.0), np.float32(2 / 255.0))
    for _ in range(100000 // 2):
        a = [LO.orc_random_float(C.byref(so)) for _ in range(2)]
        b = [LR.ref_random_float(C.byref(sr)) for _ in range(2)]
        assert a == b and so.value == sr.value
    po, pr = (C.c_float * 3)(), (C.c_float * 3)()
    rng = np.random.default_rng(1)
    for rx, ry in rng.uniform(0, 1, (20000, 2)).astype(np.float32):
        x = np.float32(np.float32(rx + rot[0]) % np.float32(1.0)) * np.float32(2) - np.float32(1)
        y = np.float32(np.float32(ry + rot[1]) % np.float32(1.0)) * np.float32(2) - np.float32(1)
        LO.orc_random_sphere_point(float(x), float(y), po)
        LR.ref_random_sphere_point(float(x), float(y), pr)
        assert po[:] == pr[:]
