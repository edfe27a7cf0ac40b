"""ctypes binding of the CPU oracle (oracle/libminote_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libminote_oracle.so")

NONE_ID = 0xFFFFFFFF
TRANS_W, TRANS_H, MULTI_W, MULTI_H, VIEW_W, VIEW_H = 256, 64, 32, 32, 192, 108
TONEMAP = {"linear": 0, "reinhard": 1, "hable": 2, "aces": 3, "uchimura": 4, "amd": 5}
AMD_DEFAULT = (16.0, 2.0, 1.0, 0.18, 0.18)          # src/gfx/modules/tonemapper.ixx:46-54
UCHIMURA_DEFAULT = (1.0, 1.0, 0.22, 0.4, 1.33, 0.0)  # src/gfx/modules/tonemapper.ixx:27-36
SUN_DIRECTION = (-0.435286462, 0.818654716, 0.374606609)  # src/gfx/modules/sky.ixx:193
SUN_ILLUMINANCE = (8.0, 8.0, 8.0)                          # src/gfx/modules/sky.ixx:194


class Camera(C.Structure):
    _fields_ = [("viewport", C.c_uint32 * 2), ("verticalFov", C.c_float), ("nearPlane", C.c_float),
                ("position", C.c_float * 3), ("yaw", C.c_float), ("pitch", C.c_float),
                ("lookSpeed", C.c_float), ("moveSpeed", C.c_float)]


class WideNode(C.Structure):
    """orc_wide_node: the quantised 8-wide BVH node of the product, as the oracle restates it."""
    _fields_ = [("w", C.c_uint32 * 16)]


class Mat4(C.Structure):
    _fields_ = [("m", (C.c_float * 4) * 4)]

    def numpy(self):
        return np.ctypeslib.as_array(self.m).copy()  # [col][row]


class PrimaryConstants(C.Structure):
    _fields_ = [("view", Mat4), ("projection", Mat4), ("invView", Mat4), ("invProjection", Mat4),
                ("prevView", Mat4), ("frameCounter", C.c_uint32)]


class SecondaryConstants(C.Structure):
    _fields_ = [("view", Mat4), ("projection", Mat4), ("invView", Mat4), ("invProjection", Mat4),
                ("cameraPos", C.c_float * 3), ("frameCounter", C.c_uint32)]


class Sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("albedo", C.c_float * 3)]


class AtmosphereParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "bottomRadius", "topRadius", "rayleighDensityExpScale", "_pad0")] + [
        ("rayleighScattering", C.c_float * 3), ("mieDensityExpScale", C.c_float),
        ("mieScattering", C.c_float * 3), ("_pad1", C.c_float),
        ("mieExtinction", C.c_float * 3), ("_pad2", C.c_float),
        ("mieAbsorption", C.c_float * 3), ("miePhaseG", C.c_float),
        ("absorptionDensity0LayerWidth", C.c_float), ("absorptionDensity0ConstantTerm", C.c_float),
        ("absorptionDensity0LinearTerm", C.c_float), ("absorptionDensity1ConstantTerm", C.c_float),
        ("absorptionDensity1LinearTerm", C.c_float), ("_pad3", C.c_float), ("_pad4", C.c_float),
        ("_pad5", C.c_float),
        ("absorptionExtinction", C.c_float * 3), ("_pad6", C.c_float),
        ("groundAlbedo", C.c_float * 3), ("_pad7", C.c_float)]


assert C.sizeof(AtmosphereParams) == 144
assert C.sizeof(PrimaryConstants) == 324
assert C.sizeof(SecondaryConstants) == 272
assert C.sizeof(Sphere) == 28

# src/gpu/scene.glsl:5-11
REFERENCE_SPHERES = [
    ((0.0000, 0.0017, 0.10000), 0.00050, (0.2, 0.7, 0.0)),
    ((-0.0008, 0.0012, 0.09983), 0.00033, (0.0, 0.2, 0.7)),
    ((0.0008, 0.0012, 0.09983), 0.00033, (0.7, 0.0, 0.2)),
    ((0.0000, 0.0008, 0.09975), 0.00025, (1.0, 1.0, 1.0)),
    ((0.0000, 0.0010, -0.00050), 0.10000, (0.5, 0.5, 0.5)),
]


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("minote_oracle.c", "minote_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    f32p, u32p, u16p, u8p, u64p = (C.POINTER(t) for t in (C.c_float, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint64))
    sig = {
        "orc_f32_to_f16": (C.c_uint16, [C.c_float]),
        "orc_f16_to_f32": (C.c_float, [C.c_uint16]),
        "orc_pack_b10g11r11": (C.c_uint32, [f32p]),
        "orc_unpack_b10g11r11": (None, [C.c_uint32, f32p]),
        "orc_unorm8": (C.c_uint8, [C.c_float]),
        "orc_camera_direction": (None, [C.POINTER(Camera), f32p]),
        "orc_camera_view": (None, [C.POINTER(Camera), C.POINTER(Mat4)]),
        "orc_camera_projection": (None, [C.POINTER(Camera), C.POINTER(Mat4)]),
        "orc_perspective": (None, [C.c_float, C.c_float, C.c_float, C.POINTER(Mat4)]),
        "orc_inverse": (None, [C.POINTER(Mat4), C.POINTER(Mat4)]),
        "orc_mat_mul": (None, [C.POINTER(Mat4), C.POINTER(Mat4), C.POINTER(Mat4)]),
        "orc_primary_constants_fill": (None, [C.POINTER(Camera), C.POINTER(Camera), C.c_uint32, C.POINTER(PrimaryConstants)]),
        "orc_secondary_constants_fill": (None, [C.POINTER(Camera), C.c_uint32, C.POINTER(SecondaryConstants)]),
        "orc_atmosphere_earth": (None, [C.POINTER(AtmosphereParams)]),
        "orc_camera_rotate": (None, [C.POINTER(Camera), C.c_float, C.c_float]),
        "orc_camera_shift": (None, [C.POINTER(Camera), f32p]),
        "orc_camera_roam": (None, [C.POINTER(Camera), f32p]),
        "orc_pcg": (C.c_uint32, [u32p]),
        "orc_random_float": (C.c_float, [u32p]),
        "orc_random_sphere_point": (None, [C.c_float, C.c_float, f32p]),
        "orc_ray_sphere": (C.c_float, [f32p, f32p, C.POINTER(Sphere)]),
        "orc_ray_triangle": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p]),
        "orc_ray_gen": (None, [C.POINTER(Mat4), C.POINTER(Mat4), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p]),
        "orc_gen_transmittance": (None, [C.POINTER(AtmosphereParams), u16p]),
        "orc_gen_multiscattering": (None, [C.POINTER(AtmosphereParams), u16p, u16p]),
        "orc_gen_sky_view": (None, [C.POINTER(AtmosphereParams), u16p, u16p, f32p, f32p, f32p, u32p]),
        "orc_sky_color": (None, [C.POINTER(AtmosphereParams), u16p, u32p, f32p, f32p, f32p]),
        "orc_sky_color_batch": (None, [C.POINTER(AtmosphereParams), u16p, u32p, f32p, C.c_uint32, f32p, f32p]),
        "orc_primary_rays_spheres": (None, [C.c_uint32, C.c_uint32, C.POINTER(PrimaryConstants), C.POINTER(Sphere), C.c_uint32, u32p, u16p, u16p, u16p]),
        "orc_secondary_rays_spheres": (None, [C.c_uint32, C.c_uint32, C.POINTER(SecondaryConstants), C.POINTER(Sphere), C.c_uint32, u32p, u16p, u16p, u8p, C.c_uint32, C.c_uint32, C.POINTER(AtmosphereParams), u16p, u32p, C.c_uint32, C.c_uint32, u16p, f32p, u64p]),
        "orc_denoise_bilateral": (None, [C.c_uint32, C.c_uint32, u16p, u16p, u16p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, u8p]),
        "orc_ray_triangle_batch": (None, [C.c_uint32, f32p, f32p, f32p, f32p, f32p, u8p, f32p]),
        "orc_wide_node_quantize": (None, [f32p, f32p, C.c_uint32, C.POINTER(WideNode)]),
        "orc_wide_node_test": (None, [C.POINTER(WideNode), C.c_uint32, f32p, f32p, f32p, u32p]),
        "orc_temporal_accumulate": (None, [C.c_uint32, C.c_uint32, f32p, u32p, u16p, C.c_int, f32p, f32p, u32p, C.c_float, f32p, f32p]),
        "orc_tonemap": (None, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_float, f32p, u8p]),
        "orc_tonemap_pixel": (None, [C.c_int, f32p, C.c_float, f32p, f32p]),
        "orc_scene_create": (C.c_void_p, [f32p, C.c_uint32, u32p, C.c_uint32, f32p]),
        "orc_scene_destroy": (None, [C.c_void_p]),
        "orc_scene_closest_hit": (C.c_uint32, [C.c_void_p, f32p, f32p, C.c_int, f32p, f32p, f32p]),
        "orc_primary_rays_tris": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(PrimaryConstants), C.c_int, C.c_uint32, C.c_uint32, u32p, u16p, u16p, u16p, f32p]),
        "orc_render_tris": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(PrimaryConstants), C.POINTER(SecondaryConstants), u8p, C.c_uint32, C.c_uint32, C.POINTER(AtmosphereParams), u16p, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, f32p, u32p, u64p]),
        "orc_render_tris_ext": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(PrimaryConstants), C.POINTER(SecondaryConstants), u8p, C.c_uint32, C.c_uint32, C.POINTER(AtmosphereParams), u16p, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, f32p, u32p, u64p, C.c_uint32, u16p]),
        "orc_gen_aerial_perspective": (None, [C.POINTER(AtmosphereParams), u16p, u16p, C.c_void_p, C.c_void_p, f32p, f32p, f32p, u16p]),
        "orc_aerial_perspective_lookup": (None, [u16p, C.c_float, C.c_float, C.c_float, f32p]),
        "orc_nee_sun_sample": (None, [C.c_float, C.c_float, f32p, f32p]),
        "orc_sun_centre_radiance": (None, [C.POINTER(AtmosphereParams), u16p, u32p, f32p, f32p]),
        "orc_resolve": (None, [C.c_uint32, f32p, f32p]),
        "orc_num_threads": (C.c_int, []),
        "orc_set_num_threads": (None, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


# ------------------------------------------------------------------ helpers


def deg(d):
    """The reference's _deg literal: radians<double, Prec = float>, i.e. fp32 arithmetic (src/stx/math.ixx:27,884)."""
    return float(np.float32(d) * (np.float32(np.pi) * np.float32(2.0)) / np.float32(360.0))


def default_camera(w=960, h=540):
    """src/app.ixx:20-32 with the window size of src/main.cpp:24."""
    c = Camera()
    c.viewport[0], c.viewport[1] = w, h
    c.verticalFov = deg(60)
    c.nearPlane = 0.001
    c.position[:] = (0.0, -0.001, 0.1)
    c.yaw = deg(90)
    c.pitch = 0.0
    c.lookSpeed = 1.0 / 256.0
    c.moveSpeed = 8.0
    return c


def make_camera(w, h, position, yaw_deg, pitch_deg, vfov_deg=60.0, near=0.001):
    c = default_camera(w, h)
    c.position[:] = position
    c.yaw = deg(yaw_deg)
    c.pitch = deg(pitch_deg)
    c.verticalFov = deg(vfov_deg)
    c.nearPlane = near
    return c


def constants(cam, prev=None, frame=1):
    pc, sc = PrimaryConstants(), SecondaryConstants()
    lib().orc_primary_constants_fill(C.byref(cam), C.byref(prev if prev is not None else cam), frame, C.byref(pc))
    lib().orc_secondary_constants_fill(C.byref(cam), frame, C.byref(sc))
    return pc, sc


def earth():
    p = AtmosphereParams()
    lib().orc_atmosphere_earth(C.byref(p))
    return p


def spheres_array(spheres=REFERENCE_SPHERES):
    arr = (Sphere * len(spheres))()
    for i, (c, r, a) in enumerate(spheres):
        arr[i].center[:] = c
        arr[i].radius = r
        arr[i].albedo[:] = a
    return arr


def f3(v):
    return (C.c_float * 3)(*v)


def sky_luts(atmo, probe_pos, sun_dir=SUN_DIRECTION, sun_ill=SUN_ILLUMINANCE):
    L = lib()
    trans = np.zeros((TRANS_H, TRANS_W, 4), np.uint16)
    multi = np.zeros((MULTI_H, MULTI_W, 4), np.uint16)
    view = np.zeros((VIEW_H, VIEW_W), np.uint32)
    L.orc_gen_transmittance(C.byref(atmo), _p(trans, C.c_uint16))
    L.orc_gen_multiscattering(C.byref(atmo), _p(trans, C.c_uint16), _p(multi, C.c_uint16))
    L.orc_gen_sky_view(C.byref(atmo), _p(trans, C.c_uint16), _p(multi, C.c_uint16), f3(probe_pos), f3(sun_dir),
                       f3(sun_ill), _p(view, C.c_uint32))
    return trans, multi, view


def sky_color(atmo, trans, view, camera_pos, dirs):
    """skyColor() of secondaryRays.comp:36-58 for an (n, 3) array of directions."""
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros_like(d)
    lib().orc_sky_color_batch(C.byref(atmo), _p(trans, C.c_uint16), _p(view, C.c_uint32), f3(camera_pos), d.shape[0],
                              _p(d, C.c_float), _p(out, C.c_float))
    return out


def load_blue_noise():
    from PIL import Image
    im = Image.open(os.path.join(ROOT, "assets", "blue_noise.png")).convert("RGBA")
    return np.ascontiguousarray(np.array(im, dtype=np.uint8))


def primary_spheres(w, h, pc, spheres):
    vis = np.zeros((h, w), np.uint32)
    depth = np.zeros((h, w), np.uint16)
    normal = np.zeros((h, w, 4), np.uint16)
    motion = np.zeros((h, w, 2), np.uint16)
    lib().orc_primary_rays_spheres(w, h, C.byref(pc), spheres, len(spheres), _p(vis, C.c_uint32),
                                   _p(depth, C.c_uint16), _p(normal, C.c_uint16), _p(motion, C.c_uint16))
    return vis, depth, normal, motion


def secondary_spheres(w, h, sc, spheres, vis, depth, normal, bn, atmo, trans, view, spp=8, bounces=8):
    c16 = np.zeros((h, w, 4), np.uint16)
    c32 = np.zeros((h, w, 4), np.float32)
    rays = C.c_uint64(0)
    lib().orc_secondary_rays_spheres(w, h, C.byref(sc), spheres, len(spheres), _p(vis, C.c_uint32),
                                     _p(depth, C.c_uint16), _p(normal, C.c_uint16), _p(bn, C.c_uint8),
                                     bn.shape[1], bn.shape[0], C.byref(atmo), _p(trans, C.c_uint16),
                                     _p(view, C.c_uint32), spp, bounces, _p(c16, C.c_uint16), _p(c32, C.c_float),
                                     C.byref(rays))
    return c16, c32, rays.value


def tonemap(mode, src, exposure=1.0, params=AMD_DEFAULT):
    h, w = src.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    src = np.ascontiguousarray(src)
    par = (C.c_float * 8)(*params)
    lib().orc_tonemap(TONEMAP[mode] if isinstance(mode, str) else mode, w, h, src.ctypes.data_as(C.c_void_p),
                      {np.dtype(np.uint16): 1, np.dtype(np.uint8): 2}.get(src.dtype, 0), exposure, par,
                      _p(out, C.c_uint8))
    return out


BILATERAL_DEFAULT = (5.0, 2.0, 0.12)  # sigma, kSigma, threshold: src/gfx/modules/denoiser.ixx:27-33


def denoise_bilateral(color16, depth16, normal16, params=BILATERAL_DEFAULT, near=0.001, frame=1):
    """Denoiser::bilateral (denoiser.ixx:36-97): RGBA16F colour + R16F depth + RGBA16F normal -> RGBA8."""
    h, w = depth16.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    c, d, n = (np.ascontiguousarray(a, np.uint16) for a in (color16, depth16, normal16))
    lib().orc_denoise_bilateral(w, h, _p(c, C.c_uint16), _p(d, C.c_uint16), _p(n, C.c_uint16), params[0], params[1],
                                params[2], near, frame, _p(out, C.c_uint8))
    return out


def ray_triangle_batch(o, d, v0, v1, v2):
    """The watertight fp32 ray/triangle test (row n4) for n rays against one triangle -> (hit bool[n], t float32[n])."""
    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    n = o.shape[0]
    hit = np.zeros(n, np.uint8)
    t = np.zeros(n, np.float32)
    lib().orc_ray_triangle_batch(n, _p(o, C.c_float), _p(d, C.c_float), f3(v0), f3(v1), f3(v2), _p(hit, C.c_uint8), _p(t, C.c_float))
    return hit.astype(bool), t


def wide_node_quantize(lo, hi, present=0xFF):
    """lo, hi: (8, 3) float32 child boxes -> WideNode (bvh_build.cu k_emit_nodes restated)."""
    lo = np.ascontiguousarray(lo, np.float32).reshape(8, 3)
    hi = np.ascontiguousarray(hi, np.float32).reshape(8, 3)
    node = WideNode()
    lib().orc_wide_node_quantize(_p(lo, C.c_float), _p(hi, C.c_float), present, C.byref(node))
    return node


def wide_node_test(node, o, d, t_best=None):
    """8-bit hit masks of the node's children for rays (o, d): (n, 3) float32 each (trace.cuh lane_node_step restated)."""
    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    n = o.shape[0]
    tb = np.full(n, 3.0e38, np.float32) if t_best is None else np.ascontiguousarray(t_best, np.float32)
    out = np.zeros(n, np.uint32)
    lib().orc_wide_node_test(C.byref(node), n, _p(o, C.c_float), _p(d, C.c_float), _p(tb, C.c_float), _p(out, C.c_uint32))
    return out


def temporal_accumulate(accum, vis, motion16, history=None, max_history=32.0):
    """orc_temporal_accumulate: history = (rgba, count, vis) of the previous call or None.  Returns (rgba, count, vis)."""
    h, w = vis.shape[:2]
    a = np.ascontiguousarray(accum, np.float32)
    v = np.ascontiguousarray(vis, np.uint32)
    m = np.ascontiguousarray(motion16, np.uint16)
    out = np.zeros((h, w, 4), np.float32)
    cnt = np.zeros((h, w), np.float32)
    if history is None:
        hr, hc, hv = out, cnt, v
    else:
        hr, hc, hv = (np.ascontiguousarray(history[0], np.float32), np.ascontiguousarray(history[1], np.float32),
                      np.ascontiguousarray(history[2], np.uint32))
    lib().orc_temporal_accumulate(w, h, _p(a, C.c_float), _p(v, C.c_uint32), _p(m, C.c_uint16), 0 if history is None else 1,
                                  _p(hr, C.c_float), _p(hc, C.c_float), _p(hv, C.c_uint32), float(max_history),
                                  _p(out, C.c_float), _p(cnt, C.c_float))
    return out, cnt, v.copy()


class Scene:
    def __init__(self, positions, indices, albedo):
        self.positions = np.ascontiguousarray(positions, np.float32)
        self.indices = np.ascontiguousarray(indices, np.uint32)
        self.albedo = np.ascontiguousarray(albedo, np.float32)
        self.ntris = self.indices.shape[0]
        self.h = lib().orc_scene_create(_p(self.positions, C.c_float), self.positions.shape[0],
                                        _p(self.indices, C.c_uint32), self.ntris, _p(self.albedo, C.c_float))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def closest_hit(self, o, d, use_bvh=True):
        t, u, v = C.c_float(), C.c_float(), C.c_float()
        i = lib().orc_scene_closest_hit(self.h, f3(o), f3(d), int(use_bvh), C.byref(t), C.byref(u), C.byref(v))
        return i, t.value, u.value, v.value

    def primary(self, w, h, pc, use_bvh=True, rows=None):
        y0, y1 = rows if rows else (0, h)
        vis = np.full((h, w), NONE_ID, np.uint32)
        depth = np.zeros((h, w), np.uint16)
        normal = np.zeros((h, w, 4), np.uint16)
        motion = np.zeros((h, w, 2), np.uint16)
        t = np.zeros((h, w), np.float32)
        lib().orc_primary_rays_tris(self.h, w, h, C.byref(pc), int(use_bvh), y0, y1, _p(vis, C.c_uint32),
                                    _p(depth, C.c_uint16), _p(normal, C.c_uint16), _p(motion, C.c_uint16),
                                    _p(t, C.c_float))
        return vis, depth, normal, motion, t

    def render(self, w, h, pc, sc, bn, atmo, trans, view, spp, bounces, use_bvh=True, rows=None, accum=None, ext=0, aerial=None):
        y0, y1 = rows if rows else (0, h)
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        vis = np.full((h, w), NONE_ID, np.uint32)
        rays = (C.c_uint64 * 2)()
        lib().orc_render_tris_ext(self.h, w, h, C.byref(pc), C.byref(sc), _p(bn, C.c_uint8), bn.shape[1], bn.shape[0],
                                  C.byref(atmo), _p(trans, C.c_uint16), _p(view, C.c_uint32), spp, bounces, int(use_bvh),
                                  y0, y1, _p(accum, C.c_float), _p(vis, C.c_uint32), rays, ext,
                                  _p(aerial, C.c_uint16) if aerial is not None else None)
        return accum, vis, (rays[0], rays[1])


EXT_NEE_SUN, EXT_SKY_AT_HIT, EXT_AERIAL = 1, 2, 4
SUN_DIRECTION = (-0.435286462, 0.818654716, 0.374606609)
SUN_ILLUMINANCE = (8.0, 8.0, 8.0)


def aerial_perspective(atmo, trans, multi, pc, camera_pos, sun_dir=SUN_DIRECTION, sun_ill=SUN_ILLUMINANCE):
    """orc_gen_aerial_perspective: the 32^3 RGBA16F camera volume (z, y, x, 4) for the camera of the constants pc."""
    out = np.zeros((32, 32, 32, 4), np.uint16)
    lib().orc_gen_aerial_perspective(C.byref(atmo), _p(trans, C.c_uint16), _p(multi, C.c_uint16), C.byref(pc.invView),
                                     C.byref(pc.invProjection), f3(camera_pos), f3(sun_dir), f3(sun_ill), _p(out, C.c_uint16))
    return out


def ray_gen(pc, x, y, w, h):
    """primaryRay.comp:40-56 for pixel (x, y): origin and direction (float32 arrays)."""
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().orc_ray_gen(C.byref(pc.invView), C.byref(pc.invProjection), int(x), int(y), w, h, _p(o, C.c_float), _p(d, C.c_float))
    return o, d


def aerial_lookup(vol, u, v, t):
    out = np.zeros(4, np.float32)
    lib().orc_aerial_perspective_lookup(_p(vol, C.c_uint16), float(u), float(v), float(t), _p(out, C.c_float))
    return out


def nee_sun_sample(u0, u1):
    l = np.zeros(3, np.float32)
    w = C.c_float()
    lib().orc_nee_sun_sample(float(u0), float(u1), _p(l, C.c_float), C.byref(w))
    return l, w.value


def sun_centre_radiance(atmo, trans, view, pos):
    out = np.zeros(3, np.float32)
    lib().orc_sun_centre_radiance(C.byref(atmo), _p(trans, C.c_uint16), _p(view, C.c_uint32), f3(pos), _p(out, C.c_float))
    return out


def resolve(accum):
    h, w = accum.shape[:2]
    out = np.zeros_like(accum)
    lib().orc_resolve(h * w, _p(np.ascontiguousarray(accum), C.c_float), _p(out, C.c_float))
    return out


def f16_to_f32(a):
    return a.view(np.float16).astype(np.float32)


def psnr(a, b, peak):
    a = np.clip(np.nan_to_num(a.astype(np.float64), nan=0.0, posinf=peak), 0, peak)
    b = np.clip(np.nan_to_num(b.astype(np.float64), nan=0.0, posinf=peak), 0, peak)
    mse = np.mean((a - b) ** 2)
    return 999.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
