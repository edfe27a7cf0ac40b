"""ctypes binding of oracle/_ref/libminote_ref.so: the reference's own GLSL shaders compiled as C++
(oracle/ref/: glsl2cpp.py pre-pass + glsl_shim.hpp + binders).

Test infrastructure only (tests/, __graft_entry__.smoke(), bench.py's CPU legs).  /root/reference is needed to BUILD
the library (in the development container); at run time only the prebuilt .so is used, which travels to the GPU box.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC_DIR = os.path.join(ROOT, "oracle", "ref")
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libminote_ref.so")
REFERENCE = "/root/reference"


def available():
    """True when the library exists or can be built (the reference tree is present)."""
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "src", "gpu"))


def build():
    if os.path.isdir(os.path.join(REFERENCE, "src", "gpu")):
        subprocess.check_call(["make", "-C", REF_SRC_DIR, "-s"])
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libminote_ref.so is missing and /root/reference is not here to build it")
    return LIB_PATH


_lib = None
_p = O._p


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    f32p, u32p, u16p, u8p = (C.POINTER(t) for t in (C.c_float, C.c_uint32, C.c_uint16, C.c_uint8))
    sig = {
        "ref_pcg": (C.c_uint32, [u32p]),
        "ref_random_float": (C.c_float, [u32p]),
        "ref_random_sphere_point": (None, [C.c_float, C.c_float, f32p]),
        "ref_ray_sphere": (C.c_float, [f32p, f32p, f32p]),
        "ref_scene_spheres": (C.c_uint32, [f32p, C.c_uint32]),
        "ref_primary_rays": (None, [C.c_uint32, C.c_uint32, C.c_void_p, u32p, u16p, u16p, u16p]),
        "ref_secondary_rays": (None, [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, u32p, u16p, u16p, u8p, C.c_uint32,
                                      C.c_uint32, u16p, u32p, u16p]),
        "ref_sky_color": (None, [C.c_void_p, u16p, u32p, f32p, C.c_uint32, f32p, f32p]),
        "ref_gen_transmittance": (None, [C.c_void_p, u16p]),
        "ref_gen_multiscattering": (None, [C.c_void_p, u16p, u16p]),
        "ref_gen_sky_view": (None, [C.c_void_p, u16p, u16p, f32p, f32p, f32p, u32p]),
        "ref_tonemap": (None, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_float, f32p, u8p]),
        "ref_denoise_bilateral": (None, [C.c_uint32, C.c_uint32, u16p, u16p, u16p, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_uint32, u8p]),
        "ref_primary_constants_fill": (None, [C.POINTER(O.Camera), C.POINTER(O.Camera), C.c_uint32, C.c_void_p]),
        "ref_secondary_constants_fill": (None, [C.POINTER(O.Camera), C.c_uint32, C.c_void_p]),
        "ref_camera_direction": (None, [C.POINTER(O.Camera), f32p]),
        "ref_camera_rotate": (None, [C.POINTER(O.Camera), C.c_float, C.c_float]),
        "ref_camera_shift": (None, [C.POINTER(O.Camera), f32p]),
        "ref_camera_roam": (None, [C.POINTER(O.Camera), f32p]),
        "ref_perspective": (None, [C.c_float, C.c_float, C.c_float, f32p]),
        "ref_look": (None, [f32p, f32p, f32p, f32p]),
        "ref_inverse": (None, [f32p, f32p]),
        "ref_mat_mul": (None, [f32p, f32p, f32p]),
        "ref_deg": (C.c_float, [C.c_float]),
        "ref_atmosphere_earth": (None, [C.c_void_p]),
        "ref_set_num_threads": (None, [C.c_int]),
        "ref_num_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def scene_spheres():
    """The sphere scene compiled into the reference's shaders (src/gpu/scene.glsl) as [(center, radius, albedo)]."""
    buf = np.zeros((16, 7), np.float32)
    n = lib().ref_scene_spheres(_p(buf, C.c_float), 16)
    return [(tuple(buf[i, :3]), buf[i, 3], tuple(buf[i, 4:7])) for i in range(n)]


def constants(cam, prev=None, frame=1):
    """The constant blocks as the reference's own host code fills them (camera.ixx + math.ixx, pathtracer.ixx:94-104,179-187)."""
    pc, sc = O.PrimaryConstants(), O.SecondaryConstants()
    lib().ref_primary_constants_fill(C.byref(cam), C.byref(prev if prev is not None else cam), frame, C.byref(pc))
    lib().ref_secondary_constants_fill(C.byref(cam), frame, C.byref(sc))
    return pc, sc


def earth():
    p = O.AtmosphereParams()
    lib().ref_atmosphere_earth(C.byref(p))
    return p


def sky_luts(atmo, probe_pos, sun_dir=O.SUN_DIRECTION, sun_ill=O.SUN_ILLUMINANCE):
    L = lib()
    trans = np.zeros((O.TRANS_H, O.TRANS_W, 4), np.uint16)
    multi = np.zeros((O.MULTI_H, O.MULTI_W, 4), np.uint16)
    view = np.zeros((O.VIEW_H, O.VIEW_W), np.uint32)
    L.ref_gen_transmittance(C.byref(atmo), _p(trans, C.c_uint16))
    L.ref_gen_multiscattering(C.byref(atmo), _p(trans, C.c_uint16), _p(multi, C.c_uint16))
    L.ref_gen_sky_view(C.byref(atmo), _p(trans, C.c_uint16), _p(multi, C.c_uint16), O.f3(probe_pos), O.f3(sun_dir),
                       O.f3(sun_ill), _p(view, C.c_uint32))
    return trans, multi, view


def primary(w, h, pc):
    vis = np.zeros((h, w), np.uint32)
    depth = np.zeros((h, w), np.uint16)
    normal = np.zeros((h, w, 4), np.uint16)
    motion = np.zeros((h, w, 2), np.uint16)
    lib().ref_primary_rays(w, h, C.byref(pc), _p(vis, C.c_uint32), _p(depth, C.c_uint16), _p(normal, C.c_uint16),
                           _p(motion, C.c_uint16))
    return vis, depth, normal, motion


def secondary(w, h, sc, vis, depth, normal, bn, atmo, trans, view):
    c16 = np.zeros((h, w, 4), np.uint16)
    lib().ref_secondary_rays(w, h, C.byref(sc), C.byref(atmo), _p(vis, C.c_uint32), _p(depth, C.c_uint16),
                             _p(normal, C.c_uint16), _p(bn, C.c_uint8), bn.shape[1], bn.shape[0],
                             _p(trans, C.c_uint16), _p(view, C.c_uint32), _p(c16, C.c_uint16))
    return c16


def sky_color(atmo, trans, view, camera_pos, dirs):
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros_like(d)
    lib().ref_sky_color(C.byref(atmo), _p(trans, C.c_uint16), _p(view, C.c_uint32), O.f3(camera_pos), d.shape[0],
                        _p(d, C.c_float), _p(out, C.c_float))
    return out


def tonemap(mode, src, exposure=1.0, params=O.AMD_DEFAULT):
    h, w = src.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    src = np.ascontiguousarray(src)
    par = (C.c_float * 8)(*params)
    lib().ref_tonemap(O.TONEMAP[mode] if isinstance(mode, str) else mode, w, h, src.ctypes.data_as(C.c_void_p),
                      {np.dtype(np.uint16): 1, np.dtype(np.uint8): 2}.get(src.dtype, 0), exposure, par,
                      _p(out, C.c_uint8))
    return out


def denoise_bilateral(color16, depth16, normal16, params=O.BILATERAL_DEFAULT, near=0.001, frame=1):
    h, w = depth16.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    c, d, n = (np.ascontiguousarray(a, np.uint16) for a in (color16, depth16, normal16))
    lib().ref_denoise_bilateral(w, h, _p(c, C.c_uint16), _p(d, C.c_uint16), _p(n, C.c_uint16), params[0], params[1],
                                params[2], near, frame, _p(out, C.c_uint8))
    return out
