"""Generates tests/golden/ref_v1.npz from oracle/_ref -- the REFERENCE'S OWN GLSL shaders (Tearnote/MinoteRT
src/gpu/*.comp|*.glsl, read from /root/reference at build time) compiled as C++ through oracle/ref/glsl_shim.hpp.

These are the golden vectors that pin the CPU restatement (oracle/minote_oracle.c): tests/test_ref_pins_oracle.py
asserts the restatement reproduces them bit for bit, with or without /root/reference at hand.  The camera constant
blocks and the atmosphere parameters come from the reference's own host code too (src/stx/math.ixx, src/gfx/camera.ixx,
Atmosphere::Params::earth of src/gfx/modules/sky.ixx, compiled with the module syntax stripped: oracle/ref/ixx2hpp.py).
Run (in the development container, where /root/reference exists):  python tests/golden/make_golden_ref.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402  (ctypes PODs, the camera POD builder and the blue-noise loader only)
import ref_lib as R  # noqa: E402


def sweep_directions():
    """Directions for skyColor(): uniform sphere, the sun-disc rim (0.2525 deg), the horizon band, near-vertical."""
    rng = np.random.default_rng(3)
    d = rng.normal(size=(3000, 3))
    sun = np.array(O.SUN_DIRECTION)
    k = 1500
    ang = np.deg2rad(rng.uniform(0.0, 0.6, k))
    phi = rng.uniform(0, 2 * np.pi, k)
    t1 = np.cross(sun, [0, 0, 1.0])
    t1 /= np.linalg.norm(t1)
    t2 = np.cross(sun, t1)
    rim = np.cos(ang)[:, None] * sun + np.sin(ang)[:, None] * (np.cos(phi)[:, None] * t1 + np.sin(phi)[:, None] * t2)
    hz = rng.normal(size=(k, 3))
    hz[:, 2] = rng.uniform(-0.02, 0.02, k)
    up = np.array([[1e-7, 0, 1], [0, 1e-6, -1], [1e-3, 1e-3, 1], [0.6, 0, 0.8]])
    a = np.concatenate([d, rim, hz, up])
    return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)


def hdr_sweep():
    rng = np.random.default_rng(7)
    hdr = np.zeros((32, 128, 4), np.float32)
    hdr[..., :3] = np.exp(rng.uniform(-12, 6, (32, 128, 3))).astype(np.float32)
    hdr[0, :8, :3] = [[0, 0, 0], [1e-8, 0, 0], [65504, 1, 0.5], [np.inf, 1, 1], [-1, 0.5, 2], [16, 16, 16], [1, 1, 1],
                      [0.18, 0.18, 0.18]]
    hdr[..., 3] = 1
    return hdr.astype(np.float16).view(np.uint16)


TONEMAPS = [("linear", ()), ("reinhard", (16.0,)), ("hable", ()), ("aces", ()), ("uchimura", O.UCHIMURA_DEFAULT),
            ("amd", O.AMD_DEFAULT)]
POSES = {  # name: (w, h, position, yaw, pitch, prev position, prev yaw, prev pitch, frame)
    "default": (96, 54, (0.0, -0.001, 0.1), 90.0, 0.0, (0.0, -0.001, 0.1), 90.0, 0.0, 1),
    "moved": (80, 60, (0.0004, -0.0012, 0.1003), 97.0, -6.0, (0.0003, -0.0012, 0.1003), 95.0, -5.0, 7),
}


def main():
    L = R.lib()
    out = {}
    s = C.c_uint32(3)
    out["pcg_seed3"] = np.array([L.ref_pcg(C.byref(s)) for _ in range(64)], np.uint32)
    s = C.c_uint32(99)
    out["random_float_seed99"] = np.array([L.ref_random_float(C.byref(s)) for _ in range(64)], np.float32)
    rng = np.random.default_rng(11)
    rr = rng.uniform(-1, 1, (512, 2)).astype(np.float32)
    pts = np.zeros((512, 3), np.float32)
    for i in range(512):
        p = (C.c_float * 3)()
        L.ref_random_sphere_point(float(rr[i, 0]), float(rr[i, 1]), p)
        pts[i] = p[:]
    out["sphere_point_in"], out["sphere_point_out"] = rr, pts
    sph = R.scene_spheres()
    out["scene_spheres"] = np.array([list(c) + [r] + list(a) for c, r, a in sph], np.float32)
    ro = (np.array([0.0, -0.001, 0.1]) + rng.normal(scale=2e-4, size=(512, 3))).astype(np.float32)
    rd = rng.normal(size=(512, 3))
    rd[:, 1] = np.abs(rd[:, 1]) + 1.0
    rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
    ts = np.zeros((512, len(sph)), np.float32)
    for i in range(512):
        for k in range(len(sph)):
            ts[i, k] = L.ref_ray_sphere(O.f3(ro[i]), O.f3(rd[i]), O._p(out["scene_spheres"][k], C.c_float))
    out["ray_sphere_o"], out["ray_sphere_d"], out["ray_sphere_t"] = ro, rd, ts

    atmo = R.earth()
    out["atmosphere_earth"] = np.frombuffer(bytes(atmo), np.uint8).copy()
    bn = O.load_blue_noise()
    for name, (w, h, pos, yaw, pitch, ppos, pyaw, ppitch, frame) in POSES.items():
        cam = O.make_camera(w, h, pos, yaw, pitch)
        prev = O.make_camera(w, h, ppos, pyaw, ppitch)
        pc, sc = R.constants(cam, prev, frame=frame)
        out[name + "_camera"] = np.frombuffer(bytes(cam) + bytes(prev), np.uint8).copy()
        out[name + "_primary_constants"] = np.frombuffer(bytes(pc), np.uint8).copy()
        out[name + "_secondary_constants"] = np.frombuffer(bytes(sc), np.uint8).copy()
        trans, multi, view = R.sky_luts(atmo, cam.position[:])
        if name == "default":   # transmittance and multi-scattering do not depend on the camera
            out["trans"], out["multi"] = trans, multi
        assert np.array_equal(trans, out["trans"]) and np.array_equal(multi, out["multi"])
        out[name + "_view"] = view
        vis, depth, normal, motion = R.primary(w, h, pc)
        c16 = R.secondary(w, h, sc, vis, depth, normal, bn, atmo, trans, view)
        out.update({name + "_vis": vis, name + "_depth": depth, name + "_normal": normal, name + "_motion": motion,
                    name + "_color16": c16})
        out[name + "_denoised"] = R.denoise_bilateral(c16, depth, normal, frame=frame)
        for mode, par in TONEMAPS:
            out[name + "_ldr_" + mode] = R.tonemap(mode, c16, 1.0, par + (0.0,) * (8 - len(par)))
        if name == "default":
            dirs = sweep_directions()
            out["sky_dirs"], out["sky_colors"] = dirs, R.sky_color(atmo, trans, view, cam.position[:], dirs)
    sweep = hdr_sweep()
    out["hdr_sweep16"] = sweep
    for mode, par in TONEMAPS:
        out["sweep_ldr_" + mode] = R.tonemap(mode, sweep, 0.37, par + (0.0,) * (8 - len(par)))
    path = os.path.join(HERE, "ref_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
