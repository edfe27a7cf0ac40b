"""Generates tests/golden/oracle_v1.npz from the CPU oracle.

The reference (Tearnote/MinoteRT) ships no golden vectors and cannot run here (GLSL/Vulkan/MSVC), so
these fixtures pin OUR restatement: they freeze the oracle's outputs at the time its KATs (PCG,
ray/sphere, matrices - SURVEY.md §8c) were checked, so later edits cannot silently change them, and
they give the GPU tests a committed target that does not depend on rebuilding the oracle.
Run:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from minotert_b200 import scenes  # noqa: E402


def main():
    L = O.lib()
    out = {}
    s = C.c_uint32(3)
    out["pcg_seed3"] = np.array([L.orc_pcg(C.byref(s)) for _ in range(16)], np.uint32)
    cam = O.default_camera(96, 54)
    pc, sc = O.constants(cam, frame=1)
    out["primary_constants_96x54"] = np.frombuffer(bytes(pc), np.uint8).copy()
    out["secondary_constants_96x54"] = np.frombuffer(bytes(sc), np.uint8).copy()
    atmo = O.earth()
    out["atmosphere_earth"] = np.frombuffer(bytes(atmo), np.uint8).copy()
    trans, multi, view = O.sky_luts(atmo, cam.position[:])
    out["sky_transmittance"], out["sky_multiscattering"], out["sky_view"] = trans, multi, view
    bn = O.load_blue_noise()
    sp = O.spheres_array()
    vis, depth, normal, motion = O.primary_spheres(96, 54, pc, sp)
    c16, c32, rays = O.secondary_spheres(96, 54, sc, sp, vis, depth, normal, bn, atmo, trans, view, 8, 8)
    out.update(spheres_vis=vis, spheres_depth=depth, spheres_normal=normal, spheres_motion=motion,
               spheres_color16=c16, spheres_rays=np.array([rays], np.uint64))
    for mode, params in [("linear", (0.0,)), ("reinhard", (8.0,)), ("hable", (0.0,)), ("aces", (0.0,)),
                         ("uchimura", O.UCHIMURA_DEFAULT), ("amd", O.AMD_DEFAULT)]:
        out["spheres_ldr_" + mode] = O.tonemap(mode, c16, 1.0, params)
    pos, idx, alb, v = scenes.cornell()
    camc = O.make_camera(64, 64, v["position"], v["yaw_deg"], v["pitch_deg"])
    pcc, scc = O.constants(camc, frame=1)
    tr2, mu2, vw2 = O.sky_luts(atmo, camc.position[:])
    scn = O.Scene(pos, idx, alb)
    acc, cvis, crays = scn.render(64, 64, pcc, scc, bn, atmo, tr2, vw2, 2, 2, use_bvh=False)
    out.update(cornell_positions_sha=np.frombuffer(__import__("hashlib").sha256(pos.tobytes() + idx.tobytes() + alb.tobytes()).digest(), np.uint8).copy(),
               cornell_vis=cvis, cornell_accum=acc, cornell_rays=np.array(crays, np.uint64))
    np.savez_compressed(os.path.join(HERE, "oracle_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_v1.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
