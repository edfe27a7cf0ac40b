"""Generates tests/golden/oracle_v2.npz from the CPU oracle: the two stages after the path tracer.

Same role as make_golden.py (the reference ships no vectors and cannot run here): freezes the oracle's bilateral
denoiser (src/gpu/denoise/bilateral.comp) and temporal reprojection (contract in oracle/minote_oracle.h) outputs on
the reference's own sphere scene at 96x54, so later edits of the oracle cannot silently change them.
Run:  python tests/golden/make_golden_v2.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402

W, H, FRAMES = 96, 54, 4


def camera_at(k):
    """the reference's initial camera (app.ixx:20-32), turned and moved a little per frame"""
    cam = O.default_camera(W, H)
    O.lib().orc_camera_rotate(C.byref(cam), 5.0 * k, -1.5 * k)
    O.lib().orc_camera_shift(C.byref(cam), (C.c_float * 3)(0.00004 * k, 0.0, 0.00002 * k))
    return cam


def compute():
    atmo, bn, sp = O.earth(), O.load_blue_noise(), O.spheres_array()
    out, hist, prev = {}, None, camera_at(0)
    for f in range(1, FRAMES + 1):
        cam = camera_at(f - 1)
        pc, sc = O.constants(cam, prev=prev, frame=f)
        trans, multi, view = O.sky_luts(atmo, cam.position[:])
        vis, depth, normal, motion = O.primary_spheres(W, H, pc, sp)
        c16, c32, rays = O.secondary_spheres(W, H, sc, sp, vis, depth, normal, bn, atmo, trans, view, 2, 3)
        if f == 1:
            out["denoised_frame1"] = O.denoise_bilateral(c16, depth, normal, O.BILATERAL_DEFAULT, cam.nearPlane, 1)
            out["denoised_frame1_sigma2"] = O.denoise_bilateral(c16, depth, normal, (2.0, 1.5, 0.3), cam.nearPlane, 7)
        acc = np.concatenate([c32[..., :3], np.ones((H, W, 1), np.float32)], -1)  # the frame average with a sample count of 1
        hist = O.temporal_accumulate(acc, vis, motion, hist, 8.0)
        prev = cam
    out["motion_frame%d" % FRAMES] = motion
    out["temporal_rgba"], out["temporal_count"] = hist[0], hist[1]
    return out


def main():
    out = compute()
    np.savez_compressed(os.path.join(HERE, "oracle_v2.npz"), **out)
    print("wrote oracle_v2.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
