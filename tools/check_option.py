#!/usr/bin/env python
"""Bitwise A/B of an mrt_set_option switch on the GPU box: renders one frame of a scene with and without the option(s)
and compares visibility, the fp32 primary hit, the accumulator and the LDR framebuffer bit for bit.
usage: tools/check_option.py <scene> <w> <h> <spp> <bounces> name=value [name=value ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from minotert_b200 import capi, host, scenes  # noqa: E402
import oracle_lib as O  # noqa: E402  (blue noise asset + atmosphere parameters only)


def render(scene, w, h, spp, bounces, opts):
    pos, idx, alb, v = getattr(scenes, scene)()
    c = capi.Context(0)
    try:
        c.upload_blue_noise(O.load_blue_noise())
        for k, val in opts:
            c.set_option(k, val)
        c.upload_mesh(pos, idx, alb)
        c.build()
        cam = host.make_camera(w, h, v["position"], v["yaw_deg"], v["pitch_deg"])
        pc, sc = host.camera_constants(cam, cam, 1)
        c.atmosphere(host.atmosphere_earth())
        c.sky_view(cam.position[:], (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
        c.primary_rays(w, h, pc)
        c.secondary_rays(sc, spp, bounces, int(os.environ.get("CHECK_FLAGS", "0")))  # CHECK_FLAGS=8: sun sampling
        c.tonemap("amd", 1.0, (16.0, 2.0, 1.0, 0.18, 0.18), capi.BUF_ACCUM)
        out = {n: c.readback(b) for n, b in (("vis", capi.BUF_VISIBILITY), ("accum", capi.BUF_ACCUM), ("ldr", capi.BUF_LDR),
                                              ("depth", capi.BUF_DEPTH), ("normal", capi.BUF_NORMAL))}
        st = c.stats()
        out["overflows"] = st.stack_overflows
        return out
    finally:
        c.close()


def main():
    scene, w, h, spp, bounces = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in sys.argv[6:]]
    if not opts:  # no option given: print the hashes of the buffers (compare two library variants, MINOTERT_LIB_DIR)
        import hashlib
        a = render(scene, w, h, spp, bounces, [])
        print("HASHES", " ".join(f"{k}:{hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()[:12]}" for k, v in a.items() if k != "overflows"),
              "overflows", a["overflows"])
        return
    a = render(scene, w, h, spp, bounces, [])
    b = render(scene, w, h, spp, bounces, opts)
    ok = True
    for k in a:
        if k == "overflows":
            same = a[k] == b[k] == 0
        else:
            same = np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8))
        ok &= bool(same)
        print(f"{k:8s} {'identical' if same else 'DIFFERENT'}")
    print("OPTION PARITY", "OK" if ok else "FAILED", opts)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
