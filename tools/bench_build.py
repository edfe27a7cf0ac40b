#!/usr/bin/env python
"""BVH build time (row n2) per scene, builder and loop placement: warm builds, CUDA events around the whole build
(mrt_stats.ms_build), kernel launches per build.  One JSON line per combination.

    python tools/bench_build.py [--scenes hall_260k scene_1m] [--repeat 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", nargs="+", default=["hall_260k", "scene_1m"])
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()
    from minotert_b200 import capi, scenes
    ctx = capi.Context(0)
    for name in args.scenes:
        pos, idx, alb, _ = getattr(scenes, name)()
        ctx.upload_mesh(pos, idx, alb)
        for builder in ("lbvh", "ploc"):
            for device_loop in (0, 1):
                ctx.set_option("builder", 1 if builder == "ploc" else 0)
                ctx.set_option("build_device_loop", device_loop)
                ctx.build()  # first build of this configuration: allocations, module load
                ms, launches = [], 0
                for _ in range(args.repeat):
                    ctx.stats_reset()
                    ctx.build()
                    st = ctx.stats()
                    ms.append(st.ms_build)
                    launches = st.kernel_launches
                ntri = int(idx.shape[0])
                print(json.dumps({"scene": name, "triangles": ntri, "builder": builder, "device_loop": bool(device_loop),
                                  "ms_build_median": float(np.median(ms)), "ms_build_min": float(np.min(ms)),
                                  "kernel_launches": int(launches), "wide_nodes": int(st.num_wide_nodes),
                                  "Mtris_per_s": ntri / (float(np.median(ms)) * 1e-3) / 1e6,
                                  "algorithmic_GBps": ntri * 410 / (float(np.median(ms)) * 1e-3) / 1e9}), flush=True)
        for mode, label in ((capi.BUILD_REFIT, "refit"),):
            ctx.set_option("builder", 1)
            ctx.set_option("build_device_loop", 1)
            ctx.build()
            ms = []
            for _ in range(args.repeat):
                ctx.update_positions(pos)
                ctx.build(mode)
                ms.append(ctx.stats().ms_build)
            print(json.dumps({"scene": name, "mode": label, "ms_median": float(np.median(ms))}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
