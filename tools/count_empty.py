#!/usr/bin/env python
"""Share of node steps that hit no child (diagnostic build -DTRACE_COUNT_EMPTY, which reports them through
mrt_stats.stack_overflows): tools/build_variant.sh empty "-DTRACE_COUNT_EMPTY"; MINOTERT_LIB_DIR=variants/empty python tools/count_empty.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from minotert_b200 import capi, scenes
for name, w, h, spp, bounces in (("hall_260k", 1920, 1080, 1, 2), ("scene_1m", 1920, 1080, 1, 1)):
    pos, idx, alb, view = getattr(scenes, name)()
    ctx = capi.Context(0)
    ctx.upload_blue_noise(O.load_blue_noise()); ctx.atmosphere(O.earth())
    cam = O.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
    ctx.upload_mesh(pos, idx, alb); ctx.build()
    ctx.set_option("count_visits", 1)
    pc, sc = O.constants(cam, frame=1)
    as_capi = lambda x, T: T.from_buffer_copy(bytes(x))
    ctx.sky_view(cam.position[:], O.SUN_DIRECTION, O.SUN_ILLUMINANCE)
    ctx.stats_reset()
    ctx.primary_rays(w, h, as_capi(pc, capi.PrimaryConstants))
    st0 = ctx.stats()
    ctx.secondary_rays(as_capi(sc, capi.SecondaryConstants), spp, bounces)
    st = ctx.stats()
    print(name, "primary: node visits", st0.node_visits, "empty", st0.stack_overflows, "share %.3f" % (st0.stack_overflows / max(1, st0.node_visits)),
          "| all: node visits", st.node_visits, "empty", st.stack_overflows, "share %.3f" % (st.stack_overflows / max(1, st.node_visits)),
          "| bounce only share %.3f" % ((st.stack_overflows - st0.stack_overflows) / max(1, st.node_visits - st0.node_visits)))
    ctx.close()
