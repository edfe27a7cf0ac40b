#!/usr/bin/env python
"""One warm build + N refits of a scene (for launch lists under ncu): python tools/refit_once.py scene_1m 3"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minotert_b200 import capi, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "scene_1m"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
pos, idx, alb, _ = getattr(scenes, name)()
ctx.upload_mesh(pos, idx, alb)
ctx.set_option("builder", 1)
ctx.build()
ctx.build()
print("ms_build", ctx.stats().ms_build)
for _ in range(reps):
    ctx.update_positions(pos)
    ctx.build(capi.BUILD_REFIT)
    print("ms_refit", ctx.stats().ms_build)
ctx.close()
