#!/usr/bin/env python
"""Phase times of the cooperative build loops: run with a -DBUILD_PROFILE variant (tools/build_variant.sh bprof "-DBUILD_PROFILE",
MINOTERT_LIB_DIR=variants/bprof); the kernels print the time between phase boundaries (CTA 0, thread 0).  The LAST build's lines count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minotert_b200 import capi, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "scene_1m"
ctx = capi.Context(0)
pos, idx, alb, _ = getattr(scenes, name)()
ctx.upload_mesh(pos, idx, alb)
ctx.set_option("builder", 1)
ctx.build()
print("==== warm build", flush=True)
ctx.build()
print("ms_build", ctx.stats().ms_build)
ctx.close()
