mkdir -p gpurun_out
MINOTERT_LIB_DIR=$PWD/variants/prof timeout 100 python - > gpurun_out/prof.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.')
from minotert_b200 import capi, scenes
ctx=capi.Context(0)
pos,idx,alb,_=scenes.scene_1m()
ctx.upload_mesh(pos,idx,alb)
ctx.build(); ctx.sync()
print("==== second build", flush=True)
ctx.build(); ctx.sync()
PY
sed -n '/==== second/,$p' gpurun_out/prof.log | head -150
