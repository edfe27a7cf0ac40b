set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_temporal.py tests/test_abi.py -q > gpurun_out/pytest_temporal.log 2>&1; echo "temporal tests rc=$?"
tail -40 gpurun_out/pytest_temporal.log
for o in "--opt trace_ctas_per_sm=2" "--opt trace_ctas_per_sm=3"; do
  for c in 2 3; do
  echo "== $o contexts $c"; timeout 60 python tools/bench_inflight.py --share 1 --frames 120 --contexts $c $o 2>&1 | cut -c1-200
  done
done
for o in "--opt trace_ctas_per_sm=2" "--opt trace_ctas_per_sm=3"; do
echo "== 1m $o"; timeout 60 python tools/bench_inflight.py --share 1 --frames 120 --contexts 3 --workload scene_1m_1080p $o 2>&1 | cut -c1-200
done
timeout 100 python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench as B
from PIL import Image
from minotert_b200 import capi, host
gen, w, h, spp, bounces = B.WORKLOADS["hall_260k_1080p"]
pos, idx, alb, view = B.make_scene(gen)
bn = np.ascontiguousarray(np.array(Image.open("assets/blue_noise.png").convert("RGBA"), np.uint8))
c = capi.Context(0); c.upload_blue_noise(bn); c.atmosphere(host.atmosphere_earth()); c.upload_mesh(pos, idx, alb); c.build()
cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
c.sky_view(list(cam.position), (-0.435286462, 0.818654716, 0.374606609), (8.0, 8.0, 8.0))
ms = []
import ctypes as C
prev = cam
for f in range(1, 9):
    cur = host.Camera.from_buffer_copy(bytes(prev)); host.load().minote_camera_rotate(C.byref(cur), 3.0, 0.5)
    pc, sc = host.camera_constants(cur, prev, f)
    c.primary_rays(w, h, pc); c.secondary_rays(sc, spp, bounces); c.temporal_accumulate(32.0)
    ms.append(c.stats().ms_temporal); prev = cur
cnt = c.readback(capi.BUF_TEMPORAL_COUNT)
print("temporal ms per 1080p frame:", [round(x, 4) for x in ms], "mean history", float(cnt.mean()))
PY
