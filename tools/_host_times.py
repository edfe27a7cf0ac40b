import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, "/root/repo")
import torch
from PIL import Image
from minotert_b200 import host, scenes
w, h = 1920, 1080
pos, idx, alb, view = scenes.scene_1m()
bn = np.ascontiguousarray(np.array(Image.open("/root/repo/assets/blue_noise.png").convert("RGBA"), np.uint8))
K = int(os.environ.get('K','2'))
MODE = os.environ.get('MODE','full')
r = host.Renderer(w, h, bn, frames_in_flight=K)
r.set_mesh(pos, idx, alb); r.configure(samples=1, bounces=1); r.set_option("async_update", 1)
cam = host.make_camera(w, h, view["position"], view["yaw_deg"], view["pitch_deg"])
frames = [torch.from_numpy(np.ascontiguousarray(scenes.animate(pos, f / 60.0))).pin_memory().numpy() for f in range(8)]
fbs = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(K)]
T = np.zeros(5)
N = 200
for f in range(N + 10):
    if f == 10: T[:] = 0; t_all = time.perf_counter()
    t0 = time.perf_counter()
    host.freecam_update(cam, 1 / 60, up=True, moving=True, cursor=(2.0, 0.0))
    t1 = time.perf_counter()
    if MODE == 'full': r.update_mesh(frames[f % 8], refit=True)
    elif MODE == 'same': r.update_mesh(frames[0], refit=True)
    t2 = time.perf_counter()
    r.draw(cam)
    t3 = time.perf_counter()
    if os.environ.get('NOREAD') != '1': r.read_framebuffer_async(C.c_void_p(fbs[f % K].data_ptr()), fbs[f % K].numel())
    t4 = time.perf_counter()
    r.wait_framebuffer(K - 1)
    t5 = time.perf_counter()
    T += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
r.wait_framebuffer(0)
tot = time.perf_counter() - t_all
print("ms/frame %.3f; host ms: freecam %.3f update %.3f draw %.3f readback_async %.3f wait %.3f" % ((tot / N * 1e3,) + tuple(T / N * 1e3)))
st = r.stats()
print("gpu ms: build %.3f primary %.3f secondary %.3f tonemap %.3f sky %.3f" % (st.ms_build, st.ms_primary, st.ms_secondary, st.ms_tonemap, st.ms_sky))
r.close()
