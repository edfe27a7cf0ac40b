"""ctypes binding of libminote_host.so -- the C++20-module host side (minotert_b200/host/*.cppm) that
mirrors the reference's src/gfx interface: Camera, Pathtracer, Atmosphere/Sky, Tonemapper,
Renderer::draw, Freecam.  Python only marshals arguments; the matrix math and the call order live in
the C++ modules, the rendering in libminotert.so."""
import ctypes as C
import os

import numpy as np

from . import capi

LIB_PATH = os.path.join(capi.LIB_DIR, "libminote_host.so")


class Camera(C.Structure):
    """src/gfx/camera.ixx:8-22 (POD shared with the C++ Camera class)."""
    _fields_ = [("viewport", C.c_uint32 * 2), ("verticalFov", C.c_float), ("nearPlane", C.c_float),
                ("position", C.c_float * 3), ("yaw", C.c_float), ("pitch", C.c_float),
                ("lookSpeed", C.c_float), ("moveSpeed", C.c_float)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise capi.MinoteError(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C minotert_b200/host)")
    capi.load()  # libminotert.so first (rpath $ORIGIN also finds it)
    L = C.CDLL(LIB_PATH)
    vp, u32, cam = C.c_void_p, C.c_uint32, C.POINTER(Camera)
    L.minote_camera_constants.argtypes = [cam, cam, u32, vp, vp]
    L.minote_camera_constants.restype = None
    L.minote_camera_direction.argtypes = [cam, C.POINTER(C.c_float)]
    L.minote_camera_rotate.argtypes = [cam, C.c_float, C.c_float]
    L.minote_camera_shift.argtypes = [cam, C.POINTER(C.c_float)]
    L.minote_camera_roam.argtypes = [cam, C.POINTER(C.c_float)]
    L.minote_camera_default.argtypes = [cam, u32, u32]
    L.minote_deg.argtypes = [C.c_double]
    L.minote_deg.restype = C.c_float
    L.minote_atmosphere_earth.argtypes = [vp]
    L.minote_freecam_update.argtypes = [cam, u32, C.c_float, C.c_float, C.c_float]
    for n in ("minote_camera_direction", "minote_camera_rotate", "minote_camera_shift", "minote_camera_roam",
              "minote_camera_default", "minote_atmosphere_earth", "minote_freecam_update", "minote_app_destroy"):
        getattr(L, n).restype = None
    L.minote_app_create.argtypes = [C.c_int, u32, u32, vp, u32, u32]
    L.minote_app_create.restype = vp
    L.minote_app_create_in_flight.argtypes = [C.c_int, C.c_int, u32, u32, vp, u32, u32]
    L.minote_app_create_in_flight.restype = vp
    L.minote_app_frame_context.argtypes = [vp, C.c_int]
    L.minote_app_frame_context.restype = vp
    L.minote_app_frames_in_flight.argtypes = [vp]
    L.minote_app_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.minote_app_stats_reset.argtypes = [vp]
    L.minote_app_destroy.argtypes = [vp]
    L.minote_app_error.argtypes = [vp]
    L.minote_app_error.restype = C.c_char_p
    L.minote_app_context.argtypes = [vp]
    L.minote_app_context.restype = vp
    L.minote_app_set_spheres.argtypes = [vp, vp, u32]
    L.minote_app_set_mesh.argtypes = [vp, vp, u32, vp, u32, vp]
    L.minote_app_update_mesh.argtypes = [vp, vp, u32, C.c_int]
    L.minote_app_configure.argtypes = [vp, u32, u32, C.c_int, C.c_int, C.c_float]
    L.minote_app_set_denoise.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float]
    L.minote_app_set_temporal.argtypes = [vp, C.c_int, C.c_float]
    L.minote_app_set_sky_extensions.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.minote_app_resize.argtypes = [vp, u32, u32]
    L.minote_app_draw.argtypes = [vp, cam]
    L.minote_app_read_framebuffer.argtypes = [vp, vp, C.c_size_t]
    L.minote_app_stats.argtypes = [vp, C.POINTER(capi.Stats)]
    L.minote_app_read_framebuffer_async.argtypes = [vp, vp, C.c_size_t]
    L.minote_app_wait_framebuffer.argtypes = [vp, C.c_int]
    L.minote_app_frame_time.argtypes = [vp]
    L.minote_app_frame_time.restype = C.c_float
    L.minote_app_frame_count.argtypes = [vp]
    L.minote_app_frame_count.restype = u32
    _lib = L
    return L


def deg(d):
    return float(load().minote_deg(float(d)))


def default_camera(w=960, h=540):
    c = Camera()
    load().minote_camera_default(C.byref(c), w, h)
    return c


def make_camera(w, h, position, yaw_deg, pitch_deg, vfov_deg=60.0, near=0.001):
    c = default_camera(w, h)
    c.position[:] = [float(x) for x in position]
    c.yaw, c.pitch, c.verticalFov, c.nearPlane = deg(yaw_deg), deg(pitch_deg), deg(vfov_deg), near
    return c


def camera_constants(cam, prev=None, frame=1):
    """Pathtracer's UBO fill (pathtracer.ixx:94-104, 178-188) by the C++ host module."""
    pc, sc = capi.PrimaryConstants(), capi.SecondaryConstants()
    load().minote_camera_constants(C.byref(cam), C.byref(prev) if prev is not None else None, frame,
                                   C.cast(C.byref(pc), C.c_void_p), C.cast(C.byref(sc), C.c_void_p))
    return pc, sc


def atmosphere_earth():
    p = capi.AtmosphereParams()
    load().minote_atmosphere_earth(C.cast(C.byref(p), C.c_void_p))
    return p


def freecam_update(cam, frame_time, up=False, down=False, left=False, right=False, floating=False, moving=False,
                   cursor=(0.0, 0.0)):
    keys = (1 if up else 0) | (2 if down else 0) | (4 if left else 0) | (8 if right else 0) | (16 if floating else 0) | \
           (32 if moving else 0)
    load().minote_freecam_update(C.byref(cam), keys, cursor[0], cursor[1], frame_time)


class Renderer:
    """Cuda::Provider + Renderer::Provider and the per-frame Renderer::serv->draw(camera) call."""

    def __init__(self, width, height, blue_noise_rgba8, device=0, frames_in_flight=1):
        """frames_in_flight: 1..3 frame contexts draw() rotates through (the reference keeps 3, renderer.ixx:36); the
        scene lives in frame context 0 and is borrowed by the others (mrt_scene_share)."""
        self.L = load()
        bn = np.ascontiguousarray(blue_noise_rgba8, np.uint8)
        self.h = self.L.minote_app_create_in_flight(device, frames_in_flight, width, height, bn.ctypes.data_as(C.c_void_p),
                                                    bn.shape[1], bn.shape[0])
        if not self.h:
            raise capi.MinoteError("Renderer: " + self.L.minote_app_error(None).decode())
        self.size = (width, height)

    def close(self):
        if getattr(self, "h", None):
            self.L.minote_app_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, s):
        if s != 0:
            raise capi.MinoteError(self.L.minote_app_error(self.h).decode())

    def set_spheres(self, spheres):
        arr = (capi.Sphere * max(1, len(spheres)))()
        for i, (c, r, al) in enumerate(spheres):
            arr[i].center[:] = c
            arr[i].radius = r
            arr[i].albedo[:] = al
        self._ck(self.L.minote_app_set_spheres(self.h, C.cast(arr, C.c_void_p), len(spheres)))

    def set_mesh(self, positions, indices, albedo):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        a = np.ascontiguousarray(albedo, np.float32).reshape(-1, 3)
        self._ck(self.L.minote_app_set_mesh(self.h, p.ctypes.data_as(C.c_void_p), p.shape[0], i.ctypes.data_as(C.c_void_p),
                                            i.shape[0], a.ctypes.data_as(C.c_void_p)))

    def update_mesh(self, positions, refit=True):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self._ck(self.L.minote_app_update_mesh(self.h, p.ctypes.data_as(C.c_void_p), p.shape[0], int(refit)))

    def configure(self, samples=8, bounces=8, accumulate=False, tonemap="amd", exposure=1.0, denoise="none",
                  bilateral=capi.BILATERAL_DEFAULT):
        """Renderer settings (the reference's ImGui statics).  NOTE: the C++ Renderer defaults to the reference's
        bilateral denoiser (renderer.ixx:140); this method sets the mode explicitly and its own default is "none",
        so that the ray-throughput measurements and the path-trace parity tests see the path tracer's image."""
        self._ck(self.L.minote_app_configure(self.h, samples, bounces, int(accumulate), capi.TONEMAP[tonemap], exposure))
        self._ck(self.L.minote_app_set_denoise(self.h, capi.DENOISE[denoise], *[float(x) for x in bilateral]))

    def set_sky_extensions(self, sun_sampling=False, sky_at_hit=False, aerial_perspective=False):
        """Pathtracer::sunSampling / skyAtHit / aerialPerspective (triangle scenes)."""
        self._ck(self.L.minote_app_set_sky_extensions(self.h, int(sun_sampling), int(sky_at_hit), int(aerial_perspective)))

    def set_temporal(self, enabled, max_history=32.0):
        """Temporal accumulation along GBuffer::motion (Reprojector::accumulate) instead of the denoiser."""
        self._ck(self.L.minote_app_set_temporal(self.h, int(enabled), float(max_history)))

    def draw(self, camera):
        self._ck(self.L.minote_app_draw(self.h, C.byref(camera)))

    def read_framebuffer(self, out=None):
        w, h = self.size
        if out is None:
            out = np.empty((h, w, 4), np.uint8)
        self._ck(self.L.minote_app_read_framebuffer(self.h, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def read_framebuffer_into(self, host_ptr, nbytes):
        self._ck(self.L.minote_app_read_framebuffer(self.h, host_ptr, nbytes))

    def read_framebuffer_async(self, host_ptr, nbytes):
        self._ck(self.L.minote_app_read_framebuffer_async(self.h, host_ptr, nbytes))

    def wait_framebuffer(self, frames_in_flight=0):
        self._ck(self.L.minote_app_wait_framebuffer(self.h, frames_in_flight))

    def stats(self):
        s = capi.Stats()
        self._ck(self.L.minote_app_stats(self.h, C.byref(s)))
        return s

    def frame_time(self):
        """Renderer_impl::frameTime(): moving average over 0.25 s of the time between draw() calls, in seconds."""
        return float(self.L.minote_app_frame_time(self.h))

    def frame_count(self):
        return self.L.minote_app_frame_count(self.h)

    def frames_in_flight(self):
        return self.L.minote_app_frames_in_flight(self.h)

    def set_option(self, name, value):
        """mrt_set_option on every frame context."""
        self._ck(self.L.minote_app_set_option(self.h, name.encode(), int(value)))

    def stats_reset(self):
        self._ck(self.L.minote_app_stats_reset(self.h))

    def context(self, frame=0):
        """Borrowed capi-level view of one of the renderer's mrt_contexts (buffers, stream, options): frame context
        `frame` (0 owns the scene; the only one with one frame in flight), or the one the last draw() recorded into
        (frame=-1)."""
        ctx = capi.Context.__new__(capi.Context)
        ctx.L = capi.load()
        ctx.h = C.c_void_p(self.L.minote_app_context(self.h) if frame == 0 else self.L.minote_app_frame_context(self.h, frame))
        if not ctx.h:
            raise capi.MinoteError(f"Renderer: no frame context {frame}")
        ctx.device = 0
        ctx.size = self.size
        ctx.close = lambda: None  # not owned
        return ctx
