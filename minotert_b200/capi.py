"""ctypes binding of libminotert.so -- the C ABI declared in include/minotert.h.

This is a thin binding, not an implementation: every render call lands in the CUDA library.
There is no CPU fallback; if the library is missing or no CUDA device exists the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MINOTERT_LIB_DIR: directory holding an alternative build of libminotert.so / libminote_host.so
# (kernel-tuning variants built with different -D flags); default is the in-tree build.
LIB_DIR = os.environ.get("MINOTERT_LIB_DIR", _HERE)
LIB_PATH = os.path.join(LIB_DIR, "libminotert.so")

MISS_ID = 0xFFFFFFFF
(BUF_VISIBILITY, BUF_DEPTH, BUF_NORMAL, BUF_MOTION, BUF_COLOR, BUF_ACCUM, BUF_LDR, BUF_TRANSMITTANCE,
 BUF_MULTISCATTERING, BUF_SKY_VIEW, BUF_HIT_T, BUF_DENOISED, BUF_BVH_NODES, BUF_BVH_TRIS, BUF_TEMPORAL,
 BUF_TEMPORAL_COUNT) = range(16)
BUF_AERIAL = 16
BUILD_FULL, BUILD_REFIT = 0, 1
SECONDARY_ACCUMULATE, SECONDARY_SORT_RAYS, SECONDARY_FRAME_SUM, SECONDARY_NEE_SUN, SECONDARY_SKY_AT_HIT = 1, 2, 4, 8, 16
SECONDARY_AERIAL = 32
TEMPORAL_RESET = 1
TONEMAP = {"linear": 0, "reinhard": 1, "hable": 2, "aces": 3, "uchimura": 4, "amd": 5}
DENOISE = {"none": 0, "bilateral": 1}   # Renderer_impl::denoise modes (src/gfx/renderer.ixx:129-132)
BILATERAL_DEFAULT = (5.0, 2.0, 0.12)   # sigma, kSigma, threshold (src/gfx/modules/denoiser.ixx:27-33)


class Mat4(C.Structure):
    _fields_ = [("m", (C.c_float * 4) * 4)]


class PrimaryConstants(C.Structure):
    _fields_ = [("view", Mat4), ("projection", Mat4), ("invView", Mat4), ("invProjection", Mat4),
                ("prevView", Mat4), ("frameCounter", C.c_uint32)]


class SecondaryConstants(C.Structure):
    _fields_ = [("view", Mat4), ("projection", Mat4), ("invView", Mat4), ("invProjection", Mat4),
                ("cameraPos", C.c_float * 3), ("frameCounter", C.c_uint32)]


class Sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("albedo", C.c_float * 3)]


class AtmosphereParams(C.Structure):
    _fields_ = [("raw", C.c_float * 36)]


class Stats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("secondary_rays", C.c_uint64), ("ms_primary", C.c_float),
                ("ms_secondary", C.c_float), ("ms_trace", C.c_float), ("ms_tonemap", C.c_float),
                ("ms_build", C.c_float), ("ms_sky", C.c_float), ("kernel_launches", C.c_uint32),
                ("stack_overflows", C.c_uint32), ("num_triangles", C.c_uint32), ("num_wide_nodes", C.c_uint32),
                ("bvh_bytes", C.c_uint64), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64),
                ("trace_launches", C.c_uint32), ("_reserved", C.c_uint32), ("total_rays", C.c_uint64), ("sah_node_cost", C.c_float), ("sah_tri_cost", C.c_float),
                ("ms_denoise", C.c_float), ("ms_temporal", C.c_float),
                ("secondary_node_visits", C.c_uint64), ("secondary_tri_tests", C.c_uint64)]


EXPORTS = [
    "mrt_abi_version", "mrt_create", "mrt_destroy", "mrt_last_error", "mrt_set_option", "mrt_upload_blue_noise",
    "mrt_scene_set_spheres", "mrt_scene_upload_mesh", "mrt_scene_update_positions", "mrt_scene_build",
    "mrt_atmosphere", "mrt_sky_view", "mrt_sky_aerial_perspective", "mrt_set_partition", "mrt_partition_rows", "mrt_primary_rays",
    "mrt_secondary_rays", "mrt_tonemap", "mrt_buffer", "mrt_readback", "mrt_sync", "mrt_stats_get",
    "mrt_stats_reset", "mrt_stream", "mrt_trace_rays", "mrt_partition_rows_for", "mrt_readback_async",
    "mrt_readback_wait", "mrt_denoise_bilateral", "mrt_scene_share", "mrt_temporal_accumulate",
    "mrt_eval_sky_color", "mrt_eval_bounce_stream", "mrt_accum_restore", "mrt_accum_commit",
    "mrt_group_set_frames_in_flight", "mrt_group_frame_context", "mrt_group_readback_async", "mrt_group_readback_wait",
    "mrt_group_create", "mrt_group_unique_id", "mrt_group_create_rank", "mrt_group_destroy", "mrt_group_last_error",
    "mrt_group_size", "mrt_group_context", "mrt_group_set_tiles", "mrt_group_render", "mrt_group_tonemap",
    "mrt_group_gather", "mrt_group_reduce", "mrt_group_result", "mrt_group_readback", "mrt_group_sync",
]
GROUP_NCCL, GROUP_P2P = 0, 1


class MinoteError(RuntimeError):
    pass


_lib = None


def load():
    """Load libminotert.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MinoteError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C minotert_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, f32p = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
    L.mrt_abi_version.restype = C.c_int
    L.mrt_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.mrt_destroy.argtypes = [vp]
    L.mrt_destroy.restype = None
    L.mrt_last_error.argtypes = [vp]
    L.mrt_last_error.restype = C.c_char_p
    L.mrt_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.mrt_upload_blue_noise.argtypes = [vp, vp, u32, u32]
    L.mrt_scene_set_spheres.argtypes = [vp, vp, u32]
    L.mrt_scene_upload_mesh.argtypes = [vp, vp, u32, vp, u32, vp]
    L.mrt_scene_update_positions.argtypes = [vp, vp, u32]
    L.mrt_scene_build.argtypes = [vp, C.c_int]
    L.mrt_scene_share.argtypes = [vp, vp]
    L.mrt_temporal_accumulate.argtypes = [vp, C.c_float, u32]
    L.mrt_atmosphere.argtypes = [vp, vp]
    L.mrt_sky_view.argtypes = [vp, f32p, f32p, f32p]
    L.mrt_sky_aerial_perspective.argtypes = [vp, vp, f32p, f32p, f32p]
    L.mrt_set_partition.argtypes = [vp, u32, u32, u32]
    L.mrt_partition_rows.argtypes = [vp, u32, vp, C.POINTER(u32)]
    L.mrt_primary_rays.argtypes = [vp, u32, u32, vp]
    L.mrt_secondary_rays.argtypes = [vp, vp, u32, u32, u32]
    L.mrt_denoise_bilateral.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, u32]
    L.mrt_tonemap.argtypes = [vp, C.c_int, C.c_float, f32p, u32, C.c_int]
    L.mrt_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mrt_readback.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.mrt_sync.argtypes = [vp]
    L.mrt_readback_async.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.mrt_readback_wait.argtypes = [vp, C.c_int]
    L.mrt_stats_get.argtypes = [vp, C.POINTER(Stats)]
    L.mrt_stats_reset.argtypes = [vp]
    L.mrt_stream.argtypes = [vp, C.POINTER(vp)]
    L.mrt_trace_rays.argtypes = [vp, vp, vp, u32, vp, vp, C.c_int]
    L.mrt_partition_rows_for.argtypes = [u32, u32, u32, u32, vp, C.POINTER(u32)]
    L.mrt_accum_restore.argtypes = [vp, vp, C.c_size_t]
    L.mrt_accum_commit.argtypes = [vp, vp, u32]
    L.mrt_group_set_frames_in_flight.argtypes = [vp, u32]
    L.mrt_group_frame_context.argtypes = [vp, u32, u32, C.POINTER(vp)]
    L.mrt_group_readback_async.argtypes = [vp, vp, C.c_size_t]
    L.mrt_group_readback_wait.argtypes = [vp, u32]
    L.mrt_eval_sky_color.argtypes = [vp, f32p, vp, u32, vp]
    L.mrt_eval_bounce_stream.argtypes = [vp, u32, u32, u32, f32p, f32p, u32, vp]
    L.mrt_group_create.argtypes = [C.POINTER(C.c_int), u32, C.c_int, C.POINTER(vp)]
    L.mrt_group_unique_id.argtypes = [vp]
    L.mrt_group_create_rank.argtypes = [C.c_int, u32, u32, vp, C.POINTER(vp)]
    L.mrt_group_destroy.argtypes = [vp]
    L.mrt_group_destroy.restype = None
    L.mrt_group_last_error.argtypes = [vp]
    L.mrt_group_last_error.restype = C.c_char_p
    L.mrt_group_size.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.mrt_group_context.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u32)]
    L.mrt_group_set_tiles.argtypes = [vp, u32]
    L.mrt_group_render.argtypes = [vp, u32, u32, vp, vp, u32, u32, u32, u32]
    L.mrt_group_tonemap.argtypes = [vp, C.c_int, C.c_float, f32p, u32, C.c_int]
    L.mrt_group_gather.argtypes = [vp, C.c_int, u32]
    L.mrt_group_reduce.argtypes = [vp, u32]
    L.mrt_group_result.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp)]
    L.mrt_group_readback.argtypes = [vp, vp, C.c_size_t]
    L.mrt_group_sync.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mrt_destroy", "mrt_last_error", "mrt_group_destroy", "mrt_group_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Context:
    """One mrt_context: device memory, stream and scene state on one GPU."""

    _DTYPES = {BUF_VISIBILITY: (np.uint32, 1), BUF_DEPTH: (np.uint16, 1), BUF_NORMAL: (np.uint16, 4),
               BUF_MOTION: (np.uint16, 2), BUF_COLOR: (np.uint16, 4), BUF_ACCUM: (np.float32, 4),
               BUF_LDR: (np.uint8, 4), BUF_HIT_T: (np.float32, 1), BUF_DENOISED: (np.uint8, 4),
               BUF_TEMPORAL: (np.float32, 4), BUF_TEMPORAL_COUNT: (np.float32, 1)}

    def __init__(self, device=0, _borrowed=None):
        self.L = load()
        self.device = device
        self.size = (0, 0)
        self._owned = _borrowed is None
        if _borrowed is not None:   # a context that belongs to an mrt_group
            self.h = C.c_void_p(_borrowed)
            return
        self.h = C.c_void_p()
        s = self.L.mrt_create(device, C.byref(self.h))
        if s != 0:
            raise MinoteError(f"mrt_create({device}) failed ({s}): {self.L.mrt_last_error(None).decode()}")

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.L.mrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, s):
        if s != 0:
            raise MinoteError(f"minotert error {s}: {self.L.mrt_last_error(self.h).decode()}")

    # ---- inputs
    def set_option(self, name, value):
        self._ck(self.L.mrt_set_option(self.h, name.encode(), int(value)))

    def upload_blue_noise(self, rgba8):
        a = np.ascontiguousarray(rgba8, np.uint8)
        self._ck(self.L.mrt_upload_blue_noise(self.h, _ptr(a), a.shape[1], a.shape[0]))

    def set_spheres(self, spheres):
        arr = (Sphere * max(1, len(spheres)))()
        for i, (c, r, al) in enumerate(spheres):
            arr[i].center[:] = c
            arr[i].radius = r
            arr[i].albedo[:] = al
        self._ck(self.L.mrt_scene_set_spheres(self.h, C.cast(arr, C.c_void_p), len(spheres)))

    def upload_mesh(self, positions, indices, albedo):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        a = np.ascontiguousarray(albedo, np.float32).reshape(-1, 3)
        assert a.shape[0] == i.shape[0]
        self._ck(self.L.mrt_scene_upload_mesh(self.h, _ptr(p), p.shape[0], _ptr(i), i.shape[0], _ptr(a)))

    def update_positions(self, positions):
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self._ck(self.L.mrt_scene_update_positions(self.h, _ptr(p), p.shape[0]))

    def build(self, mode=BUILD_FULL):
        self._ck(self.L.mrt_scene_build(self.h, mode))

    def share_scene(self, owner):
        """Render the mesh scene (triangles + built BVH) that `owner` holds (mrt_scene_share: frames in flight)."""
        self._ck(self.L.mrt_scene_share(self.h, owner.h))

    def atmosphere(self, params):
        self._ck(self.L.mrt_atmosphere(self.h, C.cast(C.byref(params), C.c_void_p)))

    def sky_view(self, probe_pos, sun_dir, sun_ill):
        self._ck(self.L.mrt_sky_view(self.h, _f3(probe_pos), _f3(sun_dir), _f3(sun_ill)))

    def sky_aerial_perspective(self, pc, camera_pos, sun_dir=(-0.435286462, 0.818654716, 0.374606609), sun_ill=(8.0, 8.0, 8.0)):
        """mrt_sky_aerial_perspective: the 32^3 camera volume for the camera of the primary constants pc."""
        self._ck(self.L.mrt_sky_aerial_perspective(self.h, C.cast(C.byref(pc), C.c_void_p), _f3(camera_pos), _f3(sun_dir), _f3(sun_ill)))

    def set_partition(self, rank, nranks, slab_rows=8):
        self._ck(self.L.mrt_set_partition(self.h, rank, nranks, slab_rows))

    def partition_rows(self, full_h):
        n = C.c_uint32()
        self._ck(self.L.mrt_partition_rows(self.h, full_h, None, C.byref(n)))
        rows = np.zeros(n.value, np.uint32)
        self._ck(self.L.mrt_partition_rows(self.h, full_h, _ptr(rows), C.byref(n)))
        return rows

    # ---- render calls
    def primary_rays(self, w, h, pc):
        self._ck(self.L.mrt_primary_rays(self.h, w, h, C.cast(C.byref(pc), C.c_void_p)))
        self.size = (w, h)

    def secondary_rays(self, sc, spp=8, bounces=8, flags=0):
        self._ck(self.L.mrt_secondary_rays(self.h, C.cast(C.byref(sc), C.c_void_p), spp, bounces, flags))

    def denoise_bilateral(self, params=BILATERAL_DEFAULT, near=0.001, frame=1):
        self._ck(self.L.mrt_denoise_bilateral(self.h, params[0], params[1], params[2], near, frame))

    def temporal_accumulate(self, max_history=32.0, reset=False):
        """mrt_temporal_accumulate: blend this frame into the reprojected history (BUF_TEMPORAL, BUF_TEMPORAL_COUNT)."""
        self._ck(self.L.mrt_temporal_accumulate(self.h, float(max_history), TEMPORAL_RESET if reset else 0))

    def tonemap(self, mode="amd", exposure=1.0, params=(16.0, 2.0, 1.0, 0.18, 0.18), source=BUF_COLOR):
        par = (C.c_float * 8)(*params)
        m = TONEMAP[mode] if isinstance(mode, str) else int(mode)
        self._ck(self.L.mrt_tonemap(self.h, m, exposure, par, len(params), source))

    # ---- outputs
    def sync(self):
        self._ck(self.L.mrt_sync(self.h))

    def buffer(self, buf):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.mrt_buffer(self.h, buf, C.byref(p), C.byref(n)))
        return p.value, n.value

    def readback(self, buf, out=None):
        _, nbytes = self.buffer(buf)
        if buf in self._DTYPES:
            dt, ch = self._DTYPES[buf]
            n = nbytes // (np.dtype(dt).itemsize * ch)
            w = self.size[0]
            shape = (n // w, w, ch) if ch > 1 else (n // w, w)
        elif buf in (BUF_BVH_NODES, BUF_BVH_TRIS):
            dt, shape = np.uint32, (nbytes // 4,)
        elif buf == BUF_SKY_VIEW:
            dt, shape = np.uint32, (108, 192)
        elif buf == BUF_TRANSMITTANCE:
            dt, shape = np.uint16, (64, 256, 4)
        elif buf == BUF_AERIAL:
            dt, shape = np.uint16, (32, 32, 32, 4)
        else:
            dt, shape = np.uint16, (32, 32, 4)
        if out is None:
            out = np.empty(shape, dt)
        self._ck(self.L.mrt_readback(self.h, buf, _ptr(out), nbytes))
        return out

    def readback_into(self, buf, host_ptr, nbytes):
        self._ck(self.L.mrt_readback(self.h, buf, host_ptr, nbytes))

    def readback_async(self, buf, host_ptr, nbytes):
        self._ck(self.L.mrt_readback_async(self.h, buf, host_ptr, nbytes))

    def readback_wait(self, frames_in_flight=0):
        self._ck(self.L.mrt_readback_wait(self.h, frames_in_flight))

    def stats(self):
        s = Stats()
        self._ck(self.L.mrt_stats_get(self.h, C.byref(s)))
        return s

    def stats_reset(self):
        self._ck(self.L.mrt_stats_reset(self.h))

    def stream(self):
        p = C.c_void_p()
        self._ck(self.L.mrt_stream(self.h, C.byref(p)))
        return p.value

    def trace_rays(self, origins, directions, brute_force=False):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        ids = np.empty(o.shape[0], np.uint32)
        t = np.empty(o.shape[0], np.float32)
        self._ck(self.L.mrt_trace_rays(self.h, _ptr(o), _ptr(d), o.shape[0], _ptr(ids), _ptr(t), int(brute_force)))
        return ids, t

    def accum_commit(self, src=None, flags=0):
        """mrt_accum_commit: MRT_BUF_ACCUM (+)= the frame `src` (default: this context) rendered with SECONDARY_FRAME_SUM."""
        self._ck(self.L.mrt_accum_commit(self.h, (src or self).h, flags))
        if src is not None:
            self.size = src.size

    def accum_restore(self, accum):
        """mrt_accum_restore: resume a progressive render from a dumped accumulator (readback(BUF_ACCUM))."""
        a = np.ascontiguousarray(accum, np.float32)
        self._ck(self.L.mrt_accum_restore(self.h, _ptr(a), a.nbytes))

    def eval_sky_color(self, camera_pos, directions):
        """skyColor() evaluated on the GPU for an (n, 3) array of directions (mrt_eval_sky_color)."""
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        out = np.empty_like(d)
        self._ck(self.L.mrt_eval_sky_color(self.h, _f3(camera_pos), _ptr(d), d.shape[0], _ptr(out)))
        return out

    def eval_bounce_stream(self, frame, x, y, position, normal, n):
        """n consecutive bounces of pixel (x, y): (n, 9) = r0, r1, origin, direction, rng state bits."""
        out = np.empty((n, 9), np.float32)
        self._ck(self.L.mrt_eval_bounce_stream(self.h, frame, x, y, _f3(position), _f3(normal), n, _ptr(out)))
        return out


class Group:
    """mrt_group: n contexts (replicated scene) + the exchange step of the finished image, all inside the library
    (NCCL called from C++; no torch.distributed on the data path).

    Group(devices=[0, 1, ...])                 one process drives the listed devices (ncclCommInitAll)
    Group(devices=[0, 0, 0], transport="p2p")  device-to-device copies instead of NCCL; contexts may share a GPU
    Group.rank_of(device, rank, nranks, uid)   one process per GPU, uid = Group.unique_id() made on rank 0
    """

    def __init__(self, devices=None, transport="nccl", _handle=None, _device=None):
        self.L = load()
        self.h = C.c_void_p()
        if _handle is not None:
            self.h = _handle
            devices = [_device]
        else:
            arr = (C.c_int * len(devices))(*devices)
            s = self.L.mrt_group_create(arr, len(devices), GROUP_P2P if transport == "p2p" else GROUP_NCCL, C.byref(self.h))
            if s != 0:
                raise MinoteError(f"mrt_group_create failed ({s}): {self.L.mrt_group_last_error(None).decode()}")
        n, nl = C.c_uint32(), C.c_uint32()
        self._ck(self.L.mrt_group_size(self.h, C.byref(n), C.byref(nl)))
        self.nranks, self.nlocal = n.value, nl.value
        self.contexts, self.ranks = [], []
        for i in range(self.nlocal):
            c, r = C.c_void_p(), C.c_uint32()
            self._ck(self.L.mrt_group_context(self.h, i, C.byref(c), C.byref(r)))
            self.contexts.append(Context(device=devices[i], _borrowed=c.value))
            self.ranks.append(r.value)

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        L = load()
        s = L.mrt_group_unique_id(buf)
        if s != 0:
            raise MinoteError(f"mrt_group_unique_id failed ({s}): {L.mrt_group_last_error(None).decode()}")
        return bytes(buf)

    @classmethod
    def rank_of(cls, device, rank, nranks, unique_id):
        L = load()
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        s = L.mrt_group_create_rank(device, rank, nranks, buf, C.byref(h))
        if s != 0:
            raise MinoteError(f"mrt_group_create_rank failed ({s}): {L.mrt_group_last_error(None).decode()}")
        return cls(_handle=h, _device=device)

    def _ck(self, s):
        if s != 0:
            raise MinoteError(f"minotert group error {s}: {self.L.mrt_group_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            for c in self.contexts:
                c.close()
            self.L.mrt_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tiles(self, slab_rows=8):
        self._ck(self.L.mrt_group_set_tiles(self.h, slab_rows))

    def set_frames_in_flight(self, frames):
        """mrt_group_set_frames_in_flight: renders that carry SECONDARY_FRAME_SUM go round-robin over `frames` frame
        contexts per rank.  Returns them as [local rank][slot] (borrowed Context objects): the caller prepares each like the
        rank's own context -- upload_blue_noise, share_scene(rank's context), atmosphere, sky_view."""
        self._ck(self.L.mrt_group_set_frames_in_flight(self.h, frames))
        self.frame_contexts = []
        for i in range(self.nlocal):
            slots = []
            for k in range(frames):
                c = C.c_void_p()
                self._ck(self.L.mrt_group_frame_context(self.h, i, k, C.byref(c)))
                slots.append(self.contexts[i] if frames == 1 else Context(device=self.contexts[i].device, _borrowed=c.value))
            self.frame_contexts.append(slots)
        return self.frame_contexts

    def render(self, w, h, pc, sc, spp, bounces, flags=0, frame_stride=0):
        self._ck(self.L.mrt_group_render(self.h, w, h, C.cast(C.byref(pc), C.c_void_p), C.cast(C.byref(sc), C.c_void_p),
                                         spp, bounces, flags, frame_stride))
        for c in self.contexts:
            c.size = (w, h)
        self.size = (w, h)

    def tonemap(self, mode="amd", exposure=1.0, params=(16.0, 2.0, 1.0, 0.18, 0.18), source=BUF_ACCUM):
        par = (C.c_float * 8)(*params)
        m = TONEMAP[mode] if isinstance(mode, str) else int(mode)
        self._ck(self.L.mrt_group_tonemap(self.h, m, exposure, par, len(params), source))

    def gather(self, buf=BUF_LDR, root=0):
        self._ck(self.L.mrt_group_gather(self.h, buf, root))
        self._gathered = buf

    def reduce(self, root=0):
        self._ck(self.L.mrt_group_reduce(self.h, root))

    def result(self):
        p, n, s = C.c_void_p(), C.c_size_t(), C.c_void_p()
        self._ck(self.L.mrt_group_result(self.h, C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value

    def readback(self, out=None):
        """The gathered image on the root as a numpy array in image order (blocking)."""
        dt, ch = Context._DTYPES[self._gathered]
        w, h = self.size
        if out is None:
            out = np.empty((h, w, ch) if ch > 1 else (h, w), dt)
        self._ck(self.L.mrt_group_readback(self.h, _ptr(out), out.nbytes))
        return out

    def readback_into(self, host_ptr, nbytes):
        self._ck(self.L.mrt_group_readback(self.h, host_ptr, nbytes))

    def readback_async(self, host_ptr, nbytes):
        self._ck(self.L.mrt_group_readback_async(self.h, host_ptr, nbytes))

    def readback_wait(self, keep_in_flight=0):
        self._ck(self.L.mrt_group_readback_wait(self.h, keep_in_flight))

    def sync(self):
        self._ck(self.L.mrt_group_sync(self.h))
