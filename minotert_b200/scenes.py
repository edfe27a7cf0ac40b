"""Deterministic procedural scenes for the BASELINE.json configs (synthetic inputs).

Only integer hashing and +,-,*,/,sqrt in float64 are used (no libm transcendentals), so the same
arrays come out on every machine.  Units are kilometres like the reference scene
(src/gpu/scene.glsl): 1 m = 0.001, and the path tracer's ray offset is 1e-6 (1 mm).

Each generator returns (positions float32 [V,3], indices uint32 [T,3], albedo float32 [T,3], view)
where `view` = dict(position, yaw_deg, pitch_deg) is a camera that looks at the scene.
"""
import numpy as np

M = 0.001  # one metre in scene units (km)
GROUND_Z = 0.0985


def hash_u32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def hash01(x):
    return hash_u32(x).astype(np.float64) / 4294967296.0


def _lattice(ix, iy, iz, seed):
    h = hash_u32(ix.astype(np.uint32) * np.uint32(73856093) ^ iy.astype(np.uint32) * np.uint32(19349663)
                 ^ iz.astype(np.uint32) * np.uint32(83492791) ^ np.uint32(seed * 2654435761 & 0xFFFFFFFF))
    return h.astype(np.float64) / 4294967296.0


def value_noise(p, seed):
    """Trilinear value noise in [0,1) with smoothstep weights; p float64 [...,3]."""
    f = np.floor(p)
    t = p - f
    t = t * t * (3.0 - 2.0 * t)
    i = f.astype(np.int64)
    out = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (t[..., 0] if dx else 1 - t[..., 0]) * (t[..., 1] if dy else 1 - t[..., 1]) * \
                    (t[..., 2] if dz else 1 - t[..., 2])
                out = out + w * _lattice(i[..., 0] + dx, i[..., 1] + dy, i[..., 2] + dz, seed)
    return out


def fbm(p, seed, octaves=4):
    amp, total, norm = 1.0, 0.0, 0.0
    for o in range(octaves):
        total = total + amp * value_noise(p * (2.0 ** o), seed + 101 * o)
        norm += amp
        amp *= 0.5
    return total / norm


def tri_albedo(ntris, seed, lo=0.2, hi=0.9, group=1):
    ids = (np.arange(ntris, dtype=np.uint32) // np.uint32(group)) * np.uint32(3) + np.uint32((seed * 7919) & 0xFFFFFFFF)
    a = np.stack([hash01(ids), hash01(ids + np.uint32(1)), hash01(ids + np.uint32(2))], axis=1)
    return (lo + (hi - lo) * a).astype(np.float32)


def grid_indices(nx, ny, base=0):
    """Two triangles per cell of an (nx+1) x (ny+1) vertex grid, row-major vertices."""
    i, j = np.meshgrid(np.arange(nx, dtype=np.uint32), np.arange(ny, dtype=np.uint32), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel() + base
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    t0 = np.stack([v00, v10, v11], axis=1)
    t1 = np.stack([v00, v11, v01], axis=1)
    return np.stack([t0, t1], axis=1).reshape(-1, 3).astype(np.uint32)


def terrain_height(x, y, seed):
    p = np.stack([x / (12.0 * M), y / (12.0 * M), np.zeros_like(x)], axis=-1)
    return GROUND_Z + 3.0 * M * (fbm(p, seed, 5) - 0.5)


def terrain(n, size, seed):
    u = np.linspace(-0.5, 0.5, n + 1)
    X, Y = np.meshgrid(u * size, u * size + 0.5 * size + 2.0 * M, indexing="xy")
    Z = terrain_height(X, Y, seed)
    pos = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    return pos, grid_indices(n, n)


def cube_sphere(res):
    """Unit cube-sphere: 6 faces x res x res x 2 triangles; seams share exact coordinates."""
    u = np.linspace(-1.0, 1.0, res + 1)
    A, B = np.meshgrid(u, u, indexing="xy")
    one = np.ones_like(A)
    faces = [(one, A, B), (-one, B, A), (B, one, A), (A, -one, B), (A, B, one), (B, A, -one)]
    pos, idx = [], []
    for f, (x, y, z) in enumerate(faces):
        p = np.stack([x, y, z], axis=-1).reshape(-1, 3)
        p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
        pos.append(p)
        idx.append(grid_indices(res, res, base=f * (res + 1) ** 2))
    return np.concatenate(pos), np.concatenate(idx)


def blobs(count, res, size, seed):
    unit_p, unit_i = cube_sphere(res)
    nv = unit_p.shape[0]
    pos, idx = [], []
    for b in range(count):
        k = np.uint32(seed * 1000003 + b * 16)
        r = (0.7 + 2.0 * hash01(k)) * M
        cx = (hash01(k + np.uint32(1)) - 0.5) * size * 0.9
        cy = (hash01(k + np.uint32(2)) - 0.5) * size * 0.9 + 0.5 * size + 2.0 * M
        ground = terrain_height(np.array([cx]), np.array([cy]), seed)[0]
        cz = ground + r * 0.6 + hash01(k + np.uint32(3)) * 5.0 * M
        disp = 1.0 + 0.35 * (fbm(unit_p * 1.7 + float(b), seed + 7, 3) - 0.5)
        pos.append(unit_p * (r * disp)[:, None] + np.array([cx, cy, cz]))
        idx.append(unit_i + np.uint32(b * nv))
    return np.concatenate(pos), np.concatenate(idx)


def terrain_blobs(n_grid, n_blobs, blob_res, seed, size=60.0 * M):
    tp, ti = terrain(n_grid, size, seed)
    bp, bi = blobs(n_blobs, blob_res, size, seed)
    pos = np.concatenate([tp, bp]).astype(np.float32)
    idx = np.concatenate([ti, bi + np.uint32(tp.shape[0])]).astype(np.uint32)
    albedo = np.concatenate([tri_albedo(ti.shape[0], seed, 0.35, 0.75, group=2),
                             tri_albedo(bi.shape[0], seed + 1, 0.2, 0.9, group=2 * blob_res * blob_res)])
    view = dict(position=(0.0, -4.0 * M, GROUND_Z + 9.0 * M), yaw_deg=90.0, pitch_deg=-12.0)
    return pos, idx, albedo, view


def hall_260k(seed=2):
    """BASELINE config 2: 'Sponza-scale' ~260k triangles (256x256 terrain grid + 64 blobs)."""
    return terrain_blobs(256, 64, 13, seed)


def scene_1m(seed=5):
    """north_star target scene: ~1M triangles (512x512 terrain + 256 blobs)."""
    return terrain_blobs(512, 256, 13, seed)


def scene_10m(seed=3):
    """BASELINE config 3: ~10M triangles (1024x1024 terrain + 4096 blobs), HBM-resident BVH."""
    return terrain_blobs(1024, 4096, 13, seed, size=120.0 * M)


def small_terrain(n_grid=24, n_blobs=3, blob_res=4, seed=11):
    """A few thousand triangles or fewer: brute-force-checkable variant of the same generator."""
    return terrain_blobs(n_grid, n_blobs, blob_res, seed, size=24.0 * M)


def _quad(p0, du, dv, nu, nv):
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    pos = p0[None, None, :] + s[..., None] * du[None, None, :] + t[..., None] * dv[None, None, :]
    return pos.reshape(-1, 3), grid_indices(nu, nv)


def cornell(seed=1, tess=8):
    """BASELINE config 1: Cornell-box-class scene, ~1k triangles, open front, sky-lit."""
    S = 4.0 * M
    c = np.array([0.0, 3.0 * M, GROUND_Z])  # front-bottom centre of the room
    x0, x1, y0, y1, z0, z1 = c[0] - S / 2, c[0] + S / 2, c[1], c[1] + S, c[2], c[2] + S
    V = np.array
    quads = [  # (origin, du, dv, colour)
        (V([x0, y0, z0]), V([S, 0, 0]), V([0, S, 0]), (0.73, 0.73, 0.73)),   # floor
        (V([x0, y0, z1]), V([S, 0, 0]), V([0, S, 0]), (0.73, 0.73, 0.73)),   # ceiling
        (V([x0, y1, z0]), V([S, 0, 0]), V([0, 0, S]), (0.73, 0.73, 0.73)),   # back
        (V([x0, y0, z0]), V([0, S, 0]), V([0, 0, S]), (0.65, 0.05, 0.05)),   # left, red
        (V([x1, y0, z0]), V([0, S, 0]), V([0, 0, S]), (0.12, 0.45, 0.15)),   # right, green
    ]
    pos, idx, alb, base = [], [], [], 0

    def add(p, i, colour):
        nonlocal base
        pos.append(p)
        idx.append(i + np.uint32(base))
        alb.append(np.tile(np.array(colour, np.float32), (i.shape[0], 1)))
        base += p.shape[0]

    for o, du, dv, col in quads:
        p, i = _quad(o, du, dv, tess, tess)
        add(p, i, col)
    # two boxes (5 faces each: no bottom), rotated by exact rational sines/cosines (3-4-5, 5-12-13)
    for (bx, by, w, h, cs, sn, col) in [(-0.7 * M, 2.6 * M, 1.2 * M, 2.4 * M, 0.8, 0.6, (0.7, 0.7, 0.75)),
                                        (0.8 * M, 1.2 * M, 1.2 * M, 1.2 * M, 12 / 13, -5 / 13, (0.75, 0.7, 0.6))]:
        ex = V([cs, sn, 0.0]) * w
        ey = V([-sn, cs, 0.0]) * w
        ez = V([0.0, 0.0, 1.0]) * h
        o = V([c[0] + bx, c[1] + by, z0]) - 0.5 * ex - 0.5 * ey
        for (fo, fu, fv) in [(o + ez, ex, ey), (o, ex, ez), (o + ey, ex, ez), (o, ey, ez), (o + ex, ey, ez)]:
            p, i = _quad(fo, fu, fv, 4, 4)
            add(p, i, col)
    pos = np.concatenate(pos).astype(np.float32)
    idx = np.concatenate(idx).astype(np.uint32)
    alb = np.concatenate(alb).astype(np.float32)
    jitter = tri_albedo(idx.shape[0], seed, 0.9, 1.0)
    view = dict(position=(0.0, -1.5 * M, GROUND_Z + 2.0 * M), yaw_deg=90.0, pitch_deg=0.0)
    return pos, idx, (alb * jitter).astype(np.float32), view


def animate(positions, time, amplitude=0.4 * M):
    """BASELINE config 5: vertex displacement along z from hashed phases (no libm: triangle wave)."""
    p = np.asarray(positions, np.float64)
    phase = (p[:, 0] + p[:, 1]) / (8.0 * M) + time
    tri = np.abs((phase - np.floor(phase)) * 2.0 - 1.0) * 2.0 - 1.0
    out = p.copy()
    out[:, 2] += amplitude * tri
    return out.astype(np.float32)
