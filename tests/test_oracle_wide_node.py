"""The compressed 8-wide BVH node of the product (north_star row n3): quantiser (bvh_build.cu k_emit_nodes) and
fp32 slab test (trace.cuh lane_node_step), restated operation for operation in the oracle so that the claim the
traversal rests on is checked on the CPU: *a child box the exact ray touches is never culled* -- in spite of 8-bit
planes, the decode folded into one FMA (error <= 1/256 step), fp32 rounding and the (1 -/+ 2^-21) scaled 1/d.
The exact side is the ray/box slab test in float64 on the very fp32 ray and fp32 child boxes.  The GPU proves the
same end to end (BVH hits == brute-force hits, tests/test_gpu_mesh.py); this test covers grazing rays, axis-parallel
rays, flat boxes and flat nodes, empty slots and scenes far from the origin far more densely."""
import numpy as np
import pytest


def exact_hits(lo, hi, present, o, d, t_best):
    """(n, 8) bool: the fp32 ray o + t d, t in [0, t_best], touches the closed fp32 box (float64 arithmetic)."""
    o = o.astype(np.float64)[:, None, :]
    d = d.astype(np.float64)[:, None, :]
    lo = lo.astype(np.float64)[None, :, :]
    hi = hi.astype(np.float64)[None, :, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t1, t2), np.maximum(t1, t2)
    par = d == 0.0
    inside = (o >= lo) & (o <= hi)
    tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
    tenter, texit = tn.max(-1), tf.min(-1)
    ok = (tenter <= texit) & (texit >= 0.0) & (tenter <= t_best.astype(np.float64)[:, None])
    return ok & np.array([(present >> s) & 1 for s in range(8)], bool)[None, :]


def random_node(rng, scale, offset):
    centre = (offset + rng.uniform(-1, 1, 3) * scale).astype(np.float32)
    ext = scale * 10.0 ** rng.uniform(-3, 0, (8, 3))
    c = centre + rng.uniform(-1, 1, (8, 3)) * scale
    lo, hi = (c - ext).astype(np.float32), (c + ext).astype(np.float32)
    flat = rng.random((8, 3)) < 0.08           # flat boxes (axis-aligned triangles)
    hi = np.where(flat, lo, hi)
    hi = np.maximum(lo, hi)
    present = int(rng.integers(1, 256))
    return lo, hi, present


def rays_for(rng, lo, hi, present, n):
    slots = [s for s in range(8) if present >> s & 1]
    nlo, nhi = lo[slots].min(0), hi[slots].max(0)
    size = np.maximum(nhi - nlo, 1e-30)
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    tb = np.full(n, 3.0e38, np.float32)
    for i in range(n):
        kind = i % 8
        s = slots[int(rng.integers(len(slots)))]
        if kind == 0:    # origin inside a child box
            oo = lo[s] + rng.random(3) * (hi[s] - lo[s])
        elif kind == 1:  # far away
            oo = nlo + (rng.uniform(-1, 2, 3) * 1000.0) * size
        else:            # around the node
            oo = nlo + rng.uniform(-3, 4, 3) * size
        oo = oo.astype(np.float32)
        # aim at a point of a child box: interior, face, edge or corner (grazing rays)
        u = rng.random(3)
        snap = rng.random(3) < (0.0, 0.3, 0.6, 0.9, 0.5, 0.5, 0.5, 0.5)[kind]
        u = np.where(snap, np.round(u), u)
        target = (lo[s] + u * (hi[s] - lo[s])).astype(np.float32)
        dd = (target - oo).astype(np.float32)
        if not np.any(dd):
            dd = rng.normal(size=3).astype(np.float32)
        if kind == 4:    # axis-parallel: exact zeros, unit length (the product's rays are normalised directions)
            keep = int(rng.integers(3))
            dd = np.where(np.arange(3) == keep, 1.0 if dd[keep] >= 0 else -1.0, 0.0).astype(np.float32)
        elif kind == 5:  # nearly axis-parallel, about unit length
            nrm = np.float32(np.sqrt(np.float32(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2])))
            dd = (dd / nrm).astype(np.float32) if nrm > 0 and np.isfinite(nrm) else np.float32([1, 0, 0])
            z = int(rng.integers(3))
            dd[z] = np.float32(rng.choice([1e-30, -1e-30, 1e-12, -1e-12, 1e-6, -1e-6]))
            if not np.any(np.abs(dd) > 0.1):
                dd[(z + 1) % 3] = 1.0
        else:
            nrm = np.float32(np.sqrt(np.float32(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2])))
            if nrm > 0 and np.isfinite(nrm):
                dd = (dd / nrm).astype(np.float32)
        o[i], d[i] = oo, dd
        if kind in (6, 7):  # a closest hit already found: somewhere along the ray near the targeted box
            with np.errstate(all="ignore"):
                k = np.argmax(np.abs(dd))
                t = (np.float64(target[k]) - np.float64(oo[k])) / np.float64(dd[k])
            if np.isfinite(t) and t > 0:
                tb[i] = np.float32(t * rng.uniform(0.5, 1.5))
    return o, d, tb


# (node size, distance from the origin): the last five are scenes whose nodes are only tens of ulps of their
# coordinates wide -- where an fp32 evaluation of origin + q * step (the builder's first version) rounded planes INTO
# child boxes and culled 0.2 % of the grazing hits; the quantiser now evaluates planes exactly and never makes the
# step finer than 2 ulp of the coordinates
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (1e-3, 0.1), (1.0, 300.0), (1e-2, 50.0), (1e-4, 0.1),
                                          (1e-3, 1000.0), (1.0, 1e6), (1e-5, 1.0), (1e-6, 100.0), (1e4, 1e7)])
def test_fp32_slab_test_never_culls_an_exact_hit(oracle, scale, offset):
    rng = np.random.default_rng(int(scale * 1e6) + int(offset * 10) + 7)
    nodes, rays_per_node = 300, 256
    exact_n = kernel_n = 0
    for _ in range(nodes):
        lo, hi, present = random_node(rng, scale, offset)
        node = oracle.wide_node_quantize(lo, hi, present)
        o, d, tb = rays_for(rng, lo, hi, present, rays_per_node)
        got = oracle.wide_node_test(node, o, d, tb)
        got = (got[:, None] >> np.arange(8)[None, :]) & 1 == 1
        want = exact_hits(lo, hi, present, o, d, tb)
        culled = want & ~got
        assert not culled.any(), (f"{culled.sum()} exact hits culled; first: ray {np.argwhere(culled)[0]}, o={o[np.argwhere(culled)[0][0]]}, "
                                  f"d={d[np.argwhere(culled)[0][0]]}, box lo={lo[np.argwhere(culled)[0][1]]} hi={hi[np.argwhere(culled)[0][1]]}")
        empty = np.array([not (present >> s) & 1 for s in range(8)])
        assert not got[:, empty].any(), "an empty slot was hit"
        exact_n += int(want.sum())
        kernel_n += int(got.sum())
    assert exact_n > nodes * rays_per_node * 0.3           # the rays do aim at the boxes
    # conservative, but not wastefully so: on well-scaled nodes 8-bit planes + slack cost a bounded number of extra
    # child visits (these random children are as small as 1/1000 of their node, far smaller than a BVH's)
    if offset <= 300.0 * scale:
        assert kernel_n <= exact_n * 1.4, (exact_n, kernel_n)


def test_known_node(oracle):
    """Hand-checkable node: two unit-ish boxes on the x axis; grid step = 2^-6 (250 steps >= the extent 3)."""
    lo = np.zeros((8, 3), np.float32)
    hi = np.zeros((8, 3), np.float32)
    lo[0], hi[0] = (0, 0, 0), (1, 1, 1)
    lo[1], hi[1] = (2, 0, 0), (3, 1, 1)
    node = oracle.wide_node_quantize(lo, hi, 0b11)
    ex, ey, ez = node.w[3] & 0xFF, (node.w[3] >> 8) & 0xFF, (node.w[3] >> 16) & 0xFF
    assert (ex, ey, ez) == (127 - 6, 127 - 7, 127 - 7)      # 3/250 -> 2^-6, 1/250 -> 2^-7
    org = np.frombuffer(np.array(node.w[:3], np.uint32).tobytes(), np.float32)
    assert np.array_equal(org, np.float32([-2 / 64, -2 / 128, -2 / 128]))
    qlo_x, qhi_x = node.w[4] & 0xFFFF, node.w[10] & 0xFFFF
    # planes: box 0 x in [0, 1] -> q in [2 - 1, 2 + 64 + 1] (one step outward: 1/64 step of slack), box 1 x in [2, 3]
    assert (qlo_x & 0xFF, qhi_x & 0xFF) == (1, 67) and (qlo_x >> 8, qhi_x >> 8) == (129, 195)
    for s in range(2, 8):                                     # empty slots: (255, 0) on every axis
        for a in range(3):
            assert (node.w[4 + 2 * a + (s >> 2)] >> (8 * (s & 3))) & 0xFF == 255
            assert (node.w[10 + 2 * a + (s >> 2)] >> (8 * (s & 3))) & 0xFF == 0
    o = np.float32([[-1, 0.5, 0.5], [-1, 0.5, 0.5], [1.5, 0.5, 0.5], [1.5, 5.0, 0.5], [-1, 0.5, 0.5]])
    d = np.float32([[1, 0, 0], [-1, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0]])
    tb = np.float32([3e38, 3e38, 3e38, 3e38, 2.5])
    assert list(oracle.wide_node_test(node, o, d, tb)) == [0b11, 0b00, 0b10, 0b00, 0b01]


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.1), (1e-2, 50.0), (1e-3, 1000.0), (1.0, 1e6)])
def test_every_triangle_hit_passes_its_leaf_box(oracle, scale, offset):
    """The invariant the image depends on, at the leaf level: whenever the fp32 watertight triangle test (row n4, the
    same code the brute-force reference runs) reports a hit, the quantised slab test lets the ray into the triangle's
    leaf box -- with no closest hit yet and with the closest hit AT this triangle (tlimit = t * (1 + 2^-21)).  Rays
    are aimed at the triangle's interior, edges and vertices and displaced by a few ulps, from near and from far;
    triangles include axis-aligned (flat-box) ones; siblings share the node."""
    rng = np.random.default_rng(int(scale * 1e6) + int(offset) + 3)
    hits_total = 0
    for _ in range(150):
        centre = (offset + rng.uniform(-1, 1, 3) * scale).astype(np.float32)
        tri = (centre + rng.uniform(-1, 1, (3, 3)) * scale * 10.0 ** rng.uniform(-2, 0)).astype(np.float32)
        if rng.random() < 0.3:                      # axis-aligned triangle: a flat leaf box
            tri[:, int(rng.integers(3))] = tri[0, int(rng.integers(3))]
        lo = np.zeros((8, 3), np.float32)
        hi = np.zeros((8, 3), np.float32)
        slot = int(rng.integers(8))
        present = 1 << slot
        lo[slot], hi[slot] = tri.min(0), tri.max(0)
        for s in range(8):                          # siblings widen the node (coarser grid for this leaf)
            if s != slot and rng.random() < 0.5:
                c = centre + rng.uniform(-1, 1, 3) * scale * 4
                e = scale * 10.0 ** rng.uniform(-2, 0, 3)
                lo[s], hi[s] = (c - e).astype(np.float32), (c + e).astype(np.float32)
                present |= 1 << s
        node = oracle.wide_node_quantize(lo, hi, present)
        n = 400
        w = rng.dirichlet((1, 1, 1), n)
        snap = rng.random(n)
        w[snap < 0.3] = np.eye(3)[rng.integers(0, 3, (snap < 0.3).sum())]                    # vertices
        edge = (snap >= 0.3) & (snap < 0.6)
        w[edge, rng.integers(0, 3)] = 0.0
        w[edge] /= np.maximum(w[edge].sum(1, keepdims=True), 1e-30)                          # edges
        target = (w @ tri.astype(np.float64))
        ulp = np.spacing(np.abs(target).astype(np.float32)).astype(np.float64)
        target = (target + rng.integers(-3, 4, (n, 3)) * ulp).astype(np.float32)             # a few ulps off
        ext = float(np.max(hi[slot] - lo[slot])) + float(np.spacing(np.float32(abs(offset) + scale)))
        dist = 10.0 ** rng.uniform(-1, 3, (n, 1)) * ext
        dirs = rng.normal(size=(n, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        o = (target.astype(np.float64) - dirs * dist).astype(np.float32)
        d = (target.astype(np.float64) - o.astype(np.float64))
        d = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)).astype(np.float32)
        hit, t = oracle.ray_triangle_batch(o, d, tri[0], tri[1], tri[2])
        if not hit.any():
            continue
        open_mask = oracle.wide_node_test(node, o, d, None)
        best_mask = oracle.wide_node_test(node, o, d, np.where(hit, t, np.float32(3.0e38)))
        for name, m in (("no closest hit yet", open_mask), ("closest hit at this triangle", best_mask)):
            lost = hit & ((m >> slot) & 1 == 0)
            assert not lost.any(), (f"{name}: {lost.sum()} of {hit.sum()} triangle hits culled at the leaf box; first: o={o[lost][0]}, "
                                    f"d={d[lost][0]}, t={t[lost][0]}, tri={tri.tolist()}")
        hits_total += int(hit.sum())
    assert hits_total > 2000
