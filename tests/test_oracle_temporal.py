"""Known-answer tests for the oracle's temporal reprojection.  The reference has no such
stage (it writes GBuffer::motion, src/gpu/primaryRay.comp:73-75, and never reads it), so the contract is the one in
oracle/minote_oracle.h; these answers follow from that contract by hand."""
import numpy as np

MISS = 0xFFFFFFFF


def f16(a):
    return np.ascontiguousarray(np.asarray(a, np.float16)).view(np.uint16)


def frame(w, h, rgb, samples=1.0, ids=7, motion=(0.0, 0.0)):
    acc = np.zeros((h, w, 4), np.float32)
    acc[..., :3] = np.asarray(rgb, np.float32) * np.float32(samples)
    acc[..., 3] = samples
    vis = np.full((h, w), ids, np.uint32) if np.isscalar(ids) else np.asarray(ids, np.uint32)
    mo = np.zeros((h, w, 2), np.float32)
    mo[...] = motion
    return acc, vis, f16(mo)


def test_first_frame_is_the_frame_average(oracle):
    w, h = 9, 5
    rng = np.random.default_rng(1)
    rgb = rng.uniform(0, 4, (h, w, 3)).astype(np.float32)
    acc, vis, mo = frame(w, h, rgb, samples=4.0)
    out, cnt, hv = oracle.temporal_accumulate(acc, vis, mo, None)
    assert np.array_equal(out[..., :3], acc[..., :3] / np.float32(4.0))
    assert np.all(out[..., 3] == 1.0) and np.all(cnt == 1.0) and np.array_equal(hv, vis)
    # an empty accumulator (w = 0) resolves to black, as mrt's tonemap source does
    acc[0, 0] = 0
    out, _, _ = oracle.temporal_accumulate(acc, vis, mo, None)
    assert np.all(out[0, 0, :3] == 0)


def test_static_camera_is_a_running_mean_up_to_the_cap(oracle):
    w, h = 6, 4
    hist, vals = None, []
    for k in range(1, 8):
        v = np.float32(k * 0.5)
        vals.append(v)
        acc, vis, mo = frame(w, h, (v, 2 * v, 0.25))
        hist = oracle.temporal_accumulate(acc, vis, mo, hist, max_history=4.0)
        out, cnt, _ = hist
        if k <= 5:  # history length k-1 <= 4: plain mean of the k frames
            assert np.all(cnt == k)
            assert np.allclose(out[..., 0], np.mean(vals), rtol=1e-6) and np.allclose(out[..., 1], 2 * np.mean(vals), rtol=1e-6)
        else:       # capped: exponential moving average with weight 1 / (4 + 1)
            assert np.all(cnt == 5.0)
            assert np.allclose(out[..., 0], prev + (v - prev) * np.float32(0.2), rtol=1e-6)
        assert np.all(out[..., 2] == 0.25)
        prev = out[0, 0, 0]


def test_integer_and_half_pixel_motion(oracle):
    """motion = (ndc_now - ndc_prev) * size = 2 x the pixel displacement, y flipped (primaryRay.comp:46,75):
    the previous position of pixel (x, y) is (x - m.x / 2, y + m.y / 2)."""
    w, h = 12, 7
    ramp = (np.arange(w, dtype=np.float32)[None, :] + 100 * np.arange(h, dtype=np.float32)[:, None])
    hist_rgb = np.stack([ramp, ramp, ramp], -1)
    acc0, vis, mo0 = frame(w, h, hist_rgb)
    hist = oracle.temporal_accumulate(acc0, vis, mo0, None)
    # content moved 3 px right and 1 px down on screen: pixel (x, y) was at (x - 3, y - 1) -> motion (+6, -2)... y: prev = y + m.y/2
    acc, _, mo = frame(w, h, np.zeros(3), motion=(6.0, -2.0))
    out, cnt, _ = oracle.temporal_accumulate(acc, vis, mo, hist)
    inside = np.zeros((h, w), bool)
    inside[1:, 3:] = True
    assert np.all(cnt[inside] == 2.0) and np.all(cnt[~inside] == 1.0)
    want = ramp[:-1, :-3] * np.float32(0.5)  # mean of the history value and the black current frame
    assert np.array_equal(out[1:, 3:, 0], want)
    assert np.all(out[~inside][:, :3] == 0)
    # half a pixel to the left: prev = x + 0.5 -> taps x and x+1, weights 1/2
    acc, _, mo = frame(w, h, np.zeros(3), motion=(-1.0, 0.0))
    out, cnt, _ = oracle.temporal_accumulate(acc, vis, mo, hist)
    assert np.array_equal(out[:, :-1, 0], (ramp[:, :-1] + ramp[:, 1:]) * np.float32(0.5) * np.float32(0.5))
    # last column: the right tap is outside the image, the weights renormalise to the left tap alone
    assert np.array_equal(out[:, -1, 0], ramp[:, -1] * np.float32(0.5)) and np.all(cnt == 2.0)


def test_taps_on_another_primitive_are_rejected(oracle):
    w, h = 8, 3
    ids = np.full((h, w), 5, np.uint32)
    ids[:, 4:] = 9
    rgb = np.zeros((h, w, 3), np.float32)
    rgb[:, :4] = 1.0
    rgb[:, 4:] = 3.0
    acc0, _, mo0 = frame(w, h, rgb, ids=ids)
    hist = oracle.temporal_accumulate(acc0, ids, mo0, None)
    # half-pixel motion: pixel 3 (id 5) would mix texels 3 (id 5) and 4 (id 9): only texel 3 counts
    acc, _, mo = frame(w, h, rgb, ids=ids, motion=(-1.0, 0.0))
    out, cnt, _ = oracle.temporal_accumulate(acc, ids, mo, hist)
    assert np.all(out[:, 3, 0] == 1.0) and np.all(out[:, 4, 0] == 3.0) and np.all(cnt == 2.0)
    # every tap on another primitive (disocclusion): history dropped
    ids2 = ids.copy()
    ids2[1, 2] = 77
    acc, _, mo = frame(w, h, rgb * 2, ids=ids2)
    out, cnt, _ = oracle.temporal_accumulate(acc, ids2, mo, hist)
    assert cnt[1, 2] == 1.0 and out[1, 2, 0] == 2.0
    assert cnt[1, 1] == 2.0 and out[1, 1, 0] == 1.5
    # primary misses pass through (noise-free sky, motion 0 by definition), even over a miss history
    ids3 = np.full((h, w), MISS, np.uint32)
    acc, _, mo = frame(w, h, rgb * 4, ids=ids3)
    h3 = oracle.temporal_accumulate(acc, ids3, mo, hist)
    out, cnt, _ = oracle.temporal_accumulate(acc, ids3, mo, h3)
    assert np.all(cnt == 1.0) and np.array_equal(out[..., :3], rgb * 4)


def test_out_of_image_and_non_finite_motion_reset(oracle):
    w, h = 5, 4
    acc0, vis, mo0 = frame(w, h, (1.0, 1.0, 1.0))
    hist = oracle.temporal_accumulate(acc0, vis, mo0, None)
    for m in ((40.0, 0.0), (0.0, -60.0), (np.inf, 0.0), (np.nan, np.nan), (-np.inf, np.inf)):
        acc, _, mo = frame(w, h, (2.0, 2.0, 2.0), motion=m)
        out, cnt, _ = oracle.temporal_accumulate(acc, vis, mo, hist)
        assert np.all(cnt == 1.0) and np.all(out[..., :3] == 2.0), m
