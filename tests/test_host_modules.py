"""Host logic (C++20 modules behind libminote_host.so) against the oracle: the matrices that reach the
GPU must be bit-identical to the reference's host math.  CPU-only."""
import ctypes as C

import numpy as np
import pytest

from minotert_b200 import capi, host


def cam_pair(oracle, w, h, pos, yaw, pitch):
    return host.make_camera(w, h, pos, yaw, pitch), oracle.make_camera(w, h, pos, yaw, pitch)


@pytest.mark.parametrize("w,h,pos,yaw,pitch", [
    (960, 540, (0.0, -0.001, 0.1), 90.0, 0.0),           # src/app.ixx:20-32
    (1920, 1080, (0.0, -0.004, 0.1075), 90.0, -12.0),
    (512, 512, (0.013, 0.021, 0.1033), 217.5, 33.0),
    (3840, 2160, (-0.02, 0.05, 0.12), 1.0, -89.0),
])
def test_constants_bit_exact(oracle, w, h, pos, yaw, pitch):
    hc, oc = cam_pair(oracle, w, h, pos, yaw, pitch)
    assert bytes(hc) == bytes(oc)
    prev_h, prev_o = cam_pair(oracle, w, h, (pos[0] + 0.001, pos[1], pos[2]), yaw + 3.0, pitch)
    for frame in (1, 2, 77):
        hp, hs = host.camera_constants(hc, prev_h, frame)
        op, os_ = oracle.constants(oc, prev_o, frame)
        assert bytes(hp) == bytes(op), "primary constants differ from the oracle's restatement"
        assert bytes(hs) == bytes(os_), "secondary constants differ from the oracle's restatement"


def test_camera_controls_match_oracle(oracle):
    hc, oc = cam_pair(oracle, 960, 540, (0.0, -0.001, 0.1), 90.0, 0.0)
    L, O = host.load(), oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = (float(x) for x in rng.uniform(-40, 40, 2))
        L.minote_camera_rotate(C.byref(hc), a, b)
        O.orc_camera_rotate(C.byref(oc), a, b)
        d = [float(x) * 1e-5 for x in rng.uniform(-1, 1, 3)]
        L.minote_camera_roam(C.byref(hc), (C.c_float * 3)(*d))
        O.orc_camera_roam(C.byref(oc), oracle.f3(d))
        L.minote_camera_shift(C.byref(hc), (C.c_float * 3)(*d))
        O.orc_camera_shift(C.byref(oc), oracle.f3(d))
        assert bytes(hc) == bytes(oc)


def test_freecam_update(oracle):
    # src/freecam.ixx:51-68: moveSpeed = 0.0005 * min(frameTime, 0.1); W moves along the view direction
    cam = host.default_camera()
    host.freecam_update(cam, 0.016, up=True)
    assert cam.moveSpeed == pytest.approx(0.0005 * 0.016)
    assert cam.position[1] == pytest.approx(-0.001 + 0.0005 * 0.016, rel=1e-5)
    cam = host.default_camera()
    host.freecam_update(cam, 5.0, floating=True, moving=True, cursor=(256.0, 0.0))
    assert cam.moveSpeed == pytest.approx(0.0005 * 0.1)
    assert cam.position[2] == pytest.approx(0.1 + 0.0005 * 0.1)
    assert cam.yaw == pytest.approx(host.deg(90) - 1.0)


def test_atmosphere_earth_matches_oracle(oracle):
    assert bytes(host.atmosphere_earth()) == bytes(oracle.earth())


def test_renderer_fails_loudly_without_gpu(blue_noise):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MinoteError, match="no CPU fallback"):
        host.Renderer(64, 64, blue_noise)


def test_frames_in_flight_range_is_checked(blue_noise):
    """Cuda::Provider(device, framesInFlight): 1..3 (the reference's InflightFrames, renderer.ixx:36); anything else is
    a logic error raised before any device is touched."""
    for n in (0, 4, -1):
        with pytest.raises(capi.MinoteError, match="frames in flight"):
            host.Renderer(64, 64, blue_noise, frames_in_flight=n)
