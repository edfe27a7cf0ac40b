"""The scripted front end (SURVEY 8f-3 without a display; host/frontend.cppm + `minote_headless --script`): the events of the
reference's window (freecam.ixx:19-48) and the values of its ImGui widgets (renderer.ixx:127-223) come from a script, `present`
writes the framebuffer.  Checks: the key / mouse / cursor map moves the camera exactly like Freecam::updateCamera driven
directly; widget assignments reach the renderer (exposure, algorithm, denoiser); a presented frame is the frame the Renderer
draws through the Python bindings for the same camera and settings."""
import os
import re
import subprocess

import numpy as np
import pytest

from minotert_b200 import host

pytestmark = pytest.mark.gpu

SCRIPT = """
# frame event
0 dt 0.02                       # fixed frame time: reproducible camera path
0 set Denoiser/Algorithm None
0 set Pathtracer/Samples 2
0 set Pathtracer/Bounces 2
1 present f1.ppm
2 set Tonemapper/Exposure 4
2 present f2.ppm
3 set Tonemapper/Exposure 1
3 set Tonemapper/Algorithm Reinhard
3 set Tonemapper/HDR peak 4
3 present f3.ppm
4 set Tonemapper/Algorithm AMD
4 key W down
6 key W up
6 cursor 100 100                # first report: only remembered (Freecam() reads the initial position)
7 cursor 130 90                 # mouse not pressed: the camera must not turn
8 mouse down
9 cursor 150 95
10 mouse up
11 cursor 500 500
12 key SPACE down
13 key SPACE up
14 present f14.ppm
"""


def ppm(path, w, h):
    data = open(path, "rb").read()
    header = f"P6\n{w} {h}\n255\n".encode()
    assert data.startswith(header) and len(data) == len(header) + w * h * 3
    return np.frombuffer(data[len(header):], np.uint8).reshape(h, w, 3)


def test_scripted_front_end(oracle, blue_noise, tmp_path):
    exe = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "minote_headless")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    raw = tmp_path / "blue_noise.rgba8"
    raw.write_bytes(np.ascontiguousarray(blue_noise, np.uint8).tobytes())
    (tmp_path / "events.txt").write_text(SCRIPT)
    w, h = 240, 135
    p = subprocess.run([exe, "--script", "events.txt", str(raw), str(w), str(h)], capture_output=True, text=True, timeout=180, cwd=tmp_path)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == 15 and all("Frame time:" in l for l in lines)
    cams = np.array([[float(x) for x in re.search(r"camera (\S+) (\S+) (\S+) yaw (\S+) pitch (\S+)", l).groups()] for l in lines])
    pos, yaw, pitch = cams[:, :3], cams[:, 3], cams[:, 4]
    # W held on frames 4 and 5: two steps forward (the camera faces +y), then still
    assert np.allclose(pos[:4], pos[0]) and pos[4, 1] > pos[3, 1] and pos[5, 1] > pos[4, 1] and np.allclose(pos[6:12], pos[5])
    step = 0.0005 * 0.02 * 1.0   # Freecam: moveSpeed = 0.0005 * min(frameTime, 0.1); Camera::roam scales by it
    assert abs((pos[5, 1] - pos[3, 1]) - 2 * step) < 0.2 * step
    # cursor motion turns the camera only while the left button is down (frame 9), by the Camera's look speed
    assert np.allclose(yaw[:9], yaw[0]) and np.allclose(pitch[:9], pitch[0])
    assert yaw[9] != yaw[8] and pitch[9] != pitch[8]
    assert np.allclose(yaw[10:], yaw[9]) and np.allclose(pitch[10:], pitch[9])
    # SPACE on frame 12: up by one step
    assert pos[12, 2] > pos[11, 2] and np.allclose(pos[13:, 2], pos[12, 2])
    f1, f2, f3, f14 = (ppm(tmp_path / n, w, h) for n in ("f1.ppm", "f2.ppm", "f3.ppm", "f14.ppm"))
    assert f2.astype(int).sum() > 1.2 * f1.astype(int).sum()          # exposure 4
    assert not np.array_equal(f3, f1) and not np.array_equal(f14, f1)  # another operator; another view
    # frame 1 == Renderer::draw through the Python bindings: same settings, default camera, frame counter 2
    r = host.Renderer(w, h, blue_noise, frames_in_flight=1)
    try:
        r.set_spheres(oracle.REFERENCE_SPHERES)
        r.configure(samples=2, bounces=2, denoise="none")
        cam = host.default_camera(w, h)
        r.draw(cam)
        r.draw(cam)
        fb = r.read_framebuffer()
    finally:
        r.close()
    assert np.array_equal(fb[..., :3], f1)


def test_front_end_rejects_unknown_widgets_and_events(blue_noise, tmp_path):
    exe = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "minote_headless")
    raw = tmp_path / "blue_noise.rgba8"
    raw.write_bytes(np.ascontiguousarray(blue_noise, np.uint8).tobytes())
    for bad in ("0 set Tonemapper/Sharpness 3\n", "0 wiggle left\n", "0 set Tonemapper/Algorithm Filmic\n"):
        (tmp_path / "bad.txt").write_text(bad)
        p = subprocess.run([exe, "--script", "bad.txt", str(raw), "64", "36"], capture_output=True, text=True, timeout=120, cwd=tmp_path)
        assert p.returncode != 0 and "front" in p.stderr
