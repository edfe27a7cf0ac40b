// minote.freecam -- Freecam::updateCamera (src/freecam.ixx:51-68) driven by explicit input state
// instead of GLFW callbacks (windowing is out of scope); same movement arithmetic.
module;
#include <algorithm>
export module minote.freecam;
import minote.math;
import minote.camera;

export class Freecam {
public:
    bool up = false, down = false, left = false, right = false, floating = false, moving = false;
    vec2 offset = {0.0f, 0.0f};  // cursor motion accumulated since the last update

    void cursorMoved(vec2 delta) { offset = offset + delta; }

    // frameTime: Renderer::frameTime() in seconds
    void updateCamera(Camera& camera, float frameTime) {
        auto const framerateScale = std::min(frameTime, 0.1f);
        camera.moveSpeed = 0.0005f * framerateScale;
        offset.y() *= -1.0f;  // Y points down in window coords but up in the world
        if (moving) camera.rotate(offset.x(), offset.y());
        offset = vec2{0.0f, 0.0f};
        camera.roam({float(right) - float(left), 0.0f, float(up) - float(down)});
        camera.shift({0.0f, 0.0f, float(floating)});
    }
};
