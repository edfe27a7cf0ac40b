// minote.camera -- the reference's Camera (src/gfx/camera.ixx:8-65), same members and methods.
module;
#include <cmath>
export module minote.camera;
import minote.math;

export class Camera {
public:
    // Projection
    uvec2 viewport;
    float verticalFov;
    float nearPlane;
    // View
    vec3 position;
    float yaw;
    float pitch;
    // Movement
    float lookSpeed;
    float moveSpeed;

    [[nodiscard]] auto direction() const -> vec3 {
        return vec3{std::cos(pitch) * std::cos(yaw), std::cos(pitch) * std::sin(yaw), std::sin(pitch)};
    }
    [[nodiscard]] auto view() const -> mat4 { return look(position, direction(), vec3{0.0f, 0.0f, 1.0f}); }
    // aspect is passed as height / width, as the reference does (camera.ixx:43)
    [[nodiscard]] auto projection() const -> mat4 {
        return perspective(verticalFov, float(viewport.y()) / float(viewport.x()), nearPlane);
    }
    void rotate(float horz, float vert) {
        yaw -= horz * lookSpeed;
        if (yaw < 0_deg) yaw += 360_deg;
        if (yaw >= 360_deg) yaw -= 360_deg;
        pitch += vert * lookSpeed;
        pitch = clamp(pitch, -89_deg, 89_deg);
    }
    void shift(vec3 distance) { position += distance * moveSpeed; }
    void roam(vec3 distance) {
        vec4 const d = inverse(view()) * vec4{distance.x(), distance.y(), distance.z(), 0.0f};
        shift(vec3{d.x(), d.y(), d.z()});
    }
};
static_assert(sizeof(Camera) == 44, "Camera is a POD shared with the C shim");
