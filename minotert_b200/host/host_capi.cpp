// host_capi.cpp -- extern "C" shims over the host modules, so tools that cannot import C++20 modules
// (the Python tests and bench.py via ctypes) drive the SAME host code: Camera math, Renderer::draw,
// Freecam.  Exceptions are caught at this seam and turned into status codes + a message.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>

#include "../../include/minotert.h"
import minote.math;
import minote.camera;
import minote.cuda;
import minote.modules.sky;
import minote.modules.pathtracer;
import minote.modules.tonemapper;
import minote.modules.denoiser;
import minote.modules.reprojector;
import minote.renderer;
import minote.freecam;

namespace {
// plain pointers and a char buffer: this TU mixes textual standard headers with module imports, which
// g++ 13 only handles reliably for trivial types
struct App {
    Cuda::Provider* cuda = nullptr;
    Renderer::Provider* renderer = nullptr;
    char error[512] = {0};
};
char g_error[512] = {0};  // creation errors (single host thread, like the reference)

template <typename F>
int guarded(App* app, F&& f) {
    try {
        // one App at a time may be "current": the service pointers are process-wide like the reference's
        f();
        return 0;
    } catch (std::exception const& e) {
        std::snprintf(app ? app->error : g_error, 512, "%s", e.what());
        return -1;
    }
}
}  // namespace

extern "C" {

// ---- camera / constants (pure host math, usable without a GPU) ----
void minote_camera_constants(Camera const* cam, Camera const* prev, std::uint32_t frame, mrt_primary_constants* p,
                             mrt_secondary_constants* s) {
    if (p) {
        auto c = Pathtracer::primaryConstants(*cam, prev ? *prev : *cam, frame);
        std::memcpy(p, &c, sizeof *p);
    }
    if (s) {
        auto c = Pathtracer::secondaryConstants(*cam, frame);
        std::memcpy(s, &c, sizeof *s);
    }
}
void minote_camera_direction(Camera const* cam, float out[3]) {
    vec3 d = cam->direction();
    out[0] = d.x(); out[1] = d.y(); out[2] = d.z();
}
void minote_camera_rotate(Camera* cam, float horz, float vert) { cam->rotate(horz, vert); }
void minote_camera_shift(Camera* cam, float const d[3]) { cam->shift({d[0], d[1], d[2]}); }
void minote_camera_roam(Camera* cam, float const d[3]) { cam->roam({d[0], d[1], d[2]}); }
// the reference's initial camera (src/app.ixx:20-32)
void minote_camera_default(Camera* cam, std::uint32_t w, std::uint32_t h) {
    *cam = Camera{};
    cam->viewport = {w, h};
    cam->verticalFov = 60_deg;
    cam->nearPlane = 0.001f;
    cam->position = {0.0f, -0.001f, 0.1f};
    cam->yaw = 90_deg;
    cam->pitch = 0.0f;
    cam->lookSpeed = 1.0f / 256.0f;
    cam->moveSpeed = 8.0f;
}
float minote_deg(double d) { return deg(d); }
void minote_atmosphere_earth(mrt_atmosphere_params* out) {
    auto p = Atmosphere::Params::earth();
    std::memcpy(out, &p, sizeof *out);
}
// keys: bit0 up, bit1 down, bit2 left, bit3 right, bit4 floating, bit5 moving (mouse button held)
void minote_freecam_update(Camera* cam, std::uint32_t keys, float cursor_dx, float cursor_dy, float frame_time) {
    Freecam f;
    f.up = keys & 1u; f.down = keys & 2u; f.left = keys & 4u; f.right = keys & 8u;
    f.floating = keys & 16u; f.moving = keys & 32u;
    f.cursorMoved({cursor_dx, cursor_dy});
    f.updateCamera(*cam, frame_time);
}

// ---- renderer ----
// frames_in_flight: 1..3 frame contexts that draw() rotates through (the reference: 3, renderer.ixx:36)
void* minote_app_create_in_flight(int device, int frames_in_flight, std::uint32_t w, std::uint32_t h, std::uint8_t const* blue_noise,
                                  std::uint32_t bn_w, std::uint32_t bn_h) {
    App* app = new App();
    int s = guarded(nullptr, [&] { app->cuda = new Cuda::Provider(device, frames_in_flight); });
    if (s == 0) s = guarded(nullptr, [&] { app->renderer = new Renderer::Provider(uvec2{w, h}, blue_noise, uvec2{bn_w, bn_h}); });
    if (s != 0) {
        delete app->cuda;
        delete app;
        return nullptr;
    }
    return app;
}
void* minote_app_create(int device, std::uint32_t w, std::uint32_t h, std::uint8_t const* blue_noise, std::uint32_t bn_w,
                        std::uint32_t bn_h) {
    return minote_app_create_in_flight(device, 1, w, h, blue_noise, bn_w, bn_h);
}
void minote_app_destroy(void* a) {
    auto* app = static_cast<App*>(a);
    if (!app) return;
    delete app->renderer;
    delete app->cuda;
    delete app;
}
char const* minote_app_error(void* a) { return a ? static_cast<App*>(a)->error : g_error; }
// frame context 0: owns the scene; with one frame in flight it is also the context of every frame
mrt_context* minote_app_context(void* a) { (void)a; return Cuda::serv ? Cuda::serv->owner() : nullptr; }
// the context the last draw() recorded into (index < 0) or frame context `index`
mrt_context* minote_app_frame_context(void* a, int index) {
    (void)a;
    if (!Cuda::serv) return nullptr;
    if (index < 0) return Cuda::serv->ctx;
    return index < Cuda::serv->framesInFlight() ? Cuda::serv->frameContext(index) : nullptr;
}
int minote_app_frames_in_flight(void* a) { (void)a; return Cuda::serv ? Cuda::serv->framesInFlight() : 0; }
// mrt_set_option on every frame context
int minote_app_set_option(void* a, char const* name, std::int64_t value) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->setOption(name, value); });
}
int minote_app_stats_reset(void* a) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->resetStats(); });
}

int minote_app_set_spheres(void* a, mrt_sphere const* s, std::uint32_t n) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->setSpheres(s, n); });
}
int minote_app_set_mesh(void* a, float const* pos, std::uint32_t nverts, std::uint32_t const* idx, std::uint32_t ntris,
                        float const* albedo) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->setMesh(pos, nverts, idx, ntris, albedo); });
}
int minote_app_update_mesh(void* a, float const* pos, std::uint32_t nverts, int refit) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->updateMesh(pos, nverts, refit != 0); });
}
int minote_app_configure(void* a, std::uint32_t samples, std::uint32_t bounces, int accumulate, int tonemap_mode, float exposure) {
    return guarded(static_cast<App*>(a), [&] {
        auto& r = *Renderer::serv;
        r.pathtracer.samples = samples;
        r.pathtracer.bounces = bounces;
        r.pathtracer.accumulate = accumulate != 0;
        r.tonemapMode = static_cast<TonemapMode>(tonemap_mode);
        r.exposure = exposure;
    });
}
// sky extensions of the path tracer: sun as a sampled light, sky evaluated at the shaded point
int minote_app_set_sky_extensions(void* a, int sun_sampling, int sky_at_hit, int aerial_perspective) {
    return guarded(static_cast<App*>(a), [&] {
        Renderer::serv->pathtracer.sunSampling = sun_sampling != 0;
        Renderer::serv->pathtracer.skyAtHit = sky_at_hit != 0;
        Renderer::serv->pathtracer.aerialPerspective = aerial_perspective != 0;
    });
}
// Renderer_impl::denoise controls (renderer.ixx:140-152): mode 0 None, 1 Bilateral
int minote_app_set_denoise(void* a, int mode, float sigma, float kSigma, float threshold) {
    return guarded(static_cast<App*>(a), [&] {
        auto& r = *Renderer::serv;
        r.denoiseMode = static_cast<DenoiseMode>(mode);
        r.bilateralParams = BilateralParams{sigma, kSigma, threshold};
    });
}
// temporal accumulation along GBuffer::motion (takes the denoiser's place in draw()); max_history <= 0 keeps the default
int minote_app_set_temporal(void* a, int enabled, float max_history) {
    return guarded(static_cast<App*>(a), [&] {
        auto& r = *Renderer::serv;
        r.temporal = enabled != 0;
        if (max_history > 0.0f) r.reprojectorParams.maxHistory = max_history;
    });
}
int minote_app_resize(void* a, std::uint32_t w, std::uint32_t h) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->outputSize = {w, h}; });
}
// Renderer::serv->draw(camera)  (src/app.ixx:39)
int minote_app_draw(void* a, Camera const* cam) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->draw(*cam); });
}
int minote_app_read_framebuffer(void* a, void* host, std::size_t bytes) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->readFramebuffer(host, bytes); });
}
int minote_app_read_framebuffer_async(void* a, void* host, std::size_t bytes) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->readFramebufferAsync(host, bytes); });
}
int minote_app_wait_framebuffer(void* a, int frames_in_flight) {
    return guarded(static_cast<App*>(a), [&] { Renderer::serv->waitFramebuffer(frames_in_flight); });
}
int minote_app_stats(void* a, mrt_stats* out) {
    return guarded(static_cast<App*>(a), [&] { *out = Renderer::serv->stats(); });
}
float minote_app_frame_time(void* a) { (void)a; return Renderer::serv ? Renderer::serv->frameTime() : 0.0f; }
std::uint32_t minote_app_frame_count(void* a) { (void)a; return Cuda::serv ? Cuda::serv->frameCount() : 0; }

}  // extern "C"
