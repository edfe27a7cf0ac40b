// minote.renderer -- frame orchestration, Renderer_impl::draw(camera) (src/gfx/renderer.ixx:39-68):
// sky LUTs -> primaryRays -> secondaryRays -> denoise -> tonemap, same order and defaults (bilateral
// denoiser renderer.ixx:140-141, AMD tonemapper, exposure 1, renderer.ixx:183-184).  Presentation (swapchain blit) is
// replaced by an RGBA8 framebuffer the caller reads back.
module;
#include <cstdint>
#include <cstring>
#include <ctime>

#include "../../include/minotert.h"
export module minote.renderer;
import minote.math;
import minote.camera;
import minote.cuda;
import minote.modules.sky;
import minote.modules.pathtracer;
import minote.modules.tonemapper;
import minote.modules.denoiser;
import minote.modules.reprojector;

export enum class DenoiseMode : int { None = 0, Bilateral = 1 };  // renderer.ixx:129-132
export enum class TonemapMode : int { Linear = 0, Reinhard = 1, Hable = 2, ACES = 3, Uchimura = 4, AMD = 5 };

export class Renderer_impl : Cuda {
public:
    // blue noise: decoded RGBA8 pixels of assets/blue_noise.png (renderer.ixx:101-110)
    Renderer_impl(uvec2 outputSize, std::uint8_t const* blueNoiseRgba8, uvec2 blueNoiseSize) : outputSize(outputSize) {
        forEachFrame([&](mrt_context* c) { return mrt_upload_blue_noise(c, blueNoiseRgba8, blueNoiseSize.x(), blueNoiseSize.y()); });
        blueNoise.id = 0;
    }
    ~Renderer_impl() {
        for (auto* a : atmosphere) delete a;
    }
    Renderer_impl(Renderer_impl const&) = delete;
    auto operator=(Renderer_impl const&) -> Renderer_impl& = delete;

    // scene selection (the reference compiles its scene into the shader, src/gpu/scene.glsl)
    // and keeps one copy of it for all frames in flight: frame context 0 owns triangles + BVH, the others borrow them
    void setSpheres(mrt_sphere const* spheres, u32 n) {
        forEachFrame([&](mrt_context* c) { return mrt_scene_set_spheres(c, spheres, n); });
    }
    void setMesh(float const* positions, u32 nverts, std::uint32_t const* indices, u32 ntris, float const* albedo) {
        auto* const owner = Cuda::serv->owner();
        waitBorrowers();
        Cuda::serv->checkOn(owner, mrt_scene_upload_mesh(owner, positions, nverts, indices, ntris, albedo));
        Cuda::serv->checkOn(owner, mrt_scene_build(owner, MRT_BUILD_FULL));
        shareScene();
    }
    void updateMesh(float const* positions, u32 nverts, bool refit) {
        auto* const owner = Cuda::serv->owner();
        if (asyncUpdate) {
            // option async_update: the library orders upload, refit / rebuild and the frames in flight on the GPU
            // (events); the borrowed scenes stay valid.  A refit returns at once, a rebuild when its own stream is done --
            // the frames already recorded keep rendering meanwhile (positions: page-locked, left alone until then)
            Cuda::serv->checkOn(owner, mrt_scene_update_positions(owner, positions, nverts));
            Cuda::serv->checkOn(owner, mrt_scene_build(owner, refit ? MRT_BUILD_REFIT : MRT_BUILD_FULL));
            return;
        }
        waitBorrowers();  // frames in flight still read the nodes a refit rewrites in place
        Cuda::serv->checkOn(owner, mrt_scene_update_positions(owner, positions, nverts));
        Cuda::serv->checkOn(owner, mrt_scene_build(owner, refit ? MRT_BUILD_REFIT : MRT_BUILD_FULL));
        shareScene();
    }

    void draw(Camera const& camera) {
        updateFrameTime();
        // Begin the frame: the next frame context (a progressive accumulator or a temporal history lives in one
        // context, so with either the frame stays put)
        Cuda::serv->nextFrame(!pathtracer.accumulate && !temporal);
        // Initial temporal resource values
        if (Cuda::serv->frameCount() == 1) prevCamera = camera;

        // transmittance / multi-scattering depend only on the (constant) parameters: the reference
        // rebuilds them every frame (renderer.ixx:56), here once per parameter set and frame context
        auto*& atmo = atmosphere[Cuda::serv->frameSlot()];
        if (!atmo) atmo = new Atmosphere(Atmosphere::Params::earth());
        auto sky = Sky();
        auto skyView = sky.createView(*atmo, camera.position);
        if (pathtracer.aerialPerspective) {  // camera-dependent: rebuilt with the view (32^3 froxels, ~20 us)
            auto const constants = Pathtracer::primaryConstants(camera, prevCamera, Cuda::serv->frameCount());
            mrt_primary_constants raw;
            std::memcpy(&raw, &constants, sizeof raw);
            sky.createAerialPerspective(*atmo, raw, camera.position);
        }
        auto gbuffer = pathtracer.primaryRays(outputSize, camera, prevCamera);
        auto pathtraced = pathtracer.secondaryRays(gbuffer, camera, *atmo, skyView, blueNoise);
        // temporal accumulation (off by default: the reference has none) takes the denoiser's place in the chain
        auto filtered = temporal ? reprojector.accumulate(pathtraced, gbuffer.visibility, gbuffer.motion, reprojectorParams)
                                 : denoise(pathtraced, gbuffer.depth, gbuffer.normal, camera);
        framebuffer = tonemap(filtered);

        // Temporal preservation
        prevCamera = camera;
    }

    // moving-average frame time in seconds, refreshed every 0.25 s (renderer.ixx:37,70-71,113-125); the reference shows it
    // in its ImGui overlay, here the caller prints it
    [[nodiscard]] auto frameTime() const -> float { return m_frameTime; }

    // host copy of the output framebuffer (RGBA8); blocks until the frame is done
    void readFramebuffer(void* host, std::size_t bytes) const { framebuffer.readback(host, bytes); }
    // frames in flight (renderer.ixx:36): start copying this frame out while the next draw() is issued
    void readFramebufferAsync(void* host, std::size_t bytes) const { framebuffer.readbackAsync(host, bytes); }
    // returns once all but the `framesInFlight` most recent frames have landed in host memory
    void waitFramebuffer(int framesInFlight = 0) const {
        int const n = Cuda::serv->framesInFlight(), cur = Cuda::serv->frameSlot();
        for (int back = 0; back < n; back++) {  // back = 0: the context of the newest frame
            auto* const c = Cuda::serv->frameContext((cur - back + n) % n);
            Cuda::serv->checkOn(c, mrt_readback_wait(c, back < framesInFlight ? 1 : 0));
        }
    }
    // stats of the current frame context; the running totals (rays, kernel launches) summed over all frames in flight
    [[nodiscard]] auto stats() const -> mrt_stats {
        mrt_stats s;
        Cuda::serv->check(mrt_stats_get(Cuda::serv->ctx, &s));
        for (int i = 0; i < Cuda::serv->framesInFlight(); i++) {
            auto* const c = Cuda::serv->frameContext(i);
            if (c == Cuda::serv->ctx) continue;
            mrt_stats o;
            Cuda::serv->checkOn(c, mrt_stats_get(c, &o));
            s.total_rays += o.total_rays;
            s.kernel_launches += o.kernel_launches;
        }
        return s;
    }
    void resetStats() const {
        forEachFrame([](mrt_context* c) { return mrt_stats_reset(c); });
    }
    void setOption(char const* name, std::int64_t value) {
        forEachFrame([&](mrt_context* c) { return mrt_set_option(c, name, value); });
        if (std::strcmp(name, "async_update") == 0) asyncUpdate = value != 0;
    }

    uvec2 outputSize;
    Pathtracer pathtracer;
    Tonemapper tonemapper;
    Denoiser denoiser;
    Reprojector reprojector;
    bool asyncUpdate = false;  // updateMesh(refit) without host stalls (mrt_set_option "async_update")
    bool temporal = false;  // reproject + accumulate along GBuffer::motion instead of denoising
    ReprojectorParams reprojectorParams = ReprojectorParams::make_default();
    // ImGui statics of Renderer_impl::denoise (renderer.ixx:140-141)
    DenoiseMode denoiseMode = DenoiseMode::Bilateral;
    BilateralParams bilateralParams = BilateralParams::make_default();
    // ImGui statics of Renderer_impl::tonemap (renderer.ixx:183-187)
    float exposure = 1.0f;
    TonemapMode tonemapMode = TonemapMode::AMD;
    float reinhardMax = 8.0f;
    UchimuraParams uchimuraParams = UchimuraParams::make_default();
    AMDParams amdParams = AMDParams::make_default();
    DeviceImage framebuffer;

private:
    template <typename F>
    void forEachFrame(F&& call) const {
        for (int i = 0; i < Cuda::serv->framesInFlight(); i++) {
            auto* const c = Cuda::serv->frameContext(i);
            Cuda::serv->checkOn(c, call(c));
        }
    }
    void waitBorrowers() const {
        for (int i = 1; i < Cuda::serv->framesInFlight(); i++) Cuda::serv->checkOn(Cuda::serv->frameContext(i), mrt_sync(Cuda::serv->frameContext(i)));
    }
    void shareScene() const {
        for (int i = 1; i < Cuda::serv->framesInFlight(); i++)
            Cuda::serv->checkOn(Cuda::serv->frameContext(i), mrt_scene_share(Cuda::serv->frameContext(i), Cuda::serv->owner()));
    }

    // renderer.ixx:127-161
    auto denoise(DeviceImage color, DeviceImage depth, DeviceImage normal, Camera const& camera) -> DeviceImage {
        switch (denoiseMode) {
        case DenoiseMode::None: return color;
        case DenoiseMode::Bilateral: return denoiser.bilateral(color, depth, normal, camera, bilateralParams);
        default: Cuda::serv->raise("Unknown denoise mode");
        }
    }

    auto tonemap(DeviceImage color) -> DeviceImage {
        switch (tonemapMode) {
        case TonemapMode::Linear: return tonemapper.linear(color, exposure);
        case TonemapMode::Reinhard: return tonemapper.reinhard(color, exposure, reinhardMax);
        case TonemapMode::Hable: return tonemapper.hable(color, exposure);
        case TonemapMode::ACES: return tonemapper.aces(color, exposure);
        case TonemapMode::Uchimura: return tonemapper.uchimura(color, exposure, uchimuraParams);
        case TonemapMode::AMD: return tonemapper.amd(color, exposure, amdParams);
        default: Cuda::serv->raise("Unknown tonemap mode");
        }
    }

    // renderer.ixx:113-125 (Window::getTime -> the monotonic clock)
    static constexpr double FrameTimeUpdate = 0.25;
    static auto now() -> double {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
    }
    void updateFrameTime() {
        framesSinceLastCheck += 1;
        auto const currentTime = now();
        auto const timeElapsed = currentTime - lastFrameTimeCheck;
        if (timeElapsed >= FrameTimeUpdate) {
            m_frameTime = float(timeElapsed / double(framesSinceLastCheck));
            lastFrameTimeCheck = currentTime;
            framesSinceLastCheck = 0;
        }
    }
    float m_frameTime = 0.0f;
    double lastFrameTimeCheck = now();
    unsigned framesSinceLastCheck = 0;

    Camera prevCamera{};
    DeviceImage blueNoise;
    Atmosphere* atmosphere[Cuda_impl::MaxFramesInFlight] = {nullptr, nullptr, nullptr};  // one per frame context
};

export class Renderer {
public:
    class Provider {
    public:
        Provider(uvec2 outputSize, std::uint8_t const* blueNoiseRgba8, uvec2 blueNoiseSize)
            : inst(new Renderer_impl(outputSize, blueNoiseRgba8, blueNoiseSize)), prev(serv) {
            serv = inst;
        }
        ~Provider() {
            serv = prev;
            delete inst;
        }
        Provider(Provider const&) = delete;
        auto operator=(Provider const&) -> Provider& = delete;

    private:
        Renderer_impl* inst;
        Renderer_impl* prev;
    };
    static inline Renderer_impl* serv = nullptr;
};
