// minote_headless -- the App::run loop (src/app.ixx:19-45) without a window: builds the services,
// sets the reference's initial camera, runs Freecam + Renderer::draw for N frames and writes the
// last framebuffer as a PPM.  Like the reference (InflightFrames = 3, src/gfx/renderer.ixx:36) it keeps up to 3
// frames in flight: draw() rotates the frame contexts and the framebuffer of frame i is copied out while frames
// i+1, i+2 render.  Usage: minote_headless <blue_noise.rgba8> [frames] [width height] [out.ppm] [frames in flight]
// With `--script events.txt` as the first arguments the frames are driven by the scripted front end (minote.frontend: the
// events of the reference's window + ImGui panels):  minote_headless --script events.txt <blue_noise.rgba8> [w h]
// With `--gpus N` as the first arguments the same frames are rendered tile-partitioned over N GPUs through the
// C ABI's mrt_group_* (replicated scene, NCCL gather of the RGBA8 framebuffer every frame):
//        minote_headless --gpus N <blue_noise.rgba8> [frames] [width height] [out.ppm]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <exception>

#include "../../include/minotert.h"
import minote.math;
import minote.camera;
import minote.cuda;
import minote.renderer;
import minote.freecam;
import minote.modules.pathtracer;
import minote.modules.sky;
import minote.frontend;

// Tile-partitioned multi-GPU loop: one process, N devices, everything behind the C ABI.
static int run_group(int ngpus, std::uint8_t const* bn, int frames, u32 w, u32 h, char const* out) {
    int devices[64];
    for (int i = 0; i < ngpus; i++) devices[i] = i;
    mrt_group* g = nullptr;
    auto fail = [&](char const* what) {
        std::fprintf(stderr, "%s: %s\n", what, mrt_group_last_error(g));
        if (g) mrt_group_destroy(g);
        return EXIT_FAILURE;
    };
    if (mrt_group_create(devices, u32(ngpus), MRT_GROUP_NCCL, &g) != MRT_OK) return fail("mrt_group_create");
    mrt_sphere const spheres[5] = {{{0.0000f, 0.0017f, 0.10000f}, 0.00050f, {0.2f, 0.7f, 0.0f}},
                                   {{-0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.0f, 0.2f, 0.7f}},
                                   {{0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.7f, 0.0f, 0.2f}},
                                   {{0.0000f, 0.0008f, 0.09975f}, 0.00025f, {1.0f, 1.0f, 1.0f}},
                                   {{0.0000f, 0.0010f, -0.00050f}, 0.10000f, {0.5f, 0.5f, 0.5f}}};
    auto camera = Camera{{w, h}, 60_deg, 0.001f, {0.0f, -0.001f, 0.1f}, 90_deg, 0.0f, 1.0f / 256.0f, 8.0f};
    auto const earth = Atmosphere::Params::earth();
    float const sunDir[3] = {-0.435286462f, 0.818654716f, 0.374606609f}, sunIll[3] = {8.0f, 8.0f, 8.0f};  // sky.ixx:193-194
    float const amd[5] = {16.0f, 2.0f, 1.0f, 0.18f, 0.18f};                                                // tonemapper.ixx:46-54
    for (u32 i = 0; i < u32(ngpus); i++) {
        mrt_context* c = nullptr;
        mrt_group_context(g, i, &c, nullptr);
        float const probe[3] = {camera.position.x(), camera.position.y(), camera.position.z()};
        if (mrt_upload_blue_noise(c, bn, 256, 256) != MRT_OK || mrt_scene_set_spheres(c, spheres, 5) != MRT_OK ||
            mrt_atmosphere(c, reinterpret_cast<mrt_atmosphere_params const*>(&earth)) != MRT_OK ||
            mrt_sky_view(c, probe, sunDir, sunIll) != MRT_OK) {
            std::fprintf(stderr, "rank %u setup: %s\n", i, mrt_last_error(c));
            mrt_group_destroy(g);
            return EXIT_FAILURE;
        }
    }
    if (mrt_group_set_tiles(g, 8) != MRT_OK) return fail("mrt_group_set_tiles");
    std::size_t const fbBytes = std::size_t(w) * h * 4;
    auto* fb = static_cast<std::uint8_t*>(std::malloc(fbBytes));
    for (int i = 0; i < frames; i++) {
        auto const pc = Pathtracer::primaryConstants(camera, camera, u32(i + 1));
        auto const sc = Pathtracer::secondaryConstants(camera, u32(i + 1));
        if (mrt_group_render(g, w, h, reinterpret_cast<mrt_primary_constants const*>(&pc),
                             reinterpret_cast<mrt_secondary_constants const*>(&sc), 8, 8, 0, 0) != MRT_OK) return fail("mrt_group_render");
        if (mrt_group_tonemap(g, MRT_TONEMAP_AMD, 1.0f, amd, 5, MRT_BUF_COLOR) != MRT_OK) return fail("mrt_group_tonemap");
        if (mrt_group_gather(g, MRT_BUF_LDR, 0) != MRT_OK) return fail("mrt_group_gather");
        if (mrt_group_readback(g, fb, fbBytes) != MRT_OK) return fail("mrt_group_readback");
    }
    if (FILE* f = std::fopen(out, "wb")) {
        std::fprintf(f, "P6\n%u %u\n255\n", w, h);
        for (std::size_t i = 0; i < std::size_t(w) * h; i++) std::fwrite(&fb[4 * i], 1, 3, f);
        std::fclose(f);
    }
    std::free(fb);
    mrt_group_destroy(g);
    std::printf("rendered %d frame(s) tile-partitioned over %d GPU(s)\n", frames, ngpus);
    return EXIT_SUCCESS;
}

int main(int argc, char** argv) try {
    int ngpus = 0;
    if (argc > 2 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 'g') {  // --gpus N
        ngpus = std::atoi(argv[2]);
        argc -= 2;
        argv += 2;
    }
    char const* script = nullptr;
    if (argc > 2 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 's') {  // --script file
        script = argv[2];
        argc -= 2;
        argv += 2;
    }
    if (script) {
        if (argc < 2) {
            std::fprintf(stderr, "usage: minote_headless --script events.txt <blue_noise.rgba8> [width height]\n");
            return EXIT_FAILURE;
        }
        u32 const w = argc > 3 ? std::atoi(argv[2]) : 960, h = argc > 3 ? std::atoi(argv[3]) : 540;
        std::size_t const bnBytes = 256 * 256 * 4;
        auto* bn = static_cast<std::uint8_t*>(std::malloc(bnBytes));
        FILE* bf = std::fopen(argv[1], "rb");
        if (!bf || std::fread(bn, 1, bnBytes, bf) != bnBytes) {
            std::fprintf(stderr, "cannot read 256x256 RGBA8 blue noise from %s\n", argv[1]);
            return EXIT_FAILURE;
        }
        std::fclose(bf);
        Cuda::Provider cuda(0, 1);  // every frame is shown before the next one is drawn: one frame in flight
        Renderer::Provider renderer(uvec2{w, h}, bn, uvec2{256u, 256u});
        mrt_sphere const spheres[5] = {{{0.0000f, 0.0017f, 0.10000f}, 0.00050f, {0.2f, 0.7f, 0.0f}},
                                       {{-0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.0f, 0.2f, 0.7f}},
                                       {{0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.7f, 0.0f, 0.2f}},
                                       {{0.0000f, 0.0008f, 0.09975f}, 0.00025f, {1.0f, 1.0f, 1.0f}},
                                       {{0.0000f, 0.0010f, -0.00050f}, 0.10000f, {0.5f, 0.5f, 0.5f}}};
        Renderer::serv->setSpheres(spheres, 5);
        auto camera = Camera{{w, h}, 60_deg, 0.001f, {0.0f, -0.001f, 0.1f}, 90_deg, 0.0f, 1.0f / 256.0f, 8.0f};  // src/app.ixx:20-32
        Frontend ui(script);
        std::size_t const fbBytes = std::size_t(w) * h * 4;
        auto* fb = static_cast<std::uint8_t*>(std::malloc(fbBytes));
        float frameTime = 1.0f / 60.0f;
        auto now = [] {
            timespec ts;
            clock_gettime(CLOCK_MONOTONIC, &ts);
            return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
        };
        for (int i = 0;; i++) {
            double const t0 = now();
            bool const more = ui.beginFrame(i, *Renderer::serv);          // Window::poll + ImGui widgets
            ui.freecam.updateCamera(camera, ui.fixedFrameTime > 0.0f ? ui.fixedFrameTime : frameTime);
            Renderer::serv->draw(camera);
            if (ui.wantsPresent()) {
                Renderer::serv->readFramebuffer(fb, fbBytes);
                ui.presentAll(fb, w, h);
            }
            frameTime = float(now() - t0);
            std::printf("frame %d  Frame time: %.2f ms  camera %.8f %.8f %.8f yaw %.6f pitch %.6f\n", i, frameTime * 1000.0f,
                        camera.position.x(), camera.position.y(), camera.position.z(), camera.yaw, camera.pitch);
            if (!more) break;
        }
        std::free(fb);
        std::free(bn);
        return EXIT_SUCCESS;
    }
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <blue_noise.rgba8 (256x256 raw)> [frames] [width height] [out.ppm] [frames in flight 1..3]\n", argv[0]);
        return EXIT_FAILURE;
    }
    int const frames = argc > 2 ? std::atoi(argv[2]) : 8;
    u32 const w = argc > 4 ? std::atoi(argv[3]) : 960, h = argc > 4 ? std::atoi(argv[4]) : 540;  // src/main.cpp:24
    char const* out = argc > 5 ? argv[5] : "minote.ppm";
    int const inFlight = argc > 6 ? std::atoi(argv[6]) : Cuda_impl::MaxFramesInFlight;
    // plain buffers: this TU mixes textual standard headers with module imports (see host_capi.cpp)
    std::size_t const bnBytes = 256 * 256 * 4;
    auto* bn = static_cast<std::uint8_t*>(std::malloc(bnBytes));
    FILE* bf = std::fopen(argv[1], "rb");
    if (!bf || std::fread(bn, 1, bnBytes, bf) != bnBytes) {
        std::fprintf(stderr, "cannot read 256x256 RGBA8 blue noise from %s\n", argv[1]);
        return EXIT_FAILURE;
    }
    std::fclose(bf);
    if (ngpus > 0) {
        int const rc = run_group(ngpus, bn, frames, w, h, out);
        std::free(bn);
        return rc;
    }
    Cuda::Provider cuda(0, inFlight);
    Renderer::Provider renderer(uvec2{w, h}, bn, uvec2{256u, 256u});
    // the reference's compiled-in scene (src/gpu/scene.glsl:5-11)
    mrt_sphere const spheres[5] = {{{0.0000f, 0.0017f, 0.10000f}, 0.00050f, {0.2f, 0.7f, 0.0f}},
                                   {{-0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.0f, 0.2f, 0.7f}},
                                   {{0.0008f, 0.0012f, 0.09983f}, 0.00033f, {0.7f, 0.0f, 0.2f}},
                                   {{0.0000f, 0.0008f, 0.09975f}, 0.00025f, {1.0f, 1.0f, 1.0f}},
                                   {{0.0000f, 0.0010f, -0.00050f}, 0.10000f, {0.5f, 0.5f, 0.5f}}};
    Renderer::serv->setSpheres(spheres, 5);
    auto camera = Camera{{w, h}, 60_deg, 0.001f, {0.0f, -0.001f, 0.1f}, 90_deg, 0.0f, 1.0f / 256.0f, 8.0f};  // src/app.ixx:20-32
    auto freecam = Freecam();
    float frameTime = 1.0f / 60.0f;
    std::size_t const fbBytes = std::size_t(w) * h * 4;
    // one host buffer per frame that may still be copying, plus the one being filled next
    std::uint8_t* ring[Cuda_impl::MaxFramesInFlight + 1];
    for (int k = 0; k <= inFlight; k++) ring[k] = static_cast<std::uint8_t*>(std::malloc(fbBytes));
    std::uint8_t* fb = ring[0];
    auto now = [] {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
    };
    for (int i = 0; i < frames; i++) {
        double const t0 = now();
        freecam.updateCamera(camera, frameTime);
        Renderer::serv->draw(camera);
        fb = ring[i % (inFlight + 1)];
        Renderer::serv->readFramebufferAsync(fb, fbBytes);
        Renderer::serv->waitFramebuffer(inFlight - 1);  // all but the newest inFlight - 1 frames are in host memory
        frameTime = float(now() - t0);
        // the reference's overlay: moving average over 0.25 s (src/gfx/renderer.ixx:113-125); 0 until the first window closes
        std::printf("Frame time: %.2f ms\n", (Renderer::serv->frameTime() > 0.0f ? Renderer::serv->frameTime() : frameTime) * 1000.0f);
    }
    Renderer::serv->waitFramebuffer(0);
    if (FILE* f = std::fopen(out, "wb")) {
        std::fprintf(f, "P6\n%u %u\n255\n", w, h);
        for (std::size_t i = 0; i < std::size_t(w) * h; i++) std::fwrite(&fb[4 * i], 1, 3, f);
        std::fclose(f);
    }
    for (int k = 0; k <= inFlight; k++) std::free(ring[k]);
    std::free(bn);
    return EXIT_SUCCESS;
} catch (std::exception const& e) {
    std::fprintf(stderr, "Uncaught exception on main thread: %s\n", e.what());
    return EXIT_FAILURE;
}
