// minote_headless -- the App::run loop (src/app.ixx:19-45) without a window: builds the services,
// sets the reference's initial camera, runs Freecam + Renderer::draw for N frames and writes the
// last framebuffer as a PPM.  Like the reference (InflightFrames = 3, src/gfx/renderer.ixx:36) it keeps up to 3
// frames in flight: draw() rotates the frame contexts and the framebuffer of frame i is copied out while frames
// i+1, i+2 render.  Usage: minote_headless <blue_noise.rgba8> [frames] [width height] [out.ppm] [frames in flight]
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

int main(int argc, char** argv) try {
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
        std::printf("Frame time: %.2f ms\n", frameTime * 1000.0f);  // src/gfx/renderer.ixx:124
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
