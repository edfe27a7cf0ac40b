// minote.frontend -- the reference's interactive front end (SURVEY 8f-3) without a display: the events a GLFW window would
// deliver and the values its ImGui widgets would hold come from a script, and "present" writes the framebuffer to a file.
// What it mirrors:
//   * Freecam::registerEvents (src/freecam.ixx:19-48): the same key -> direction map (W/Up, S/Down, A/Left, D/Right,
//     Space = float up), cursor motion accumulated into Freecam::offset, left mouse button = look around;
//   * the ImGui statics of Renderer_impl::tonemap / ::denoise (src/gfx/renderer.ixx:127-161,163-223): every widget is
//     addressed by its header and label ("Tonemapper/Exposure", "Tonemapper/Algorithm", "Tonemapper/HDR peak", ...,
//     "Denoiser/Algorithm", "Denoiser/Sigma", ...) and clamped to the slider's range;
//   * the "Frame time: x ms" overlay (src/gfx/renderer.ixx:113-125) -> status line;
//   * blitAndPresent (src/gfx/renderer.ixx:225-263) -> `present <file.ppm>`.
// A window (GLFW), ImGui and a swapchain do not exist in this environment (no display, no network to fetch them): the
// seam they plug into is exactly this class -- a GLFW callback would call key()/cursor()/mouseButton(), an ImGui panel
// would call set() with the same names.
// Script: one event per line, `#` comments:   <frame> key <W|S|A|D|UP|DOWN|LEFT|RIGHT|SPACE> <down|up>
//   <frame> cursor <x> <y>      (absolute window position, as GLFW reports it)
//   <frame> mouse <down|up>     (left button)
//   <frame> set <Header/Label> <value>      <frame> present <file.ppm>      <frame> dt <seconds> (fixed frame time; 0 = measured)
module;
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
export module minote.frontend;
import minote.math;
import minote.camera;
import minote.freecam;
import minote.renderer;

export class Frontend {
public:
    struct Event {
        int frame;
        std::string kind, a, b;
    };
    Freecam freecam;
    float fixedFrameTime = 0.0f;       // > 0: Freecam sees this frame time (reproducible camera paths)
    std::vector<std::string> presents;  // files to write this frame

    explicit Frontend(char const* scriptPath) {
        FILE* f = std::fopen(scriptPath, "r");
        if (!f) throw std::runtime_error(std::string("cannot open front-end script ") + scriptPath);
        char line[512];
        while (std::fgets(line, sizeof line, f)) {
            if (char* h = std::strchr(line, '#')) *h = 0;
            char kind[64] = "", a[256] = "", b[128] = "";
            int frame = 0;
            // "set" takes a label that may contain spaces: <frame> set <Header/Label words...> <value>
            int const n = std::sscanf(line, "%d %63s", &frame, kind);
            if (n < 2) continue;
            Event e{frame, kind, "", ""};
            char const* rest = std::strstr(line, kind) + std::strlen(kind);
            if (e.kind == "set") {
                std::string r = trim(rest);
                auto const cut = r.find_last_of(" \t");
                if (cut == std::string::npos) throw std::runtime_error("front-end script: set needs a label and a value: " + r);
                e.a = trim(r.substr(0, cut).c_str());
                e.b = r.substr(cut + 1);
            } else {
                std::sscanf(rest, "%255s %127s", a, b);
                e.a = a;
                e.b = b;
            }
            events.push_back(e);
        }
        std::fclose(f);
        std::stable_sort(events.begin(), events.end(), [](Event const& x, Event const& y) { return x.frame < y.frame; });
    }

    [[nodiscard]] auto lastFrame() const -> int { return events.empty() ? 0 : events.back().frame; }

    // ---- the callbacks GLFW would invoke (freecam.ixx:21-47)
    void key(std::string const& k, bool action) {
        if (k == "W" || k == "UP") freecam.up = action;
        else if (k == "S" || k == "DOWN") freecam.down = action;
        else if (k == "A" || k == "LEFT") freecam.left = action;
        else if (k == "D" || k == "RIGHT") freecam.right = action;
        else if (k == "SPACE") freecam.floating = action;
    }
    void cursor(vec2 newPos) {
        if (haveCursor) freecam.cursorMoved(newPos - prevCursorPos);
        prevCursorPos = newPos;
        haveCursor = true;
    }
    void mouseButton(bool action) { freecam.moving = action; }

    // ---- the ImGui widgets of Renderer_impl::tonemap / ::denoise by header/label, clamped like the sliders
    static void set(Renderer_impl& r, std::string const& label, std::string const& value) {
        auto num = [&](float lo, float hi) { return std::min(std::max(float(std::atof(value.c_str())), lo), hi); };
        auto pick = [&](std::initializer_list<char const*> names) {
            int i = 0;
            for (char const* n : names) {
                if (value == n) return i;
                i++;
            }
            throw std::runtime_error("front end: " + label + " has no entry " + value);
        };
        auto& u = r.uchimuraParams;
        auto& a = r.amdParams;
        if (label == "Tonemapper/Exposure") r.exposure = num(0.1f, 10.0f);
        else if (label == "Tonemapper/Algorithm") r.tonemapMode = TonemapMode(pick({"Linear", "Reinhard", "Hable", "ACES", "Uchimura", "AMD"}));
        else if (label == "Tonemapper/HDR peak") (r.tonemapMode == TonemapMode::Reinhard ? r.reinhardMax : a.hdrMax) = num(1.0f, 32.0f);
        else if (label == "Tonemapper/Max brightness") u.maxBrightness = num(1.0f, 10.0f);
        else if (label == "Tonemapper/Contrast") (r.tonemapMode == TonemapMode::Uchimura ? u.contrast : a.contrast) = r.tonemapMode == TonemapMode::Uchimura ? num(0.1f, 2.4f) : num(0.5f, 4.0f);
        else if (label == "Tonemapper/Linear start") u.linearStart = num(0.01f, 0.9f);
        else if (label == "Tonemapper/Linear length") u.linearLength = num(0.0f, 0.9f);
        else if (label == "Tonemapper/Black tightness") u.blackTightness = num(1.0f, 3.0f);
        else if (label == "Tonemapper/Pedestal") u.pedestal = num(0.0f, 1.0f);
        else if (label == "Tonemapper/Shoulder") a.shoulder = num(0.9f, 1.0f);
        else if (label == "Tonemapper/Mid in") a.midIn = num(0.01f, 1.0f);
        else if (label == "Tonemapper/Mid out") a.midOut = num(0.01f, 0.99f);
        else if (label == "Denoiser/Algorithm") r.denoiseMode = DenoiseMode(pick({"None", "Bilateral"}));
        else if (label == "Denoiser/Sigma") r.bilateralParams.sigma = num(0.1f, 10.0f);
        else if (label == "Denoiser/kSigma") r.bilateralParams.kSigma = num(0.1f, 4.0f);
        else if (label == "Denoiser/Threshold") r.bilateralParams.threshold = num(0.01f, 1.0f);
        // settings the reference compiles in (secondaryRays.comp:128-129) or does not have
        else if (label == "Pathtracer/Samples") r.pathtracer.samples = u32(std::max(1, std::atoi(value.c_str())));
        else if (label == "Pathtracer/Bounces") r.pathtracer.bounces = u32(std::max(0, std::atoi(value.c_str())));
        else if (label == "Pathtracer/Sun sampling") r.pathtracer.sunSampling = std::atoi(value.c_str()) != 0;
        else if (label == "Pathtracer/Sky at hit") r.pathtracer.skyAtHit = std::atoi(value.c_str()) != 0;
        else if (label == "Pathtracer/Aerial perspective") r.pathtracer.aerialPerspective = std::atoi(value.c_str()) != 0;
        else throw std::runtime_error("front end: unknown widget " + label);
    }

    // Window::poll + the ImGui frame: deliver this frame's events; returns false when this is the script's last frame
    auto beginFrame(int frame, Renderer_impl& r) -> bool {
        presents.clear();
        while (next < events.size() && events[next].frame <= frame) {
            Event const& e = events[next++];
            if (e.kind == "key") key(e.a, e.b == "down");
            else if (e.kind == "cursor") cursor(vec2{float(std::atof(e.a.c_str())), float(std::atof(e.b.c_str()))});
            else if (e.kind == "mouse") mouseButton(e.a == "down");
            else if (e.kind == "set") set(r, e.a, e.b);
            else if (e.kind == "present") presents.push_back(e.a);
            else if (e.kind == "dt") fixedFrameTime = float(std::atof(e.a.c_str()));
            else throw std::runtime_error("front-end script: unknown event " + e.kind);
        }
        return next < events.size();
    }

    // main.cpp mixes textual standard headers with module imports: it sees no standard containers of this interface
    [[nodiscard]] auto wantsPresent() const -> bool { return !presents.empty(); }
    void presentAll(std::uint8_t const* rgba, u32 w, u32 h) const {
        for (auto const& path : presents) present(path.c_str(), rgba, w, h);
    }

    // blitAndPresent: RGBA8 framebuffer -> binary PPM
    static void present(char const* path, std::uint8_t const* rgba, u32 w, u32 h) {
        FILE* f = std::fopen(path, "wb");
        if (!f) throw std::runtime_error(std::string("cannot write ") + path);
        std::fprintf(f, "P6\n%u %u\n255\n", w, h);
        for (std::size_t i = 0; i < std::size_t(w) * h; i++) std::fwrite(&rgba[4 * i], 1, 3, f);
        std::fclose(f);
    }

private:
    static auto trim(char const* s) -> std::string {
        std::string r = s;
        auto const b = r.find_first_not_of(" \t\r\n");
        auto const e = r.find_last_not_of(" \t\r\n");
        return b == std::string::npos ? std::string() : r.substr(b, e - b + 1);
    }
    std::vector<Event> events;
    std::size_t next = 0;
    vec2 prevCursorPos = {0.0f, 0.0f};
    bool haveCursor = false;
};
