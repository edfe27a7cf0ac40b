// minote.cuda -- the device service of the B200 build.  It takes the place of the reference's
// Vulkan service (src/sys/vulkan.ixx:44-49,106-115): owns the mrt_contexts (CUDA device + stream +
// all device memory) and is reached by the gfx modules through the same Service<T> locator pattern
// (src/util/service.ixx:13-55).  Frames in flight: the reference allocates each frame from its own arena with
// 3 frames in flight (src/gfx/renderer.ixx:36,42-43,97); here every frame in flight is one mrt_context (its own
// stream and frame buffers), `ctx` points at the one the current frame is recorded into, and nextFrame() rotates it.
// Frame context 0 owns the scene; the others borrow it (mrt_scene_share).  Status codes from the C ABI become std::runtime_error here, which
// preserves the reference's exception behaviour (stx/except.ixx, caught in App::run).
module;
#include <cstdint>

#include "../../include/minotert.h"
export module minote.cuda;

export class Cuda_impl {
public:
    static constexpr int MaxFramesInFlight = 3;  // renderer.ixx:36
    // throws std::runtime_error when no CUDA device exists (no fallback)
    explicit Cuda_impl(int device = 0, int framesInFlight = 1);
    ~Cuda_impl() {
        for (int i = inFlight - 1; i >= 0; i--) mrt_destroy(frameCtx[i]);  // borrowers before the scene's owner
    }
    Cuda_impl(Cuda_impl const&) = delete;
    auto operator=(Cuda_impl const&) -> Cuda_impl& = delete;

    // throw on any non-OK status, carrying the context's sticky message
    void check(int status) const {
        if (status != MRT_OK) fail(status, ctx);
    }
    void checkOn(mrt_context* c, int status) const {  // for calls on a frame context other than the current one
        if (status != MRT_OK) fail(status, c);
    }
    [[noreturn]] void raise(char const* message) const;  // std::logic_error, defined in cuda_impl.cpp
    [[nodiscard]] auto frameCount() const -> std::uint32_t { return frames; }
    // vuk::Context::next_frame(): 1 on the first frame (renderer.ixx:41,52); rotate = false keeps recording into the
    // same frame context (progressive accumulation carries its accumulator from frame to frame)
    void nextFrame(bool rotate = true) {
        frames += 1;
        if (rotate) slot = (slot + 1) % inFlight;
        ctx = frameCtx[slot];
    }
    [[nodiscard]] auto framesInFlight() const -> int { return inFlight; }
    [[nodiscard]] auto frameSlot() const -> int { return slot; }
    [[nodiscard]] auto frameContext(int i) const -> mrt_context* { return frameCtx[i]; }
    [[nodiscard]] auto owner() const -> mrt_context* { return frameCtx[0]; }  // holds the scene

    mrt_context* ctx = nullptr;  // the frame context being recorded into

private:
    [[noreturn]] void fail(int status, mrt_context* on) const;  // defined in cuda_impl.cpp (keeps <string> out of importers)
    mrt_context* frameCtx[MaxFramesInFlight] = {nullptr, nullptr, nullptr};
    int inFlight = 1, slot = 0;
    std::uint32_t frames = 0;
};

// Service locator in the style of src/util/service.ixx:13-55: a Provider owns the instance for a scope,
// consumers inherit from the service class and reach the instance through Cuda::serv.
// (Written per service instead of as a class template: g++ 13's -fmodules-ts miscompiles the
// exception clean-up of a variadic constructor of a nested class of an exported template.)
export class Cuda {
public:
    class Provider {
    public:
        explicit Provider(int device = 0, int framesInFlight = 1) : inst(new Cuda_impl(device, framesInFlight)), prev(serv) { serv = inst; }
        ~Provider() {
            serv = prev;
            delete inst;
        }
        Provider(Provider const&) = delete;
        auto operator=(Provider const&) -> Provider& = delete;

    private:
        Cuda_impl* inst;
        Cuda_impl* prev;
    };
    static inline Cuda_impl* serv = nullptr;
};

// A rendered device image: borrowed pointer into the context, valid until that buffer is re-rendered
// at another size (what a vuk::Future resolves to after execution).
export struct DeviceImage {
    int id = -1;  // mrt_buffer_id
    [[nodiscard]] auto valid() const -> bool { return id >= 0; }
    [[nodiscard]] auto data() const -> void* {
        void* p = nullptr;
        std::size_t n = 0;
        Cuda::serv->check(mrt_buffer(Cuda::serv->ctx, id, &p, &n));
        return p;
    }
    [[nodiscard]] auto bytes() const -> std::size_t {
        void* p = nullptr;
        std::size_t n = 0;
        Cuda::serv->check(mrt_buffer(Cuda::serv->ctx, id, &p, &n));
        return n;
    }
    void readback(void* host, std::size_t n) const { Cuda::serv->check(mrt_readback(Cuda::serv->ctx, id, host, n)); }
    void readbackAsync(void* host, std::size_t n) const { Cuda::serv->check(mrt_readback_async(Cuda::serv->ctx, id, host, n)); }
};
