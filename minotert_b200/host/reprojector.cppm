// minote.modules.reprojector -- Reprojector::accumulate: temporal accumulation of the path tracer's image along the
// G-buffer's motion vectors (SURVEY 8f rank 2).  The reference has no such module: Pathtracer::primaryRays writes
// GBuffer::motion (src/gfx/modules/pathtracer.ixx:63-69, src/gpu/primaryRay.comp:73-75) and Renderer_impl::draw never
// reads it (src/gfx/renderer.ixx:61).  Written in the style of its sibling modules (Denoiser, Tonemapper): images in,
// parameter struct, image out.
module;
#include "../../include/minotert.h"
export module minote.modules.reprojector;
import minote.cuda;

export class Reprojector : Cuda {
public:
    struct Params {
        float maxHistory;  // cap of the per-pixel history length: a new frame weighs at least 1 / (maxHistory + 1)

        static auto make_default() -> Params { return Params{32.0f}; }
    };

    // color: the frame's radiance; visibility / motion: this frame's G-buffer (prevView = the previous frame's view)
    auto accumulate(DeviceImage color, DeviceImage visibility, DeviceImage motion, Params params, bool reset = false) -> DeviceImage {
        (void)color; (void)visibility; (void)motion;  // resident in the context; passed for interface parity
        Cuda::serv->check(mrt_temporal_accumulate(Cuda::serv->ctx, params.maxHistory, reset ? MRT_TEMPORAL_RESET : 0u));
        return DeviceImage{MRT_BUF_TEMPORAL};
    }
};

export using ReprojectorParams = Reprojector::Params;
