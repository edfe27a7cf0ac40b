// minote.modules.denoiser -- Denoiser::bilateral (src/gfx/modules/denoiser.ixx:20-97): same name, argument
// order, BilateralParams struct and defaults.  The push constants {sigma, kSigma, threshold, nearPlane,
// frameCounter} (denoiser.ixx:78-91) go to the C ABI by value; the output is the RGBA8 image of denoiser.ixx:56.
module;
#include "../../include/minotert.h"
export module minote.modules.denoiser;
import minote.camera;
import minote.cuda;

export class Denoiser : Cuda {
public:
    struct BilateralParams {
        float sigma;
        float kSigma;
        float threshold;

        static auto make_default() -> BilateralParams { return BilateralParams{5.0f, 2.0f, 0.12f}; }  // :27-33
    };

    auto bilateral(DeviceImage color, DeviceImage depth, DeviceImage normal, Camera const& camera, BilateralParams params)
        -> DeviceImage {
        (void)color; (void)depth; (void)normal;  // resident in the context; passed for interface parity
        Cuda::serv->check(mrt_denoise_bilateral(Cuda::serv->ctx, params.sigma, params.kSigma, params.threshold, camera.nearPlane,
                                                Cuda::serv->frameCount()));
        return DeviceImage{MRT_BUF_DENOISED};
    }
};

export using BilateralParams = Denoiser::BilateralParams;
