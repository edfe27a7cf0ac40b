// minote.modules.tonemapper -- Tonemapper::{linear,reinhard,hable,aces,uchimura,amd}
// (src/gfx/modules/tonemapper.ixx:57-373), same names, parameter structs and defaults.
module;
#include "../../include/minotert.h"
export module minote.modules.tonemapper;
import minote.cuda;

export class Tonemapper : Cuda {
public:
    struct UchimuraParams {
        float maxBrightness, contrast, linearStart, linearLength, blackTightness, pedestal;
        static auto make_default() -> UchimuraParams { return {1.0f, 1.0f, 0.22f, 0.4f, 1.33f, 0.0f}; }  // :27-36
    };
    struct AMDParams {
        float hdrMax, contrast, shoulder, midIn, midOut;
        static auto make_default() -> AMDParams { return {16.0f, 2.0f, 1.0f, 0.18f, 0.18f}; }  // :46-54
    };

    auto linear(DeviceImage input, float exposure) -> DeviceImage { return run(MRT_TONEMAP_LINEAR, input, exposure, nullptr, 0); }
    auto reinhard(DeviceImage input, float exposure, float hdrMax) -> DeviceImage {
        return run(MRT_TONEMAP_REINHARD, input, exposure, &hdrMax, 1);
    }
    auto hable(DeviceImage input, float exposure) -> DeviceImage { return run(MRT_TONEMAP_HABLE, input, exposure, nullptr, 0); }
    auto aces(DeviceImage input, float exposure) -> DeviceImage { return run(MRT_TONEMAP_ACES, input, exposure, nullptr, 0); }
    auto uchimura(DeviceImage input, float exposure, UchimuraParams const& p) -> DeviceImage {
        return run(MRT_TONEMAP_UCHIMURA, input, exposure, &p.maxBrightness, 6);
    }
    auto amd(DeviceImage input, float exposure, AMDParams const& p) -> DeviceImage {
        return run(MRT_TONEMAP_AMD, input, exposure, &p.hdrMax, 5);
    }

private:
    auto run(int mode, DeviceImage input, float exposure, float const* params, unsigned n) -> DeviceImage {
        Cuda::serv->check(mrt_tonemap(Cuda::serv->ctx, mode, exposure, params, n, input.id));
        return DeviceImage{MRT_BUF_LDR};
    }
};

export using UchimuraParams = Tonemapper::UchimuraParams;
export using AMDParams = Tonemapper::AMDParams;
