// Implementation unit of minote.cuda: the throwing paths (kept out of the module interface).
module;
#include <stdexcept>
#include <string>

#include "../../include/minotert.h"
module minote.cuda;

Cuda_impl::Cuda_impl(int device, int framesInFlight) {
    if (framesInFlight < 1 || framesInFlight > MaxFramesInFlight)
        throw std::logic_error("frames in flight must be 1.." + std::to_string(MaxFramesInFlight));
    for (int i = 0; i < framesInFlight; i++) {
        if (int s = mrt_create(device, &frameCtx[i]); s != MRT_OK) {
            std::string const why = mrt_last_error(nullptr);
            for (int k = i - 1; k >= 0; k--) mrt_destroy(frameCtx[k]);
            throw std::runtime_error("mrt_create failed: " + why);
        }
    }
    inFlight = framesInFlight;
    ctx = frameCtx[0];
    // With several frames in flight the persistent traversal grids of different frames should co-run instead of
    // each filling every SM: 3 of the 6 resident CTAs per SM measured best (config 2, 3 frames: 3824 -> 3958
    // Mrays/s; 2 frames: 3757 -> 3820; profiles/r1_results.md).
    if (inFlight > 1)
        for (int i = 0; i < inFlight; i++) mrt_set_option(frameCtx[i], "trace_ctas_per_sm", 3);
}

void Cuda_impl::fail(int status, mrt_context* on) const {
    throw std::runtime_error(std::string("minotert: ") + mrt_last_error(on) + " (status " + std::to_string(status) + ")");
}

void Cuda_impl::raise(char const* message) const { throw std::logic_error(message); }
