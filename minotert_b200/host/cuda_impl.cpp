// Implementation unit of minote.cuda: the throwing paths (kept out of the module interface).
module;
#include <stdexcept>
#include <string>

#include "../../include/minotert.h"
module minote.cuda;

Cuda_impl::Cuda_impl(int device) {
    if (int s = mrt_create(device, &ctx); s != MRT_OK)
        throw std::runtime_error(std::string("mrt_create failed: ") + mrt_last_error(nullptr));
}

void Cuda_impl::fail(int status) const {
    throw std::runtime_error(std::string("minotert: ") + mrt_last_error(ctx) + " (status " + std::to_string(status) + ")");
}

void Cuda_impl::raise(char const* message) const { throw std::logic_error(message); }
