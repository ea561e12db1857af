// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/b200track.h"

namespace b200 {
void set_error(const std::string& msg);
}

#define B200_CU_TRY(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            b200::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
            return B200TRACK_ERR_CUDA;                                                     \
        }                                                                                  \
    } while (0)
