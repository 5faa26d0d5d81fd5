// glg_common.cuh - error plumbing and small device helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/glg_b200.h"

namespace glg {

void set_error(const char* fmt, ...);
int launch_status(const char* what);

#define GLG_REQUIRE(cond, ...)                   \
    do {                                         \
        if (!(cond)) {                           \
            glg::set_error(__VA_ARGS__);         \
            return GLG_ERR_ARG;                  \
        }                                        \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;
constexpr float INF = __builtin_huge_valf();

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// points of one track record {right[N], left[N], centre[N]} (include/glg_b200.h)
struct TrackView {
    const float2* right;
    const float2* left;
    const float2* centre;
    int N;
};

}  // namespace glg
