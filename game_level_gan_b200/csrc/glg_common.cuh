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

// One track record in HBM / shared memory (include/glg_b200.h): 3*N float2,
//   line[0..N)   = right boundary REVERSED (line[v] = right[N-1-v])
//   line[N..2N)  = left boundary
//   centre[0..N) = centre points
// so that `line` is the polyline right-end ... start line ... left-end (the reference's own
// line_bounds order, games/race.py:175) and wall w of the polyline joins line[w] and line[w+1]:
// w < N-1 right walls, w == N-1 the start line, w >= N left walls (2N-1 walls).
struct TrackView {
    const float2* line;
    const float2* centre;
    int N;
};

}  // namespace glg
