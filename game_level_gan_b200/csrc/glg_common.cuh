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

// Optional per-phase cycle accounting of the step kernel (debug builds with -DGLG_PHASE_CLOCKS only):
// lane 0 of every warp adds the cycles since its previous mark to g_phase[i]; read with glg_debug_phases().
#ifdef GLG_PHASE_CLOCKS
static __device__ unsigned long long g_phase[32];   // one copy per translation unit; only glg_race.cu uses it
static __device__ unsigned long long g_trace[8192 * 4];   // per CTA (last launch wins): start, after wait, end [ns], SM id
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define GLG_TRACE(slot) do { if (threadIdx.x == 0 && blockIdx.x < 8192) glg::g_trace[blockIdx.x * 4 + (slot)] = ((slot) == 3) ? (unsigned long long)glg::smid() : glg::globaltimer_ns(); } while (0)
#define GLG_MARK(i)                                                                  \
    do {                                                                             \
        const long long now_ = clock64();                                            \
        if ((threadIdx.x & 31) == 0) atomicAdd(&glg::g_phase[i], (unsigned long long)(now_ - mark_)); \
        mark_ = now_;                                                                \
    } while (0)
#define GLG_MARK_INIT long long mark_ = clock64()
#else
#define GLG_MARK(i) do {} while (0)
#define GLG_MARK_INIT do {} while (0)
#define GLG_TRACE(slot) do {} while (0)
#endif

#ifndef GLG_CHAIN_BACKOFF_NS
#define GLG_CHAIN_BACKOFF_NS 200    // sleep between polls of a chained car's stamp (glg_race_rollout)
#endif

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
