// glg_exact.cuh - fp32 primitives with the reference's exact rounding behaviour.
//
// The reference (torch on CPU) evaluates every geometric predicate as separately rounded fp32
// mul / mul / sub (games/race.py:238, 300-301), the 2x2 rotation as mul, mul, add (:324), and the
// 2-norm as sqrt(fma(y, y, x*x)) (ATen norm kernel) - SURVEY.md 8.2, re-verified against the
// reference-generated fixtures through oracle/race_oracle.c.  nvcc would contract a*b-c*d into an
// FMA, which changes signs near zero, so everything that must match bit-for-bit goes through the
// explicit round-to-nearest intrinsics below (the file is also compiled with --fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glg {

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
// a*b - c*d with three roundings
__device__ __forceinline__ float det2(float a, float b, float c, float d) {
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
__device__ __forceinline__ int sgn(float v) { return (v > 0.f) - (v < 0.f); }
// torch.norm(p=2) over a pair: sqrt(fma(y, y, x*x))
__device__ __forceinline__ float norm2(float x, float y) {
    return __fsqrt_rn(__fmaf_rn(y, y, __fmul_rn(x, x)));
}

struct P2 { float x, y; };

// orientation value of r with respect to p->q  (games/race.py:230-238; game_helpers.cpp:22-37)
__device__ __forceinline__ float turn_val(P2 p, P2 q, P2 r) {
    const float ax = xsub(q.x, p.x), ay = xsub(q.y, p.y);
    const float bx = xsub(r.x, q.x), by = xsub(r.y, q.y);
    return det2(ay, bx, ax, by);
}
__device__ __forceinline__ int turn(P2 p, P2 q, P2 r) { return sgn(turn_val(p, q, r)); }

// r inside the axis-aligned box of p, q  (games/race.py:240-245; game_helpers.cpp:39-47)
__device__ __forceinline__ bool in_box(P2 p, P2 q, P2 r) {
    return r.x <= fmaxf(p.x, q.x) && r.x >= fminf(p.x, q.x) &&
           r.y <= fmaxf(p.y, q.y) && r.y >= fminf(p.y, q.y);
}

// wall (p,q) against probe (a,b), games/race.py:248-269.
// Returns the "collision" mask of special=True (general case + special cases 2..4);
// start_on = special case 1 (probe start a lies on the wall).
__device__ __forceinline__ bool cross_tables(P2 p, P2 q, P2 a, P2 b, bool& start_on) {
    const int o1 = turn(p, q, a);
    const int o2 = turn(p, q, b);
    const int o3 = turn(a, b, p);
    const int o4 = turn(a, b, q);
    bool hit = (o1 != o2) && (o3 != o4);
    start_on = (o1 == 0) && in_box(p, q, a);
    hit = hit || ((o2 == 0) && in_box(p, q, b));
    hit = hit || ((o3 == 0) && in_box(a, b, p));
    hit = hit || ((o4 == 0) && in_box(a, b, q));
    return hit;
}

// segments intersect incl. touching (special=False, games/race.py:269; game_helpers.cpp:49-66)
__device__ __forceinline__ bool segments_cross(P2 p, P2 q, P2 a, P2 b) {
    bool so;
    const bool hit = cross_tables(p, q, a, b, so);
    return hit || so;
}

// Same result as ray_wall_t below, arranged for the common case: when none of the four orientation values is
// zero, no special case of games/race.py:256-267 can fire and the general case is a test of sign BITS.
__device__ __forceinline__ float ray_wall_t_fast(P2 p, P2 q, P2 s, P2 d, P2 f) {
    const float wqx = xsub(q.x, p.x), wqy = xsub(q.y, p.y);
    const float fsx = xsub(f.x, s.x), fsy = xsub(f.y, s.y);
    const float v1 = det2(wqy, xsub(s.x, q.x), wqx, xsub(s.y, q.y));       // turn_val(p, q, s)
    const float v2 = det2(wqy, xsub(f.x, q.x), wqx, xsub(f.y, q.y));       // turn_val(p, q, f)
    const float v3 = det2(fsy, xsub(p.x, f.x), fsx, xsub(p.y, f.y));       // turn_val(s, f, p)
    const float v4 = det2(fsy, xsub(q.x, f.x), fsx, xsub(q.y, f.y));       // turn_val(s, f, q)
    const float INF_ = __int_as_float(0x7f800000);
    bool hit, so = false;
    if (fabsf(v1) > 0.f && fabsf(v2) > 0.f && fabsf(v3) > 0.f && fabsf(v4) > 0.f) {   // false for +-0 and for NaN
        const unsigned x = (__float_as_uint(v1) ^ __float_as_uint(v2)) & (__float_as_uint(v3) ^ __float_as_uint(v4));
        hit = (x >> 31) != 0u;                                  // o1 != o2 && o3 != o4 with all four non-zero
    } else {
        hit = cross_tables(p, q, s, f, so);
        hit = hit && !so;
    }
    const float psx = xsub(p.x, s.x), psy = xsub(p.y, s.y);
    const float num = det2(psy, wqx, psx, wqy);                 // :300
    const float den = det2(d.y, wqx, d.x, wqy);                 // :301
    float t = INF_;
    if (so) t = 0.f;                                            // :303
    else if (hit) t = xdiv(num, den);                           // :304
    if (t < 0.f) t = INF_;                                      // :306
    return t;
}

// ray parameter t of wall (p,q) for the ray (s, d) with far point f, games/race.py:287-306.
// Returns +inf for "no hit"; may return NaN (0/0) exactly where the reference does.
__device__ __forceinline__ float ray_wall_t(P2 p, P2 q, P2 s, P2 d, P2 f) {
    bool so;
    bool hit = cross_tables(p, q, s, f, so);
    hit = hit && !so;                                           // :292
    const float wqx = xsub(q.x, p.x), wqy = xsub(q.y, p.y);
    const float psx = xsub(p.x, s.x), psy = xsub(p.y, s.y);
    const float num = det2(psy, wqx, psx, wqy);                 // :300
    const float den = det2(d.y, wqx, d.x, wqy);                 // :301
    float t = __int_as_float(0x7f800000);
    if (so) t = 0.f;                                            // :303
    else if (hit) t = xdiv(num, den);                           // :304
    if (t < 0.f) t = __int_as_float(0x7f800000);                // :306
    return t;
}

}  // namespace glg
