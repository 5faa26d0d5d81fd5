// glg_sensors.cuh - the ray-cast sensors of Race.step (games/race.py:459-489, 271-308).
//
// For one car (one warp): O rays from the car's position, directions = heading rotated by the O
// fixed angles; the reading of a ray is min over all 2(N-1)+1 walls (right, left, start line - the
// finish line is not a wall) of the ray parameter t given by the reference formula.
//
//   sensors_brute : every ray x every wall with the literal formula (the reference loop).
//   sensors_fast  : exact angular pruning, see the comment at sensors_fast.
//
// Both return, in lane i (< O), the un-clamped reading of ray i (+inf = nothing hit, NaN where the
// reference produces NaN).
#pragma once
#include "glg_common.cuh"
#include "glg_exact.cuh"

namespace glg {

constexpr int QUEUE_CAP = 192;   // (wall, ray) candidates buffered per warp before a dense evaluation

struct SensorScratch {           // per-warp shared memory
    float ray[GLG_MAX_RAYS][6];  // dx, dy, fx, fy per ray (+2 pad)
    int tmin[GLG_MAX_RAYS];      // running min of t as ordered int bits (t >= 0 or +inf)
    unsigned nan_mask;           // rays that saw a NaN
    int pad[3];
    unsigned short queue[QUEUE_CAP];
};

__host__ __device__ inline size_t sensor_scratch_offset(int N) {
    return ((size_t)3 * N * sizeof(float2) + 15) & ~(size_t)15;
}

// wall j in the reference's order: right 0..S-1, left 0..S-1, start line (games/race.py:166-172)
__device__ __forceinline__ void wall_points(const TrackView& tv, int j, P2& p, P2& q) {
    const int S = tv.N - 1;
    float2 a, b;
    if (j < S) { a = tv.right[j]; b = tv.right[j + 1]; }
    else if (j < 2 * S) { a = tv.left[j - S]; b = tv.left[j - S + 1]; }
    else { a = tv.left[0]; b = tv.right[0]; }
    p = P2{a.x, a.y};
    q = P2{b.x, b.y};
}

// ray i of the car: direction = heading @ R(angle_i) (games/race.py:462-470), far point :289
__device__ __forceinline__ void ray_setup(const glg_race_params& pr, int i, P2 s, P2 nd, P2& d, P2& f) {
    const float rc = pr.ray_cos[i], rs = pr.ray_sin[i];
    d = P2{xadd(xmul(nd.x, rc), xmul(nd.y, rs)), xadd(xmul(nd.x, -rs), xmul(nd.y, rc))};
    f = P2{xadd(s.x, xmul(1000.f, d.x)), xadd(s.y, xmul(1000.f, d.y))};
}

// warp-wide min with NaN propagation (torch.min semantics, games/race.py:308)
__device__ __forceinline__ float warp_min_nan(float t) {
    const bool nan = __any_sync(FULL, t != t);
    if (t != t) t = INF;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t = fminf(t, __shfl_xor_sync(FULL, t, off));
    return nan ? __int_as_float(0x7fc00000) : t;
}

__device__ __noinline__ float sensors_brute(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd,
                                            SensorScratch* /*scratch*/)
{
    const int lane = lane_id();
    const int O = pr.num_rays;
    const int walls = 2 * (tv.N - 1) + 1;
    float mine = INF;
    for (int i = 0; i < O; ++i) {
        P2 d, f;
        ray_setup(pr, i, s, nd, d, f);
        float t = INF;
        bool nan = false;
        for (int j = lane; j < walls; j += 32) {
            P2 p, q;
            wall_points(tv, j, p, q);
            const float tj = ray_wall_t(p, q, s, d, f);
            if (tj != tj) nan = true;
            else t = fminf(t, tj);
        }
        if (nan) t = __int_as_float(0x7fc00000);
        t = warp_min_nan(t);
        if (lane == i) mine = t;
    }
    return mine;
}

// ---------------------------------------------------------------------------------------------
// sensors_fast - exact angular pruning.
//
// Seen from the car, ray i points at angle theta_i = -pi + i*2pi/O relative to the heading
// (torch.linspace, race.py:462), so in "sector units" f = (phi + pi) * O / 2pi ray i sits at f = i.
// A wall (p,q) can only produce a hit for ray i in the reference formula if
//   (a) i lies in the angular span [f_p, f_q] (taken the short way, resolved by the sign of
//       cross(p-s, q-s)), widened by a margin m, or
//   (b) an end point lies within eps_perp of the LINE of ray i (then the fp32 orientation signs
//       o3/o4 of race.py:238 are not trustworthy) - this makes ray i and its opposite i+O/2
//       candidates, or
//   (c) the car lies (to fp32 resolution) on the wall's own line, so o1 is not trustworthy - then
//       every ray is a candidate for that wall.
// The margin covers the atan2 approximation (< 2e-5 rad), the fp32 evaluation of the frame
// coordinates and the worst-case rounding of the reference's o3/o4 (|R|~1000: the sign of
// R x (p - far) is decided by ~0.35/1000 units of perpendicular distance; eps_perp = 2e-3).
// Candidates are compacted into a per-warp queue and evaluated, 32 at a time, with the literal
// reference formula (ray_wall_t) - so every reported number is produced by the same arithmetic
// as the brute-force loop; pruning only removes pairs that provably evaluate to +inf.
// Preconditions checked per car, else the car falls back to sensors_brute: |heading|^2 in
// [0.5, 2], all points within 200 units of the car (far points are then outside every wall's box,
// which rules out special case 2 of race.py:265).
// tests/test_race_variants.py compares FAST against BRUTE bit-for-bit on every fixture.
// ---------------------------------------------------------------------------------------------
constexpr float EPS_PERP = 2e-3f;      // see (b)
constexpr float ETA_ANGLE = 1e-4f;     // atan2 approximation + fp32 slack, radians
constexpr float PI_F = 3.14159265358979f;

// atan2 with |error| < 2e-5 rad (A&S 4.4.47 odd polynomial on [0,1] + octant folding)
__device__ __forceinline__ float atan2_approx(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float r = __fdividef(mn, fmaxf(mx, 1e-30f));
    const float r2 = r * r;
    float pl = fmaf(r2, 0.0208351f, -0.0851330f);
    pl = fmaf(pl, r2, 0.1801410f);
    pl = fmaf(pl, r2, -0.3302995f);
    pl = fmaf(pl, r2, 0.9998660f);
    float th = pl * r;
    if (ay > ax) th = 0.5f * PI_F - th;
    if (x < 0.f) th = PI_F - th;
    return copysignf(th, y);
}

__device__ __forceinline__ void queue_flush(const TrackView& tv, const glg_race_params& pr, P2 s,
                                            SensorScratch* sc, int qn)
{
    const int lane = lane_id();
    __syncwarp();
    for (int e = lane; e < qn; e += 32) {
        const int code = sc->queue[e];
        const int j = code >> 5, i = code & 31;
        P2 p, q;
        wall_points(tv, j, p, q);
        const P2 d{sc->ray[i][0], sc->ray[i][1]}, f{sc->ray[i][2], sc->ray[i][3]};
        const float t = ray_wall_t(p, q, s, d, f);
        if (t != t) atomicOr(&sc->nan_mask, 1u << i);
        else atomicMin(&sc->tmin[i], __float_as_int(t));     // t >= 0 (or -0.0) or +inf: int order == float order
    }
    __syncwarp();
}

__device__ __forceinline__ float sensors_fast(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd,
                                              SensorScratch* sc)
{
    const int lane = lane_id();
    const int O = pr.num_rays;          // even (host falls back to BRUTE otherwise)
    const int N = tv.N;
    const int V = 2 * N;                // vertices in the order right[N-1..0], left[0..N-1]
    const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
    bool safe = d2 > 0.5f && d2 < 2.f;

    if (lane < O) {
        P2 d, f;
        ray_setup(pr, lane, s, nd, d, f);
        sc->ray[lane][0] = d.x; sc->ray[lane][1] = d.y; sc->ray[lane][2] = f.x; sc->ray[lane][3] = f.y;
        sc->tmin[lane] = 0x7f800000;
    }
    if (lane == 0) sc->nan_mask = 0;
    __syncwarp();

    const float sect = (float)O * (0.5f / PI_F);      // radians -> sector units
    const float m_eta = ETA_ANGLE * sect;
    const float m_eps = EPS_PERP * sect;
    const unsigned all_rays = (O == 32) ? FULL : ((1u << O) - 1u);
    const int halfO = O >> 1;
    int qn = 0;
    float far2 = 0.f;

    // 31 walls per pass: lane l classifies vertex base+l, lanes 0..30 own wall (base+l, base+l+1)
    for (int base = 0; base < V - 1; base += 31) {
        const int v = base + lane;
        const bool vin = v < V;
        const int vc = vin ? v : V - 1;
        const float2 pt = (vc < N) ? tv.right[N - 1 - vc] : tv.left[vc - N];
        const float ux = pt.x - s.x, uy = pt.y - s.y;
        const float r2 = fmaf(ux, ux, uy * uy);
        far2 = fmaxf(far2, r2);
        // frame coordinates: a along the heading, bq along (nd.y, -nd.x); ray i = angle theta_i
        const float fa = fmaf(ux, nd.x, uy * nd.y);
        const float fb = fmaf(ux, nd.y, -(uy * nd.x));
        const float phi = atan2_approx(fb, fa);
        const float f = fmaf(phi, sect, 0.5f * (float)O);           // [0, O]
        const float m = fmaf(m_eps, rsqrtf(fmaxf(r2, 1e-12f)), m_eta);
        // neighbour (vertex v+1) through the warp
        const float ux1 = __shfl_down_sync(FULL, ux, 1), uy1 = __shfl_down_sync(FULL, uy, 1);
        const float f1 = __shfl_down_sync(FULL, f, 1), m1 = __shfl_down_sync(FULL, m, 1);
        const float r21 = __shfl_down_sync(FULL, r2, 1);
        unsigned mask = 0;
        if (lane < 31 && v + 1 < V) {
            const float cr = fmaf(ux, uy1, -(uy * ux1));             // cross(u_p, u_q)
            const float tau = fmaf(5e-7f, r2 + r21, 2e-6f);
            const float mm = fmaxf(m, m1);
            if (fabsf(cr) <= tau || mm > 0.45f) {
                mask = all_rays;                                     // (c), or a point almost at the car
            } else {
                // short-way span from f to f1: its direction is the sign of cr.  In this frame
                // (a, b) = (u.nd, u.(nd.y,-nd.x)) is a reflection of (x, y), so cross > 0 <=> f decreases.
                float lo = f, hi = f1;
                if (cr > 0.f) { lo = f1; hi = f; }
                if (hi < lo) hi += (float)O;                         // wrap through f = O == 0
                const int ilo = (int)ceilf(lo - mm), ihi = (int)floorf(hi + mm);
                int cnt = ihi - ilo + 1;                             // rays inside the span (+margin)
                if (cnt > 0) {
                    cnt = min(cnt, O);
                    int st = ilo % O; if (st < 0) st += O;
                    const unsigned run = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);
                    mask = (st == 0) ? (run & all_rays)                      // rotate within O bits
                                     : (((run << st) | (run >> (O - st))) & all_rays);
                    // (b): rays sharing a line with an end point: add the opposite rays
                    const float np0 = rintf(f), np1 = rintf(f1);
                    if (fabsf(f - np0) <= m) { int r = ((int)np0 + halfO) % O; mask |= 1u << r; }
                    if (fabsf(f1 - np1) <= m1) { int r = ((int)np1 + halfO) % O; mask |= 1u << r; }
                }
            }
        }
        // wall index in the reference's order for vertex pair (v, v+1)
        int wall;
        if (v < N - 1) wall = N - 2 - v;                              // right wall (right[N-2-v] -> right[N-1-v])
        else if (v == N - 1) wall = 2 * (N - 1);                      // start line
        else wall = (N - 1) + (v - N);                                // left wall
        // compact the candidates of this pass into the queue
        while (true) {
            const bool has = mask != 0;
            const unsigned bal = __ballot_sync(FULL, has);
            if (bal == 0) break;
            const int cntb = __popc(bal);
            if (qn + cntb > QUEUE_CAP) { queue_flush(tv, pr, s, sc, qn); qn = 0; }
            if (has) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                sc->queue[qn + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)((wall << 5) | i);
            }
            qn += cntb;
        }
    }
    // preconditions (uniform): every point near enough that far points are outside all wall boxes
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) far2 = fmaxf(far2, __shfl_xor_sync(FULL, far2, off));
    safe = safe && far2 < 200.f * 200.f;
    if (!safe) return sensors_brute(tv, pr, s, nd, sc);
    queue_flush(tv, pr, s, sc, qn);
    float t = INF;
    if (lane < O) {
        t = __int_as_float(sc->tmin[lane]);
        if (sc->nan_mask & (1u << lane)) t = __int_as_float(0x7fc00000);
    }
    return t;
}

}  // namespace glg
