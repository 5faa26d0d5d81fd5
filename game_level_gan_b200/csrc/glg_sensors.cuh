// glg_sensors.cuh - wall collision and ray-cast sensors of Race.step for one car (= one warp).
//
// Reference: games/race.py:385-432 (collision with walls / finish line through _segment_collisions,
// :213-269) and :459-489 (sensors through _smallest_distance, :271-308).
//
// For a car at position s with heading nd there are O rays; the reading of ray i is the min over all
// 2N-1 walls (right walls, start line, left walls - the finish line is not a wall) of the ray
// parameter t given by the reference formula.
//
//   BRUTE : every ray x every wall and every wall x path with the literal formula.
//   FAST  : one pass over the polyline vertices that (1) bins every vertex into the angular sector
//           between two rays and emits only the (wall, ray) pairs that can possibly hit, (2) tests the
//           wall's box against the car's path box; candidates of both kinds are then evaluated with
//           the SAME literal formula, so every number that is reported comes out of the reference's
//           arithmetic.  See the comment at scan_fast for why the pruning is exact.
#pragma once
#include "glg_common.cuh"
#include "glg_exact.cuh"

namespace glg {

constexpr int QUEUE_CAP = 192;   // (wall, ray) candidates evaluated per car before falling back to BRUTE

struct SensorScratch {           // per-warp shared memory (fixed part)
    float4 ray[GLG_MAX_RAYS];    // dx, dy, far x, far y per ray
    int tmin[GLG_MAX_RAYS];      // running min of t as ordered int bits (t >= 0, -0.0 or +inf)
    unsigned nan_mask;           // rays that saw a NaN
    int pad[3];
    unsigned short queue[QUEUE_CAP];
};

// shared memory carve-up of the step kernel:
//   [record 3N float2][mbarrier 16 B][P x SensorScratch][P x maskbuf(2N u32, padded to 32)]
__host__ __device__ inline size_t smem_barrier_offset(int N) {
    return ((size_t)3 * N * sizeof(float2) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t smem_scratch_offset(int N) { return smem_barrier_offset(N) + 16; }
__host__ __device__ inline int maskbuf_len(int N) { return ((2 * N + 31) / 31 + 1) * 32; }
__host__ __device__ inline size_t smem_maskbuf_offset(int N, int P) {
    return smem_scratch_offset(N) + (size_t)P * sizeof(SensorScratch);
}
__host__ __device__ inline size_t smem_total(int N, int P) {
    return smem_maskbuf_offset(N, P) + (size_t)P * maskbuf_len(N) * sizeof(unsigned);
}

// wall w of the polyline in the reference's orientation (games/race.py:166-168): right walls and the
// start line run against the polyline direction (right[j] -> right[j+1], left[0] -> right[0]).
__device__ __forceinline__ void wall_by_line_index(const TrackView& tv, int w, P2& p, P2& q) {
    const float2 a = tv.line[w], b = tv.line[w + 1];
    const bool rev = w < tv.N;
    p = rev ? P2{b.x, b.y} : P2{a.x, a.y};
    q = rev ? P2{a.x, a.y} : P2{b.x, b.y};
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

// ---- BRUTE ----------------------------------------------------------------------------------
__device__ __noinline__ bool collide_brute(const TrackView& tv, P2 op, P2 np) {
    bool hit = false;
    for (int w = lane_id(); w < 2 * tv.N - 1; w += 32) {                    // race.py:406-407
        P2 p, q;
        wall_by_line_index(tv, w, p, q);
        hit = hit || segments_cross(p, q, op, np);
    }
    return __any_sync(FULL, hit);
}

__device__ __noinline__ float sensors_brute(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd) {
    const int lane = lane_id();
    const int O = pr.num_rays;
    float mine = INF;
    for (int i = 0; i < O; ++i) {
        P2 d, f;
        ray_setup(pr, i, s, nd, d, f);
        float t = INF;
        bool nan = false;
        for (int w = lane; w < 2 * tv.N - 1; w += 32) {
            P2 p, q;
            wall_by_line_index(tv, w, p, q);
            const float tw = ray_wall_t(p, q, s, d, f);
            if (tw != tw) nan = true;
            else t = fminf(t, tw);
        }
        if (nan) t = __int_as_float(0x7fc00000);
        t = warp_min_nan(t);
        if (lane == i) mine = t;
    }
    return mine;
}

// ---- FAST -----------------------------------------------------------------------------------
// Exactness of the sensor pruning.  Seen from the car, ray i points at angle theta_i = -pi + i*2pi/O
// relative to the heading (torch.linspace, race.py:462); in "sector units" f = (phi + pi) * O / 2pi ray i
// sits at f = i.  In the reference formula a wall (p,q) can only yield a finite t for ray i if
//   (a) i lies in the angular span [f_p, f_q] of the wall (taken the short way, which the sign of
//       cross(p-s, q-s) resolves), widened by a margin m, or
//   (b) an end point lies within eps_perp of the LINE of ray i - then the fp32 signs o3/o4 of
//       race.py:238 are not trustworthy - which makes ray i and its opposite i+O/2 candidates, or
//   (c) the car lies (to fp32 resolution) on the wall's own line, so o1 is not trustworthy - then
//       every ray is a candidate for that wall.
// m covers the atan2 approximation (< 2e-5 rad), the fp32 evaluation of the frame coordinates, and
// the worst-case rounding of o3/o4 (|far - s| ~ 1000: the sign of R x (p - far) is decided within
// ~0.35/1000 units of perpendicular distance; eps_perp = 2e-3 leaves a factor 5).
// Preconditions, checked per car, else the car takes the brute-force path: |heading|^2 in [0.5, 2]
// and every point within 200 units (far points are then outside every wall's box: special case 2 of
// race.py:265 cannot fire).
// Collision pruning: a wall whose box (plus 1e-4) does not meet the path's box cannot satisfy any
// special case (they contain a box test themselves), and the general case could only misfire for
// four points collinear to ~1e-5 with disjoint boxes.
// tests/test_race_gpu.py compares FAST with BRUTE and with the oracles bit-for-bit.
constexpr float EPS_PERP = 2e-3f;
constexpr float ETA_ANGLE = 1e-4f;
constexpr float BOX_MARGIN = 1e-4f;
constexpr float PI_F = 3.14159265358979f;

__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// atan2 with |error| < 2e-5 rad (Abramowitz & Stegun 4.4.47 on [0,1] + octant folding)
__device__ __forceinline__ float atan2_approx(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float r = mn * rcp_fast(fmaxf(mx, 1e-30f));
    const float r2 = r * r;
    float pl = fmaf(r2, 0.0208351f, -0.0851330f);
    pl = fmaf(pl, r2, 0.1801410f);
    pl = fmaf(pl, r2, -0.3302995f);
    pl = fmaf(pl, r2, 0.9998660f);
    float th = pl * r;
    th = (ay > ax) ? (0.5f * PI_F - th) : th;
    th = (x < 0.f) ? (PI_F - th) : th;
    return copysignf(th, y);
}

__device__ __forceinline__ void queue_flush(const TrackView& tv, P2 s, SensorScratch* sc, int qn) {
    const int lane = lane_id();
    __syncwarp();
    for (int e = lane; e < qn; e += 32) {
        const int code = sc->queue[e];
        const int w = code >> 5, i = code & 31;
        P2 p, q;
        wall_by_line_index(tv, w, p, q);
        const float4 r = sc->ray[i];
        const float t = ray_wall_t(p, q, s, P2{r.x, r.y}, P2{r.z, r.w});
        if (t != t) atomicOr(&sc->nan_mask, 1u << i);
        else atomicMin(&sc->tmin[i], __float_as_int(t));     // int order == float order on {-0, [0, inf]}
    }
    __syncwarp();
}

struct ScanResult {
    bool wall_hit;   // the path op -> s crosses a wall (valid if need_col)
    bool safe;       // preconditions of the pruning held; otherwise the caller must use sensors_brute
    int queued;      // candidates in the queue
};

// OC: compile-time number of rays (0 = take it from pr at run time)
// maskbuf: per-warp u32[maskbuf_len(N)]; slot pass*32+lane holds the non-empty candidate-ray mask of the wall
// that lane owned in that pass (written and later re-read by the same lane; `busy` remembers which passes).
template <int OC>
__device__ __forceinline__ ScanResult scan_fast(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd, P2 op,
                                                bool need_col, SensorScratch* sc, unsigned* maskbuf)
{
    const int lane = lane_id();
    const int O = OC ? OC : pr.num_rays;        // even (the host routes odd O to BRUTE)
    const int N = tv.N;
    const int V = 2 * N;
    const int halfO = O >> 1;
    const unsigned all_rays = (O == 32) ? FULL : ((1u << O) - 1u);

    if (lane < O) {
        P2 d, f;
        ray_setup(pr, lane, s, nd, d, f);
        sc->ray[lane] = make_float4(d.x, d.y, f.x, f.y);
        sc->tmin[lane] = 0x7f800000;
    }
    if (lane == 0) sc->nan_mask = 0;

    const float sect = (float)O * (0.5f / PI_F);         // radians -> sector units
    const float m_eta = ETA_ANGLE * sect;
    const float m_eps = EPS_PERP * sect;
    const float fO = (float)O, fhalf = 0.5f * (float)O;
    // path box relative to s (empty when no collision test is wanted)
    const float ox = op.x - s.x, oy = op.y - s.y;
    const float bx0 = need_col ? fminf(ox, 0.f) - BOX_MARGIN : INF, bx1 = need_col ? fmaxf(ox, 0.f) + BOX_MARGIN : -INF;
    const float by0 = fminf(oy, 0.f) - BOX_MARGIN, by1 = fmaxf(oy, 0.f) + BOX_MARGIN;

    int mine = 0;                   // candidates found by this lane
    unsigned busy = 0;              // passes (<= 32, the host routes longer tracks to BRUTE) with candidates
    float far2 = 0.f;
    bool hit = false;
    const int passes = (V - 1 + 30) / 31;
    // 31 walls per pass: lane l handles vertex 31*pass+l, lanes 0..30 own wall (v, v+1)
#pragma unroll 1
    for (int pass = 0; pass < passes; ++pass) {
        const int v = pass * 31 + lane;
        const float2 pt = tv.line[min(v, V - 1)];
        const float ux = pt.x - s.x, uy = pt.y - s.y;
        const float r2 = fmaf(ux, ux, uy * uy);
        far2 = fmaxf(far2, r2);
        // frame coordinates: (fa, fb) = (u . nd, u . (nd.y, -nd.x)); ray i has angle theta_i there
        const float fa = fmaf(ux, nd.x, uy * nd.y);
        const float fb = fmaf(ux, nd.y, -(uy * nd.x));
        const float f = fmaf(atan2_approx(fb, fa), sect, fhalf);          // [0, O]
        const float m = fmaf(m_eps, rsqrt_fast(fmaxf(r2, 1e-12f)), m_eta);
        const float nr = rintf(f);
        const unsigned near = __ballot_sync(FULL, fabsf(f - nr) <= m);    // vertices on (the line of) a ray
        const float ux1 = __shfl_down_sync(FULL, ux, 1), uy1 = __shfl_down_sync(FULL, uy, 1);
        const float f1 = __shfl_down_sync(FULL, f, 1), m1 = __shfl_down_sync(FULL, m, 1);
        const float r21 = __shfl_down_sync(FULL, r2, 1);
        const bool active = lane < 31 && v + 1 < V;

        // -- collision prefilter: wall box against path box --
        const bool cflag = active && !(fmaxf(ux, ux1) < bx0 || fminf(ux, ux1) > bx1 ||
                                       fmaxf(uy, uy1) < by0 || fminf(uy, uy1) > by1);
        if (__any_sync(FULL, cflag)) {
            if (cflag) {
                P2 p, q;
                wall_by_line_index(tv, v, p, q);
                hit = hit || segments_cross(p, q, op, s);                  // race.py:406
            }
        }

        // -- sensor candidates --
        unsigned mask = 0;
        if (active) {
            const float cr = fmaf(ux, uy1, -(uy * ux1));                   // cross(u_p, u_q)
            const float tau = fmaf(5e-7f, r2 + r21, 2e-6f);
            const float mm = fmaxf(m, m1);
            if (fabsf(cr) <= tau || mm > 0.45f) {
                mask = all_rays;                                           // (c), or a point almost at the car
            } else {
                // span from f to f1 the short way; (fa, fb) is a reflection of (x, y), so cr > 0 <=> f decreases
                const float lo = cr > 0.f ? f1 : f;
                float hi = cr > 0.f ? f : f1;
                hi = hi < lo ? hi + fO : hi;                               // wrap through f = O == 0
                const int ilo = __float2int_ru(lo - mm), ihi = __float2int_rd(hi + mm);
                int cnt = ihi - ilo + 1;                                   // rays inside the widened span
                if (cnt > 0) {
                    cnt = min(cnt, O);
                    const int st = ilo >= O ? ilo - O : ilo;               // ilo in [0, O]
                    const unsigned run = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);
                    mask = (st == 0) ? (run & all_rays) : (((run << st) | (run >> (O - st))) & all_rays);
                    // (b): an end point on the line of a ray also makes the opposite ray a candidate
                    const unsigned nb = (near >> lane) & 3u;
                    if (nb) {
                        if (nb & 1u) { int r = (int)nr + halfO; r = r >= O ? r - O : r; mask |= 1u << r; }
                        if (nb & 2u) { int r = (int)rintf(f1) + halfO; r = r >= O ? r - O : r; mask |= 1u << r; }
                    }
                }
            }
        }
        if (mask) {
            maskbuf[pass * 32 + lane] = mask;
            busy |= 1u << pass;
            mine += __popc(mask);
        }
    }
    far2 = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(far2)));       // far2 >= 0
    const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
    // ---- compact all (wall, ray) candidates of the car into the queue (exclusive prefix sum over lanes) ----
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    ScanResult res;
    res.wall_hit = __any_sync(FULL, hit);
    res.safe = d2 > 0.5f && d2 < 2.f && far2 < 200.f * 200.f && total <= QUEUE_CAP;
    res.queued = total;
    if (res.safe && mine) {
        int pos = incl - mine;
        while (busy) {
            const int pass = __ffs(busy) - 1;
            busy &= busy - 1;
            unsigned mask = maskbuf[pass * 32 + lane];
            const int v = pass * 31 + lane;
            while (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                sc->queue[pos++] = (unsigned short)((v << 5) | i);
            }
        }
    }
    return res;
}

// readings of the rays after the scan: lane i (< O) gets ray i
__device__ __forceinline__ float sensors_finish(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd,
                                                SensorScratch* sc, const ScanResult& res, int O)
{
    if (!res.safe) return sensors_brute(tv, pr, s, nd);
    queue_flush(tv, s, sc, res.queued);
    const int lane = lane_id();
    float t = INF;
    if (lane < O) {
        t = __int_as_float(sc->tmin[lane]);
        if (sc->nan_mask & (1u << lane)) t = __int_as_float(0x7fc00000);
    }
    return t;
}

}  // namespace glg
