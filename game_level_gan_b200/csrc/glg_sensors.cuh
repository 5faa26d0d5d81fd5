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
//   SCAN  : one pass over the polyline vertices that (1) bins every vertex into the angular sector
//           between two rays and emits only the (wall, ray) pairs that can possibly hit, (2) tests the
//           wall's box against the car's path box; candidates of both kinds are then evaluated with
//           the SAME literal formula, so every number that is reported comes out of the reference's
//           arithmetic.  See the comment at scan_fast for why the pruning is exact.
//   FAST  : two stages.  Stage 1 is a cheap pass over the vertices that only decides WHETHER a wall can
//           meet any ray line (sign of Im z^(O/2), z = vertex in the car frame, which vanishes exactly on
//           the O/2 ray lines) and flags the few walls near the path; stage 2 runs the SCAN analysis on
//           the flagged walls only (one wall per lane).  See scan_two_stage.
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
//   [record 3N float2][mbarrier 16 B][P x SensorScratch][P x maskbuf_len(N) u32]
__host__ __device__ inline size_t smem_barrier_offset(int N) {
    return ((size_t)3 * N * sizeof(float2) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t smem_scratch_offset(int N) { return smem_barrier_offset(N) + 16; }
// per-warp u32 scratch: SCAN keeps one candidate-ray mask per (pass, lane); FAST keeps two u16 lists of
// wall indices (sensor walls, collision walls) of list_len(N) entries each
__host__ __device__ inline int list_len(int N) { return ((2 * N + 31) / 32) * 32; }
__host__ __device__ inline int maskbuf_len(int N) {
    const int a = ((2 * N + 31) / 31 + 1) * 32, b = list_len(N);      // b u32 = 2 lists of b u16
    return a > b ? a : b;
}
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


// ---- FAST: two-stage scan ---------------------------------------------------------------------
// Stage 1 (all 2N vertices, ~30 instructions each): with z = (u . nd) + i (u x nd), u = vertex - s, the
// O/2 ray LINES of the car are exactly the zero set of Im z^(O/2) (= |z|^(O/2) sin((O/2) phi)).
// A wall whose end points are farther than Rc = 3.2 Lmax from the car (Lmax = longest wall of the track,
// glg_track_extent) subtends less than 18 degrees < one sector, so it meets a ray line iff the sign of
// Im z^9 differs at its end points; "near a ray line" (perpendicular distance <= EPS_PERP, the bound of
// the reference's own fp32 sign noise, see scan_fast) is |Im z^9| <= 9 EPS_PERP |nd| |z|^8 because
// |sin 9x| <= 9 |sin x|.  A wall is FLAGGED if the sign differs, or an end point is near a ray line, or an
// end point is closer than Rc.  An unflagged wall has both end points strictly on one side of every ray
// line, farther than EPS_PERP from it, and the car is outside the wall's box: o3 == o4 != 0 for every
// ray, so no case of games/race.py:248-269 can fire - it yields +inf for all rays, exactly what
// dropping it does.  Evaluation error of Im z^9 in fp32 is < 1e-6 |z|^9, i.e. < 1e-7 r in perpendicular
// distance (r <= 200): two orders below EPS_PERP.
// Collision walls: boxes (each widened by BOX_MARGIN) can only meet if the wall's first end point is
// within |path|_1 + Lmax + 1e-3 of s.
// Stage 2: flagged walls, one per lane, get the full analysis of scan_fast (sector of both end points
// with margins, short-way span, opposite rays, car-on-the-wall's-line case) and emit (wall, ray) pairs.
constexpr float RC_FACTOR = 3.2f;          // 1 / (2 sin(alpha/2)) for alpha = 17.98 deg < 20 deg = one sector at O = 18

struct WallRays { unsigned mask; };

__device__ __forceinline__ unsigned wall_ray_mask(float ux, float uy, float ux1, float uy1, P2 nd, int O,
                                                  float sect, float fhalf, float m_eps, float m_eta, unsigned all_rays)
{
    const float fO = (float)O;
    const int halfO = O >> 1;
    const float r2 = fmaf(ux, ux, uy * uy), r21 = fmaf(ux1, ux1, uy1 * uy1);
    const float fa = fmaf(ux, nd.x, uy * nd.y), fb = fmaf(ux, nd.y, -(uy * nd.x));
    const float fa1 = fmaf(ux1, nd.x, uy1 * nd.y), fb1 = fmaf(ux1, nd.y, -(uy1 * nd.x));
    const float f = fmaf(atan2_approx(fb, fa), sect, fhalf);
    const float f1 = fmaf(atan2_approx(fb1, fa1), sect, fhalf);
    const float m = fmaf(m_eps, rsqrt_fast(fmaxf(r2, 1e-12f)), m_eta);
    const float m1 = fmaf(m_eps, rsqrt_fast(fmaxf(r21, 1e-12f)), m_eta);
    const float cr = fmaf(ux, uy1, -(uy * ux1));                       // cross(u_p, u_q)
    const float tau = fmaf(5e-7f, r2 + r21, 2e-6f);
    const float mm = fmaxf(m, m1);
    if (fabsf(cr) <= tau || mm > 0.45f) return all_rays;               // (c), or a point almost at the car
    const float lo = cr > 0.f ? f1 : f;
    float hi = cr > 0.f ? f : f1;
    hi = hi < lo ? hi + fO : hi;                                       // wrap through f = O == 0
    const int ilo = __float2int_ru(lo - mm), ihi = __float2int_rd(hi + mm);
    int cnt = ihi - ilo + 1;                                           // rays inside the widened span
    if (cnt <= 0) return 0u;
    cnt = min(cnt, O);
    const int st = ilo >= O ? ilo - O : ilo;                           // ilo in [0, O]
    const unsigned run = (cnt >= 32) ? FULL : ((1u << cnt) - 1u);
    unsigned mask = (st == 0) ? (run & all_rays) : (((run << st) | (run >> (O - st))) & all_rays);
    const float nr = rintf(f), nr1 = rintf(f1);
    if (fabsf(f - nr) <= m) { int r = (int)nr + halfO; r = r >= O ? r - O : r; mask |= 1u << r; }     // (b)
    if (fabsf(f1 - nr1) <= m1) { int r = (int)nr1 + halfO; r = r >= O ? r - O : r; mask |= 1u << r; }
    return mask;
}

// lists: per-warp u16[2 * list_len(N)] (sensor walls, then collision walls).  extent = {Rb, Lmax} of the track.
__device__ __forceinline__ ScanResult scan_two_stage(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd, P2 op,
                                                     bool need_col, SensorScratch* sc, unsigned short* lists,
                                                     float Rb, float Lmax)
{
    constexpr int O = 18;
    const int lane = lane_id();
    const int N = tv.N;
    const int V = 2 * N;
    const unsigned all_rays = (1u << O) - 1u;
    const unsigned lt = (1u << lane) - 1u;
    unsigned short* wlist = lists;
    unsigned short* clist = lists + list_len(N);

    ScanResult res{false, false, 0};
    const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
    // preconditions of the pruning (see scan_fast): heading norm, every point within 200 units of the car
    if (!(d2 > 0.5f && d2 < 2.f && fabsf(s.x) + fabsf(s.y) + Rb < 200.f)) {
        if (need_col) res.wall_hit = collide_brute(tv, op, s);
        return res;
    }
    if (lane < O) {
        P2 d, f;
        ray_setup(pr, lane, s, nd, d, f);
        sc->ray[lane] = make_float4(d.x, d.y, f.x, f.y);
        sc->tmin[lane] = 0x7f800000;
    }
    if (lane == 0) sc->nan_mask = 0;

    // ---- stage 1 ----
    GLG_MARK_INIT;
    // Lmax excludes the start line (wall N-1, as long as the track is wide): it is appended to both lists below.
    const float Rc = fmaf(RC_FACTOR, Lmax, 1e-3f);
    const float close2 = Rc * Rc * d2 * 1.0001f;                       // thresholds on |z|^2 = r^2 d2
    const float colR = fabsf(op.x - s.x) + fabsf(op.y - s.y) + Lmax + 1e-3f;
    const float col2 = need_col ? colR * colR * d2 * 1.0001f : 0.f;    // r2z - col2 < 0 never holds for 0
    const float Kn = 9.f * EPS_PERP * 1.4143f * 1.001f;                // 9 EPS_PERP |nd| with |nd| < sqrt(2)
    const int passes = (V + 31) / 32;                                  // <= 32 (N <= 512)
    // lane l handles vertices l, l+32, ...; bit `pass` of sbits / fbits / cbits describes vertex 32*pass + l.
    // The three predicates are produced as SIGN BITS (x < 0) and shifted in with one funnel shift each; the
    // bits arrive in reverse pass order and are put right after the loop.  Vertices past the polyline (the
    // last pass may read up to 31 points of whatever follows `line` in shared memory) are masked by `own`.
    unsigned sbits = 0, fbits = 0, cbits = 0;
#pragma unroll 3
    for (int pass = 0; pass < passes; ++pass) {
        const float2 pt = tv.line[pass * 32 + lane];
        const float ux = pt.x - s.x, uy = pt.y - s.y;
        const float a = fmaf(ux, nd.x, uy * nd.y);
        const float b = fmaf(ux, nd.y, -(uy * nd.x));
        const float a2 = a * a, b2 = b * b;
        const float r2z = a2 + b2;
        const float re3 = a * fmaf(-3.f, b2, a2);                      // z^3
        const float im3 = b * fmaf(3.f, a2, -b2);
        const float im9 = im3 * fmaf(3.f, re3 * re3, -(im3 * im3));    // Im (z^3)^3
        const float r4 = r2z * r2z;
        const float near = fmaf(r4 * r4, -Kn, fabsf(im9));             // < 0: within EPS_PERP of a ray line
        const float flag = fminf(near, r2z - close2);                  // < 0: near a ray line or close to the car
        sbits = __funnelshift_l(__float_as_uint(im9), sbits, 1);
        fbits = __funnelshift_l(__float_as_uint(flag), fbits, 1);
        cbits = __funnelshift_l(__float_as_uint(r2z - col2), cbits, 1);
    }
    {
        const int sh = 32 - passes;
        sbits = __brev(sbits) >> sh;
        fbits = __brev(fbits) >> sh;
        cbits = __brev(cbits) >> sh;
    }
    GLG_MARK(5);
    // wall w = (vertex w, vertex w+1): the next vertex lives in lane+1 (same pass), or in lane 0 of the next pass
    unsigned s1 = __shfl_down_sync(FULL, sbits, 1), f1 = __shfl_down_sync(FULL, fbits, 1);
    const unsigned s0 = __shfl_sync(FULL, sbits, 0), f0 = __shfl_sync(FULL, fbits, 0);
    if (lane == 31) { s1 = s0 >> 1; f1 = f0 >> 1; }
    // walls owned by this lane: w = 32*pass + lane <= V-2, except the start line w = N-1
    const int nown = (V - 2 - lane >= 0) ? ((V - 2 - lane) >> 5) + 1 : 0;
    unsigned own = nown >= 32 ? FULL : ((1u << nown) - 1u);
    if (((N - 1) & 31) == lane) own &= ~(1u << ((N - 1) >> 5));
    unsigned wbits = ((sbits ^ s1) | fbits | f1) & own;
    cbits &= own;
    int nw, nc = 0;
    {   // exclusive prefix sum of the per-lane counts, then every lane appends its own walls
        const int cnt = __popc(wbits);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += t;
        }
        nw = __shfl_sync(FULL, incl, 31);
        int pos = incl - cnt;
        while (wbits) {
            const int pass = __ffs(wbits) - 1;
            wbits &= wbits - 1;
            wlist[pos++] = (unsigned short)(pass * 32 + lane);
        }
        if (lane == 0) wlist[nw] = (unsigned short)(N - 1);            // the start line, always
        ++nw;
    }
    if (need_col) {
        unsigned live = __ballot_sync(FULL, cbits != 0u);
        while (live) {                                                 // usually zero or one round
            if (cbits) {
                const int pass = __ffs(cbits) - 1;
                cbits &= cbits - 1;
                clist[nc + __popc(live & lt)] = (unsigned short)(pass * 32 + lane);
            }
            nc += __popc(live);
            live = __ballot_sync(FULL, cbits != 0u);
        }
        if (lane == 0) clist[nc] = (unsigned short)(N - 1);
        ++nc;
    }
    __syncwarp();
    GLG_MARK(6);

    // ---- collision: exact test of the walls near the path (race.py:406) ----
    if (nc) {
        const float ox = op.x - s.x, oy = op.y - s.y;
        const float bx0 = fminf(ox, 0.f) - BOX_MARGIN, bx1 = fmaxf(ox, 0.f) + BOX_MARGIN;
        const float by0 = fminf(oy, 0.f) - BOX_MARGIN, by1 = fmaxf(oy, 0.f) + BOX_MARGIN;
        bool hit = false;
        for (int e = lane; e < nc; e += 32) {
            const int w = clist[e];
            const float2 p0 = tv.line[w], p1 = tv.line[w + 1];
            const float ux = p0.x - s.x, uy = p0.y - s.y, ux1 = p1.x - s.x, uy1 = p1.y - s.y;
            if (!(fmaxf(ux, ux1) < bx0 || fminf(ux, ux1) > bx1 || fmaxf(uy, uy1) < by0 || fminf(uy, uy1) > by1)) {
                P2 p, q;
                wall_by_line_index(tv, w, p, q);
                hit = hit || segments_cross(p, q, op, s);
            }
        }
        res.wall_hit = __any_sync(FULL, hit);
    }

    GLG_MARK(7);
    // ---- stage 2: candidate rays of the flagged walls ----
    const float sect = (float)O * (0.5f / PI_F);
    const float m_eta = ETA_ANGLE * sect, m_eps = EPS_PERP * sect, fhalf = 0.5f * (float)O;
    int total = 0;
    bool overflow = false;
    for (int base = 0; base < nw; base += 32) {
        const int e = base + lane;
        unsigned mask = 0;
        int w = 0;
        if (e < nw) {
            w = wlist[e];
            const float2 p0 = tv.line[w], p1 = tv.line[w + 1];
            mask = wall_ray_mask(p0.x - s.x, p0.y - s.y, p1.x - s.x, p1.y - s.y, nd, O, sect, fhalf, m_eps, m_eta, all_rays);
        }
        // emit one ray of every lane per round (most walls have exactly one)
        unsigned live = __ballot_sync(FULL, mask != 0u);
        while (live) {
            const int cnt = __popc(live);
            if (total + cnt > QUEUE_CAP) { overflow = true; break; }
            if (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                sc->queue[total + __popc(live & lt)] = (unsigned short)((w << 5) | i);
            }
            total += cnt;
            live = __ballot_sync(FULL, mask != 0u);
        }
        if (overflow) break;
    }
    GLG_MARK(8);
    res.safe = !overflow;
    res.queued = total;
    return res;
}

// readings of the rays after the scan: lane i (< O) gets ray i
__device__ __forceinline__ float sensors_finish(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd,
                                                SensorScratch* sc, const ScanResult& res, int O)
{
    if (!res.safe) return sensors_brute(tv, pr, s, nd);
    GLG_MARK_INIT;
    queue_flush(tv, s, sc, res.queued);
    GLG_MARK(10);
    const int lane = lane_id();
    float t = INF;
    if (lane < O) {
        t = __int_as_float(sc->tmin[lane]);
        if (sc->nan_mask & (1u << lane)) t = __int_as_float(0x7fc00000);
    }
    return t;
}

}  // namespace glg
