// glg_race_packed.cuh - the production step kernel: TWO cars per warp (16 lanes each).
//
// Same algorithm and the same arithmetic as race_step_kernel<GLG_STEP_FAST> (glg_race.cu, glg_sensors.cuh);
// what changes is the mapping.  A warp owns a track's cars 2w and 2w+1 ("groups" of 16 lanes), so that
//   * the per-car scalar work (state loads, kinematics, reward, write-back, address arithmetic) is issued
//     once per two cars instead of once per car,
//   * the list phase (flagged walls -> candidate rays -> exact evaluation) runs 16 walls per car per pass:
//     ~45 flagged walls per car fill 3 passes of a half warp, each lane evaluating the first candidate ray
//     of its wall on the spot (only further rays of a wall are queued for a short second round),
//   * the vertex pass (stage 1) costs the same (17 half-warp passes for two cars = 8.5 per car).
// Rollouts chain consecutive launches car by car (glg_race_rollout); with keep_all the car state is handed
// from launch to launch in self-validating 64-bit words ("LL") instead of through the state arrays.
// Included by glg_race.cu after StepArgs.
#pragma once
#include "glg_common.cuh"
#include "glg_exact.cuh"
#include "glg_sensors.cuh"

namespace glg {

constexpr int PK_G = 16;                       // lanes per car
constexpr int PK_QCAP = 288;                   // (wall, ray) candidates per car kept before the car falls back to brute force
constexpr int PK_RAYS = 18;                    // stage 1 is written for 9 ray lines
#ifndef GLG_S1_UNROLL
#define GLG_S1_UNROLL 3
#endif
constexpr int PK_S1_UNROLL = GLG_S1_UNROLL;    // unroll factor of the stage-1 vertex loop

struct PackedCar {                             // per-car shared scratch
    float4 ray[PK_RAYS];                       // dx, dy, far x, far y
    int tmin[PK_RAYS + 2];                     // running min of t as ordered int bits
    unsigned nan_mask;
    int qn;                                    // rays queued for the second evaluation round
    unsigned pad[2];
};

// shared memory per track: [record 3N float2][mbarrier 16 B][cars x PackedCar][cars x cq u16[LL]][cars x wlist u16[2N]]
// LL = max(list_len(N), PK_QCAP): cq holds the collision walls first and the candidate queue later.  With one
// warp per track (P <= 2) the two wall lists (2 x 2N u16 = 8N bytes) live in the record's centre points, which
// nobody reads after the progress arg-min, and the last block is not allocated.
__host__ __device__ inline unsigned pk_list_len(int N) { const int a = list_len(N); return a > PK_QCAP ? a : PK_QCAP; }
__host__ __device__ inline unsigned pk_cars_offset(int N) { return (unsigned)smem_barrier_offset(N) + 16u; }
__host__ __device__ inline unsigned pk_cq_offset(int N, int cars) { return pk_cars_offset(N) + (unsigned)cars * (unsigned)sizeof(PackedCar); }
__host__ __device__ inline unsigned pk_wlist_offset(int N, int cars) { return pk_cq_offset(N, cars) + (unsigned)cars * pk_list_len(N) * 2u; }
__host__ __device__ inline unsigned pk_track_bytes(int N, int cars) {
    const unsigned x = cars == 2 ? pk_wlist_offset(N, cars) : pk_wlist_offset(N, cars) + (unsigned)cars * 2u * (unsigned)N * 2u;
    return (x + 127u) & ~127u;
}

// min over the 16 lanes of a group (xor butterflies stay inside the group); redux.sync with two different
// member masks in one warp takes the compiler's slow divergent path
__device__ __forceinline__ unsigned group_min_u32(unsigned v) {
#pragma unroll
    for (int off = PK_G / 2; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

__device__ __forceinline__ unsigned group_ballot(bool pred, int grp) {
    return (__ballot_sync(FULL, pred) >> (grp * PK_G)) & 0xffffu;
}

// brute-force sensors of one car by its 16 lanes (rare: a precondition of the pruning failed); every lane
// of the warp calls it, lanes of a group that does not need it just ride along
__device__ __noinline__ void packed_sensors_brute(const TrackView& tv, const glg_race_params& pr, P2 s, P2 nd,
                                                  int gl, unsigned gmask, bool wanted, PackedCar* car)
{
    for (int i = 0; i < PK_RAYS; ++i) {
        P2 d, f;
        ray_setup(pr, i, s, nd, d, f);
        float t = INF;
        bool nan = false;
        for (int w = gl; w < 2 * tv.N - 1; w += PK_G) {
            P2 p, q;
            wall_by_line_index(tv, w, p, q);
            const float tw = ray_wall_t(p, q, s, d, f);
            if (tw != tw) nan = true;
            else t = fminf(t, tw);
        }
        const unsigned anynan = __ballot_sync(FULL, nan) & gmask;
        const float m = __uint_as_float(group_min_u32(__float_as_uint(fmaxf(t, 0.f))));   // t in {-0} u [0, inf]
        const bool negzero = (__ballot_sync(FULL, __float_as_uint(t) == 0x80000000u) & gmask) != 0u;
        if (wanted && gl == 0) {
            float r = (m == 0.f && negzero) ? -0.f : m;
            car->tmin[i] = __float_as_int(r);
            if (anynan) car->nan_mask |= 1u << i;
        }
    }
    __syncwarp();
}

__device__ __noinline__ bool packed_collide_brute(const TrackView& tv, P2 op, P2 np, int gl, int grp)
{
    bool hit = false;
    for (int w = gl; w < 2 * tv.N - 1; w += PK_G) {
        P2 p, q;
        wall_by_line_index(tv, w, p, q);
        hit = hit || segments_cross(p, q, op, np);
    }
    return group_ballot(hit, grp) != 0u;
}

__host__ inline PackedLayout pk_layout(int N, int cars) {
    PackedLayout l;
    l.bar_off = (unsigned)smem_barrier_offset(N);
    l.cars_off = pk_cars_offset(N);
    l.cq_off = pk_cq_offset(N, cars);
    l.wlist_off = pk_wlist_offset(N, cars);
    l.list_len = pk_list_len(N);
    l.track_bytes = pk_track_bytes(N, cars);
    return l;
}

#ifndef GLG_PACKED_MINBLOCKS
#define GLG_PACKED_MINBLOCKS 8      // 128-thread blocks per SM the register allocator targets: 8 -> 64 registers for the
                                    // two-tracks-per-CTA shape (its residency is capped at 16 CTAs/SM anyway; 8.85e8 vs 8.6e8
                                    // env-steps/s with 48), 10 -> 48 registers for one track of 3..8 cars per CTA
#endif

// TPB tracks per CTA; each track has WPT = ceil(P/2) consecutive warps.
template <int TPB>
__global__ void __launch_bounds__(128, TPB == 2 ? GLG_PACKED_MINBLOCKS : GLG_PACKED_MINBLOCKS + 2)
race_step_packed_kernel(const __grid_constant__ glg_race_params pr, const StepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int step_no = a.step_no, seq = a.seq;
    if (a.base) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        step_no += __ldcg(a.base);
        seq += __ldcg(a.base + 1);
        if (step_no > __ldcg(a.base + 2)) return;     // past the time limit: the step is a no-op
    }
    GLG_MARK_INIT;
    constexpr int O = PK_RAYS;
    const int N = a.N, B = a.B, V = 2 * N;
    const int P = pr.num_players;
    // TPB == 2 is launched for P <= 2 only: one warp per track, two tracks per CTA
    const int WPT = (TPB == 2) ? 1 : (P + 1) >> 1;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int tslot = (TPB == 2) ? warp : 0;
    const int wt = (TPB == 2) ? 0 : warp;
    const int b = blockIdx.x * TPB + tslot;
    const int grp = lane >> 4, gl = lane & 15;
    const unsigned gmask = 0xffffu << (grp * PK_G);
    const unsigned lt = (1u << gl) - 1u;
    const int p = wt * 2 + grp;
    const bool track_on = b < B;
    const bool car_on = track_on && p < P;
    const int cars = 2 * WPT;

    const int ci = wt * 2 + grp;                                          // car slot within the track
    // shared-memory carve-up (byte offsets computed by the host, pk_layout)
    unsigned char* tbase = smem_raw + (unsigned)tslot * a.pk.track_bytes;
    float2* pts = reinterpret_cast<float2*>(tbase);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tbase + a.pk.bar_off);
    PackedCar* car = reinterpret_cast<PackedCar*>(tbase + a.pk.cars_off) + ci;
    unsigned short* cq = reinterpret_cast<unsigned short*>(tbase + a.pk.cq_off) + (unsigned)ci * a.pk.list_len;
    unsigned short* wlist = (TPB == 2)
        ? reinterpret_cast<unsigned short*>(pts + 2 * N) + (unsigned)grp * 2u * (unsigned)N     // over the centre points
        : reinterpret_cast<unsigned short*>(tbase + a.pk.wlist_off) + (unsigned)ci * 2u * (unsigned)N;

    // ---- stage the track record with one bulk async copy per track ----
    const uint32_t rec_bytes = (uint32_t)(3 * N * sizeof(float2));
    if (track_on && wt == 0 && lane == 0)
        record_copy_async(pts, reinterpret_cast<const float2*>(a.geom) + (size_t)b * 3 * N, rec_bytes, bar);
    asm volatile("griddepcontrol.launch_dependents;");
    const int k = b * P + p;
    // Inputs of a chained launch (t >= 1 of a rollout) are older than the rollout itself: fetch them before waiting
    // for the previous step of the car, so that their latency is off the car's critical path.
    const bool inputs_early = a.chained != 0;
    bool ok = false;
    int act = 0;
    float2 ext = make_float2(INF, INF);
    if (inputs_early && car_on) {
        ok = a.valid[b] != 0;
        act = (int)a.actions[(size_t)p * B + b];
        ext = __ldg(reinterpret_cast<const float2*>(a.extent) + b);
    }
    // "LL" chaining (rollouts that keep every step's outputs): the previous step of this car hands its state
    // over in six 64-bit words {launch number : value} - single-copy atomic, self-validating, so neither side
    // needs a fence and the consumer needs no second round trip for the state itself.
    unsigned llw = 0;
    if (a.ll_read) {
        if (car_on && gl < 6) {
            const unsigned want = (unsigned)(seq - 1);
            const unsigned long long* w = a.ll + (size_t)k * 6 + gl;
            unsigned long long x;
            int spin = 0;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(x) : "l"(w) : "memory");
                if ((unsigned)(x >> 32) == want) break;
                __nanosleep(GLG_CHAIN_BACKOFF_NS);
                if (++spin > (1 << 22)) __trap();
            } while (true);
            llw = (unsigned)x;
        }
        __syncwarp();
    } else if (a.chained) {
        if (car_on && gl == 0) {
            const int want = seq - 1;
            int got, spin = 0;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(a.chain + k) : "memory");
                if (got == want) break;
                __nanosleep(GLG_CHAIN_BACKOFF_NS);                         // do not burn issue slots while waiting
                if (++spin > (1 << 22)) __trap();
            } while (true);
        }
        __syncwarp();
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    GLG_MARK(0);
    // ---- car state and kinematics (uniform within a group) ----
    bool alive = false, fin = false;
    float2 dir = make_float2(0.f, 1.f), pos = make_float2(0.f, 0.f);
    float spd = 0.f;
    if (a.ll_read) {                     // words: pos.x, pos.y, dir.x, dir.y, speed, flags (alive | finished << 1)
        pos.x = __uint_as_float(__shfl_sync(FULL, llw, 0, PK_G));
        pos.y = __uint_as_float(__shfl_sync(FULL, llw, 1, PK_G));
        dir.x = __uint_as_float(__shfl_sync(FULL, llw, 2, PK_G));
        dir.y = __uint_as_float(__shfl_sync(FULL, llw, 3, PK_G));
        spd = __uint_as_float(__shfl_sync(FULL, llw, 4, PK_G));
        const unsigned fl = __shfl_sync(FULL, llw, 5, PK_G);
        if (car_on) {
            alive = (fl & 1u) != 0u;
            fin = (fl & 2u) != 0u;
        }
    }
    if (car_on) {
        if (!a.ll_read) {
            alive = __ldcg(&a.st.alive[k]) != 0;
            fin = __ldcg(&a.st.finishes[k]) != 0;
            dir = __ldcg(reinterpret_cast<const float2*>(a.st.directions) + k);
            pos = __ldcg(reinterpret_cast<const float2*>(a.st.positions) + k);
            spd = __ldcg(&a.st.speeds[k]);
        }
        if (!inputs_early) {
            ok = a.valid[b] != 0;
            act = (int)a.actions[(size_t)p * B + b];
            ext = __ldg(reinterpret_cast<const float2*>(a.extent) + b);
        }
    }
    const int pc = min(p, GLG_MAX_PLAYERS - 1);
    act = min(max(act, 0), 8);
    if (!alive || !ok) act = 0;                                           // race.py:359
    const int fs = act / 3, ft = act - 3 * fs;
    const float c = pr.turn_cos[pc][fs], s_ = pr.turn_sin[pc][fs];
    const P2 nd{xadd(xmul(dir.x, c), xmul(dir.y, s_)),                    // race.py:362-364
                xadd(xmul(dir.x, -s_), xmul(dir.y, c))};
    const float v = xadd(spd, pr.speed_inc[pc][ft]);                      // race.py:367
    float nv = fminf(pr.vmax[pc], fmaxf(v, 0.f));                         // race.py:369
    const bool moving = fabsf(nv) > 1e-7f;                                // race.py:370
    const P2 op{pos.x, pos.y};
    const P2 np{xadd(pos.x, xmul(nd.x, nv)), xadd(pos.y, xmul(nd.y, nv))};   // race.py:372

    GLG_MARK(1);
    if (TPB == 2) __syncwarp();          // one warp per track: the lane that initialised the mbarrier is in this warp
    else __syncthreads();                // mbarrier init visible to the other warps of the track
    if (track_on) record_copy_wait(bar);
    const TrackView tv{pts, pts + 2 * N, N};

    GLG_MARK(2);
    // ---- progress: FIRST arg-min of |np - centre_j| (race.py:374-376), see race_step_kernel ----
    int idx;
    {
        float q1 = INF, q2 = INF;
        int j1 = 0x7fffffff;
        {
            const float2* cp = tv.centre + gl;
            const int full = N / PK_G;                         // iterations in which every lane has a point
            int j = gl;
#pragma unroll 3
            for (int it = 0; it < full; ++it, cp += PK_G, j += PK_G) {
                const float2 cpt = *cp;
                const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                const bool less = q < q1;
                q2 = less ? q1 : fminf(q2, q);
                j1 = less ? j : j1;
                q1 = less ? q : q1;
            }
            if (j < N) {                                       // tail
                const float2 cpt = *cp;
                const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                const bool less = q < q1;
                q2 = less ? q1 : fminf(q2, q);
                j1 = less ? j : j1;
                q1 = less ? q : q1;
            }
        }
        const float qmin = __uint_as_float(group_min_u32(__float_as_uint(q1)));              // q >= 0
        const float qcut = qmin * 1.000001f + 1e-45f;
        idx = (int)group_min_u32((q1 == qmin) ? (unsigned)j1 : 0x7fffffffu);
        // near ties of the ROUNDED norms: only if a second value lies within qcut (rare) redo the pass with sqrt_rn
        const unsigned n1 = __ballot_sync(FULL, q1 <= qcut) & gmask, n2 = __ballot_sync(FULL, q2 <= qcut) & gmask;
        const bool tie = __popc(n1) + __popc(n2) > 1;
        if (__any_sync(FULL, tie)) {
            const float smin = __fsqrt_rn(qmin);
            int first = 0x7fffffff;
            for (int j = gl; j < N && tie; j += PK_G) {
                const float2 cpt = tv.centre[j];
                const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                if (q <= qcut && __fsqrt_rn(q) == smin) { first = j; break; }
            }
            const int f2 = (int)group_min_u32((unsigned)first);
            if (tie) idx = f2;
        }
    }

    GLG_MARK(3);
    __syncwarp();                        // (P <= 2) the centre points become the wall lists from here on

    // ---- scan: preconditions, ray table, stage 1 over all vertices ----
    float reward = fin ? 0.f : pr.step_penalty;                            // race.py:382-383
    const bool upd = alive && moving && ok;                                // race.py:380
    const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
    const bool safe = d2 > 0.5f && d2 < 2.f && fabsf(np.x) + fabsf(np.y) + ext.x < 200.f;   // see scan_two_stage
    const bool scan_on = alive && safe;
    if (gl < 2) car->tmin[PK_RAYS + gl] = 0;
    {                                    // 18 rays, 16 lanes: lanes 0..15 take ray gl, lanes 0 and 1 also rays 16 and 17
        P2 d, f;
        ray_setup(pr, gl, np, nd, d, f);
        car->ray[gl] = make_float4(d.x, d.y, f.x, f.y);
        car->tmin[gl] = 0x7f800000;
        if (gl < O - PK_G) {
            ray_setup(pr, PK_G + gl, np, nd, d, f);
            car->ray[PK_G + gl] = make_float4(d.x, d.y, f.x, f.y);
            car->tmin[PK_G + gl] = 0x7f800000;
        }
    }
    if (gl == 0) { car->nan_mask = 0; car->qn = 0; }

    const float Lmax = ext.y;
    const float Rc = fmaf(RC_FACTOR, Lmax, 1e-3f);
    const float close2 = scan_on ? Rc * Rc * d2 * 1.0001f : 0.f;           // scan off: nothing is flagged
    const float colR = fabsf(op.x - np.x) + fabsf(op.y - np.y) + Lmax + 1e-3f;
    const bool col_on = upd && safe;
    const float col2 = col_on ? colR * colR * d2 * 1.0001f : 0.f;
    const float Kn = scan_on ? 9.f * EPS_PERP * 1.4143f * 1.001f : 0.f;
    const int passes = (V + PK_G - 1) / PK_G;                              // <= 32 (the host routes N > 256 elsewhere)
    unsigned sbits = 0, fbits = 0, cbits = 0;
    GLG_MARK(4);
    {
        const float2* vp = tv.line + gl;
        const int quads = passes >> 2;
        for (int it = 0; it < quads; ++it, vp += 4 * PK_G) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 pt = vp[u * PK_G];
                const float ux = pt.x - np.x, uy = pt.y - np.y;
                const float za = fmaf(ux, nd.x, uy * nd.y);
                const float zb = fmaf(ux, nd.y, -(uy * nd.x));
                const float a2 = za * za, b2 = zb * zb;
                const float r2z = a2 + b2;
                const float re3 = za * fmaf(-3.f, b2, a2);
                const float im3 = zb * fmaf(3.f, a2, -b2);
                const float im9 = im3 * fmaf(3.f, re3 * re3, -(im3 * im3));
                const float r4 = r2z * r2z;
                const float near = fmaf(r4 * r4, -Kn, fabsf(im9));
                const float flag = fminf(near, r2z - close2);
                sbits = __funnelshift_l(__float_as_uint(im9), sbits, 1);
                fbits = __funnelshift_l(__float_as_uint(flag), fbits, 1);
                cbits = __funnelshift_l(__float_as_uint(r2z - col2), cbits, 1);
            }
        }
        for (int it = passes & 3; it > 0; --it, vp += PK_G) {
            const float2 pt = *vp;
            const float ux = pt.x - np.x, uy = pt.y - np.y;
            const float za = fmaf(ux, nd.x, uy * nd.y);
            const float zb = fmaf(ux, nd.y, -(uy * nd.x));
            const float a2 = za * za, b2 = zb * zb;
            const float r2z = a2 + b2;
            const float re3 = za * fmaf(-3.f, b2, a2);
            const float im3 = zb * fmaf(3.f, a2, -b2);
            const float im9 = im3 * fmaf(3.f, re3 * re3, -(im3 * im3));
            const float r4 = r2z * r2z;
            const float near = fmaf(r4 * r4, -Kn, fabsf(im9));
            const float flag = fminf(near, r2z - close2);
            sbits = __funnelshift_l(__float_as_uint(im9), sbits, 1);
            fbits = __funnelshift_l(__float_as_uint(flag), fbits, 1);
            cbits = __funnelshift_l(__float_as_uint(r2z - col2), cbits, 1);
        }
    }
    {
        const int sh = 32 - passes;
        sbits = __brev(sbits) >> sh;
        fbits = __brev(fbits) >> sh;
        cbits = __brev(cbits) >> sh;
    }
    GLG_MARK(5);
    if (!scan_on) { fbits = 0; sbits = 0; }                                // (Kn = close2 = 0 can still flag |im9| < 0: never; belt and braces)
    if (!col_on) cbits = 0;
    unsigned s1 = __shfl_down_sync(FULL, sbits, 1, PK_G), f1 = __shfl_down_sync(FULL, fbits, 1, PK_G);
    const unsigned s0 = __shfl_sync(FULL, sbits, 0, PK_G), f0 = __shfl_sync(FULL, fbits, 0, PK_G);
    if (gl == PK_G - 1) { s1 = s0 >> 1; f1 = f0 >> 1; }
    const int nown = (V - 2 - gl >= 0) ? ((V - 2 - gl) >> 4) + 1 : 0;      // walls w = 16*pass + gl <= V-2
    unsigned own = nown >= 32 ? FULL : ((1u << nown) - 1u);
    if (((N - 1) & 15) == gl) own &= ~(1u << ((N - 1) >> 4));              // the start line is appended below
    unsigned wbits = scan_on ? (((sbits ^ s1) | fbits | f1) & own) : 0u;
    cbits &= own;
    int nw, nc = 0;
    {
        const int cnt = __popc(wbits);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < PK_G; off <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, off, PK_G);
            if (gl >= off) incl += t;
        }
        nw = __shfl_sync(FULL, incl, PK_G - 1, PK_G);
        int posn = incl - cnt;
#pragma unroll
        for (int r = 0; r < 3; ++r) {                                      // straight-line for the usual <= 3 walls per lane
            if (wbits) {
                const int pass = __ffs(wbits) - 1;
                wbits &= wbits - 1;
                wlist[posn++] = (unsigned short)(pass * PK_G + gl);
            }
        }
        while (wbits) {
            const int pass = __ffs(wbits) - 1;
            wbits &= wbits - 1;
            wlist[posn++] = (unsigned short)(pass * PK_G + gl);
        }
        if (scan_on) {
            if (gl == 0) wlist[nw] = (unsigned short)(N - 1);
            ++nw;
        }
    }
    {
        unsigned live = __ballot_sync(FULL, cbits != 0u);
        while (live) {                                                     // uniform over the warp; usually 0 or 1 rounds
            const unsigned mine = (live >> (grp * PK_G)) & 0xffffu;
            if (cbits) {
                const int pass = __ffs(cbits) - 1;
                cbits &= cbits - 1;
                cq[nc + __popc(mine & lt)] = (unsigned short)(pass * PK_G + gl);
            }
            nc += __popc(mine);
            live = __ballot_sync(FULL, cbits != 0u);
        }
        if (col_on) {
            if (gl == 0) cq[nc] = (unsigned short)(N - 1);
            ++nc;
        }
    }
    __syncwarp();

    GLG_MARK(6);
    // ---- collision: exact test of the walls near the path (race.py:406) ----
    bool wall_hit = false;
    {
        const float ox = op.x - np.x, oy = op.y - np.y;
        const float bx0 = fminf(ox, 0.f) - BOX_MARGIN, bx1 = fmaxf(ox, 0.f) + BOX_MARGIN;
        const float by0 = fminf(oy, 0.f) - BOX_MARGIN, by1 = fmaxf(oy, 0.f) + BOX_MARGIN;
        bool hit = false;
        for (int e = gl; e < nc; e += PK_G) {
            const int w = cq[e];
            const float2 p0 = tv.line[w], p1 = tv.line[w + 1];
            const float ux = p0.x - np.x, uy = p0.y - np.y, ux1 = p1.x - np.x, uy1 = p1.y - np.y;
            if (!(fmaxf(ux, ux1) < bx0 || fminf(ux, ux1) > bx1 || fmaxf(uy, uy1) < by0 || fminf(uy, uy1) > by1)) {
                P2 pp, qq;
                wall_by_line_index(tv, w, pp, qq);
                hit = hit || segments_cross(pp, qq, op, np);
            }
        }
        wall_hit = group_ballot(hit, grp) != 0u;
    }
    const bool brute_col = upd && !safe;
    if (__any_sync(FULL, brute_col)) {
        const bool h = packed_collide_brute(tv, op, np, gl, grp);
        if (brute_col) wall_hit = h;
    }
    __syncwarp();                        // cq is reused as the candidate queue from here on

    GLG_MARK(7);
    // ---- finish line, reward, score, state write-back (race.py:431-456) ----
    if (upd) {
        const bool dead = wall_hit;
        const float2 fl = tv.line[2 * N - 1], fr = tv.line[0];             // finish: left[N-1] -> right[N-1], race.py:169
        bool done = false;
        {
            const float x0 = fminf(op.x, np.x) - BOX_MARGIN, x1 = fmaxf(op.x, np.x) + BOX_MARGIN;
            const float y0 = fminf(op.y, np.y) - BOX_MARGIN, y1 = fmaxf(op.y, np.y) + BOX_MARGIN;
            const bool apart = fmaxf(fl.x, fr.x) < x0 || fminf(fl.x, fr.x) > x1 ||
                               fmaxf(fl.y, fr.y) < y0 || fminf(fl.y, fr.y) > y1;
            if (!apart) done = segments_cross(P2{fl.x, fl.y}, P2{fr.x, fr.y}, op, np);   // race.py:431-432
        }
        reward = xadd(reward, xsub(done ? 1.f : 0.f, dead ? 1.f : 0.f));   // race.py:434
        alive = alive && !dead && !done;                                   // race.py:414, 435
        fin = fin || done;                                                 // race.py:436
        if (gl == 0 && (dead || done))                                     // race.py:442-447 (done wins over dead)
            a.st.scores[k] = done ? step_no : idx + pr.steps_limit + 1;
    }
    if (!alive) nv = 0.f;                                                  // race.py:449
    const float drag = xsub(1.f, xmul(xsub(1.f, ft != 0 ? 1.f : 0.f), pr.drag));   // race.py:452
    const float speed = xmul(nv, drag);                                    // race.py:455
    if (a.ll_write && car_on && gl < 6) {
        const unsigned word = gl == 0 ? __float_as_uint(np.x) : gl == 1 ? __float_as_uint(np.y)
                            : gl == 2 ? __float_as_uint(nd.x) : gl == 3 ? __float_as_uint(nd.y)
                            : gl == 4 ? __float_as_uint(speed) : ((alive ? 1u : 0u) | (fin ? 2u : 0u));
        const unsigned long long x = ((unsigned long long)(unsigned)seq << 32) | word;
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(a.ll + (size_t)k * 6 + gl), "l"(x) : "memory");
    }
    if (car_on && gl == 0) {
        if (a.arrays_write) {            // (intermediate launches of an LL rollout hand the state over in the LL words only)
            reinterpret_cast<float2*>(a.st.directions)[k] = make_float2(nd.x, nd.y);
            reinterpret_cast<float2*>(a.st.positions)[k] = make_float2(np.x, np.y);
            a.st.speeds[k] = speed;
            a.st.alive[k] = alive ? 1 : 0;
            a.st.finishes[k] = fin ? 1 : 0;
        }
        a.rewards_out[(size_t)p * B + b] = reward;
        if (a.history && b == a.record_id) {                               // race.py:492-494
            float* h = a.history + ((size_t)step_no * P + p) * 6;
            h[0] = np.x; h[1] = np.y; h[2] = nd.x; h[3] = nd.y; h[4] = (float)act; h[5] = alive ? 1.f : 0.f;
        }
        if (a.chain && a.early && !a.ll_write)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.chain + k), "r"(seq) : "memory");
        if (alive && a.alive_stamp) atomicMax(&a.alive_stamp[b % GLG_ALIVE_SLOTS], seq);   // (after the release: not waited for)
    }

    GLG_MARK(8);
    // ---- stage 2: candidate rays of the flagged walls; the first ray of a wall is evaluated on the spot ----
    // (race.py:287-308; most flagged walls have exactly one candidate ray: the lane already holds the wall, so the
    //  exact test runs right here and only the 2nd, 3rd, ... rays of a wall go through the queue)
    const bool sense = alive && scan_on;                       // `alive` is post-update here: dead cars report zeros
    {
        const unsigned all_rays = (1u << O) - 1u;
        const float sect = (float)O * (0.5f / PI_F);
        const float m_eta = ETA_ANGLE * sect, m_eps = EPS_PERP * sect, fhalf = 0.5f * (float)O;
        const int nws = sense ? nw : 0;
        const int nwmax = (int)__reduce_max_sync(FULL, (unsigned)nws);
        for (int base = gl; base < nwmax; base += PK_G) {
            if (base < nws) {
                const int w = wlist[base];
                const float2 p0 = tv.line[w], p1 = tv.line[w + 1];
                unsigned mask = wall_ray_mask(p0.x - np.x, p0.y - np.y, p1.x - np.x, p1.y - np.y, nd, O, sect, fhalf, m_eps, m_eta, all_rays);
                if (mask) {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const bool rev = w < N;                            // wall_by_line_index
                    const P2 pp = rev ? P2{p1.x, p1.y} : P2{p0.x, p0.y}, qq = rev ? P2{p0.x, p0.y} : P2{p1.x, p1.y};
                    const float4 r = car->ray[i];
                    const float t = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
                    if (t != t) atomicOr(&car->nan_mask, 1u << i);
                    else atomicMin(&car->tmin[i], __float_as_int(t));
                    if (mask) {                                        // further rays of this wall: second round
                        int posn = atomicAdd(&car->qn, __popc(mask));
                        const int wcode = w << 5;
                        while (mask && posn < PK_QCAP) {
                            cq[posn++] = (unsigned short)(wcode | (__ffs(mask) - 1));
                            mask &= mask - 1;
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    const int total = car->qn;                                         // > PK_QCAP: some rays were dropped -> brute force
    const bool overflow = total > PK_QCAP;

    GLG_MARK(9);
    // ---- second round: the queued rays ----
    if (sense && !overflow) {
        for (int e = gl; e < total; e += PK_G) {
            const int code = cq[e];
            const int w = code >> 5, i = code & 31;
            P2 pp, qq;
            wall_by_line_index(tv, w, pp, qq);
            const float4 r = car->ray[i];
            const float t = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
            if (t != t) atomicOr(&car->nan_mask, 1u << i);
            else atomicMin(&car->tmin[i], __float_as_int(t));
        }
    }
    const bool brute_s = alive && (!safe || overflow);
    if (__any_sync(FULL, brute_s)) packed_sensors_brute(tv, pr, np, nd, gl, gmask, brute_s, car);
    __syncwarp();

    GLG_MARK(10);
    // ---- observation pack [P,B,O+2] (race.py:496-500): 16 lanes write 20 values in two rounds ----
    if (car_on) {
        float* out = a.states_out + ((size_t)p * B + b) * (O + 2);
        const unsigned nanm = car->nan_mask;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = r * PK_G + gl;
            if (i < O + 2) {
                float num = 0.f, den = 1.f;
                if (i < O) {
                    if (alive) {
                        float t = __int_as_float(car->tmin[i]);
                        if (nanm & (1u << i)) t = __int_as_float(0x7fc00000);
                        num = (t != t) ? t : fminf(t, pr.max_distance);
                        den = pr.max_distance;
                    }
                } else if (i == O) { num = speed; den = pr.vmax[pc]; }
                else { num = (float)idx; den = pr.progress_div; }
                out[i] = xdiv(num, den);
            }
        }
    }
    GLG_MARK(11);
    if (a.chain && !a.early && !a.ll_write) {   // publish "this car's step `seq` is complete" (all lanes' stores first)
        __syncwarp();
        if (car_on && gl == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.chain + k), "r"(seq) : "memory");
        }
    }
    // A chained launch never waited for the previous grid as a whole.  Waiting for it here, after the work, costs
    // nothing in steady state (the previous step's CTAs are done by now) and makes launches COMPLETE in order, so
    // that whatever follows the rollout in the stream finds every step's outputs and scores in memory.
    if (a.chained) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace glg
