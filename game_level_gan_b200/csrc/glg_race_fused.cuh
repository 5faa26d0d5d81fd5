// glg_race_fused.cuh - the persistent rollout kernel: ONE launch plays T steps (glg_race_rollout, GLG_ROLLOUT_FUSED).
//
// Same algorithm and the same reported arithmetic as race_step_packed_kernel (two cars per warp, 16 lanes each;
// P <= 2: one warp per track and two tracks per CTA, else one track per CTA with ceil(P/2) warps), but a warp keeps its
// track and its cars for the whole rollout:
//   * the track record is staged in shared memory ONCE (bulk async copies) and stays there - the polyline as
//     structure of arrays (xs[], ys[]), so that stage 1 handles TWO vertices per lane with packed fp32
//     instructions (FFMA2 / FMUL2 / FADD2: half the issue slots of the scalar form, and the kernel is issue-bound),
//   * the car state lives in registers from step to step; the state arrays are read once and written once,
//   * per step the only global traffic is the car's action (8 B, fetched one step ahead), its observation
//     (80 B) and its reward (4 B) - 92 B per car-step instead of 3120 B per track + 145 B per car,
//   * the progress arg-min looks at a window of 16 centre points around the previous arg-min; a certificate from
//     the last full pass (min distance of every OTHER point, minus the distance travelled since) proves that no
//     point outside the window can win or tie; when it fails the full pass runs and re-centres the window,
//   * the walls the path could touch are a subset of the flagged walls when the step is shorter than
//     2.2 Lmax (always, at racing speeds), so the collision test rides along in the flagged-wall loop,
//   * no launch, no hand-over between launches, no grid-wide dependency: cars never interact.
// Results are bit-identical to T calls of glg_race_step (tests/test_race_gpu.py, tests/test_fused_rollout_gpu.py).
// Included by glg_race.cu after glg_race_packed.cuh.
#pragma once
#include "glg_race_packed.cuh"

namespace glg {

// shared memory per track: [centre N float2][xs VP f32][ys VP f32][mbarrier 16 B][cars x PackedCar]
//                          [cars x cq u16[LL]][cars x wlist u16[2N]]; the last three blocks double as the landing
//                          zone of the polyline (array of structures, 16N bytes) before it is transposed
struct FusedLayout {
    unsigned centre_off, xs_off, ys_off, bar_off, cars_off, cq_off, wlist_off, list_len, track_bytes;
    int VP;                      // vertices rounded up to whole stage-1 passes (32 per group and pass)
};

__host__ __device__ constexpr FusedLayout fused_layout(int N, int cars) {
    FusedLayout l{};
    const unsigned V = 2u * (unsigned)N;
    l.VP = (int)((V + 31u) & ~31u);
    l.centre_off = 0;
    l.xs_off = ((unsigned)N * 8u + 15u) & ~15u;
    l.ys_off = l.xs_off + (unsigned)l.VP * 4u;
    l.bar_off = l.ys_off + (unsigned)l.VP * 4u;
    l.cars_off = l.bar_off + 16u;
    const unsigned ll0 = ((V + 31u) / 32u) * 32u;                          // list_len(N)
    l.list_len = ll0 > (unsigned)PK_QCAP ? ll0 : (unsigned)PK_QCAP;       // pk_list_len(N)
    l.cq_off = l.cars_off + (unsigned)cars * (unsigned)sizeof(PackedCar);
    l.wlist_off = l.cq_off + (unsigned)cars * l.list_len * 2u;
    unsigned end = l.wlist_off + (unsigned)cars * V * 2u;
    const unsigned landing = l.cars_off + V * 8u;                         // the polyline lands at cars_off
    if (landing > end) end = landing;
    l.track_bytes = (end + 127u) & ~127u;
    return l;
}

struct FusedArgs {
    const float* geom;
    const int64_t* actions;      // [T,P,B]
    const uint8_t* valid;
    const float* extent;
    glg_race_state st;
    float* states_out;           // [T,P,B,O+2] (keep_all) or [P,B,O+2]
    float* rewards_out;          // [T,P,B] or [P,B]
    int32_t* alive_stamp;
    float* history;
    int32_t B, N, T, first_step_no, record_id, last_seq, keep_all;
    FusedLayout lay;             // fused_layout(N, cars), for the instantiations that do not know N at compile time
};

// ---- packed fp32 (two values per lane and instruction) ---------------------------------------------------
// NOTE: ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into one FFMA2 even under --fmad=false (unlike the scalar
// forms).  These helpers are therefore used ONLY for the approximate stage-1 predicates, whose margins cover any
// rounding; everything that reaches an output goes through the scalar round-to-nearest primitives of glg_exact.cuh.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return bits_f2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// ray_setup (glg_sensors.cuh) with the table entry already in registers
__device__ __forceinline__ void ray_setup_cs(float rc, float rs, P2 s, P2 nd, P2& d, P2& f) {
    d = P2{xadd(xmul(nd.x, rc), xmul(nd.y, rs)), xadd(xmul(nd.x, -rs), xmul(nd.y, rc))};
    f = P2{xadd(s.x, xmul(1000.f, d.x)), xadd(s.y, xmul(1000.f, d.y))};
}

// wall w of the polyline in the reference's orientation, from the structure-of-arrays copy (see wall_by_line_index)
__device__ __forceinline__ void wall_soa(const float* xs, const float* ys, int N, int w, P2& p, P2& q) {
    const float x0 = xs[w], x1 = xs[w + 1], y0 = ys[w], y1 = ys[w + 1];
    const bool rev = w < N;
    p = rev ? P2{x1, y1} : P2{x0, y0};
    q = rev ? P2{x0, y0} : P2{x1, y1};
}

// one evaluated (wall, ray) pair into the car's per-ray minima: +inf (no hit, the common case) changes nothing and is
// not sent; NaN is recorded in the mask (torch.min propagates it)
__device__ __forceinline__ void fused_report(PackedCar* car, int i, float tw) {
    if (tw != tw) atomicOr(&car->nan_mask, 1u << i);
    else if (tw < INF) atomicMin(&car->tmin[i], __float_as_int(tw));
}

// brute-force sensors of one car by its 16 lanes (rare: a precondition of the pruning failed), SoA polyline
__device__ __noinline__ void fused_sensors_brute(const float* xs, const float* ys, int N, const glg_race_params& pr, P2 s, P2 nd,
                                                 int gl, unsigned gmask, bool wanted, PackedCar* car)
{
    for (int i = 0; i < PK_RAYS; ++i) {
        P2 d, f;
        ray_setup(pr, i, s, nd, d, f);
        float t = INF;
        bool nan = false;
        for (int w = gl; w < 2 * N - 1; w += PK_G) {
            P2 p, q;
            wall_soa(xs, ys, N, w, p, q);
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

// every wall against the path (rare: pruning preconditions failed, or a step longer than 2.2 Lmax)
__device__ __noinline__ bool fused_collide_all(const float* xs, const float* ys, int N, P2 op, P2 np, int gl, int grp, bool boxes)
{
    const float bx0 = fminf(op.x, np.x) - BOX_MARGIN, bx1 = fmaxf(op.x, np.x) + BOX_MARGIN;
    const float by0 = fminf(op.y, np.y) - BOX_MARGIN, by1 = fmaxf(op.y, np.y) + BOX_MARGIN;
    bool hit = false;
    for (int w = gl; w < 2 * N - 1; w += PK_G) {
        const float x0 = xs[w], x1 = xs[w + 1], y0 = ys[w], y1 = ys[w + 1];
        if (boxes && (fmaxf(x0, x1) < bx0 || fminf(x0, x1) > bx1 || fmaxf(y0, y1) < by0 || fminf(y0, y1) > by1)) continue;
        P2 p, q;
        wall_soa(xs, ys, N, w, p, q);
        hit = hit || segments_cross(p, q, op, np);
    }
    return group_ballot(hit, grp) != 0u;
}

// Residency.  P <= 2 (64-thread CTAs, two tracks each): config 2 is 2048 CTAs on 148 SMs, and a persistent kernel must
// hold them all at once (a second wave would double the time): 14 CTAs per SM = 2072 slots, which leaves 72 registers
// per thread.  P > 2 (one track per CTA, up to 128 threads): 8 CTAs per SM -> 64 registers.
#ifndef GLG_FUSED_MINBLOCKS2
#define GLG_FUSED_MINBLOCKS2 14
#endif
#ifndef GLG_FUSED_MINBLOCKS1
#define GLG_FUSED_MINBLOCKS1 8
#endif
#ifndef GLG_F_ACT_SMEM
#define GLG_F_ACT_SMEM 1          // actions fetched 16 steps at a time by cp.async (0: one __ldg per step, a step ahead)
#endif
#ifndef GLG_F_S1_UNROLL
#define GLG_F_S1_UNROLL 3         // stage-1 passes per loop trip (9 passes at N = 130)
#endif
#ifndef GLG_F_RAYTAB
#define GLG_F_RAYTAB 1            // ray table in shared memory (0: per-lane indexed loads from the parameter bank)
#endif
constexpr int S1_UNROLL = GLG_F_S1_UNROLL;
constexpr int FW = 16;                         // centre points in the arg-min window (one per lane of a group)
constexpr int FW_BACK = 5;                     // of which behind the last arg-min (cars mostly advance)

// NC / CARS: track length N = L+2 and car slots per track known at compile time (0 = taken from the arguments).  The
// instantiations for the reference's own track length (RaceConfig.max_segments = 128 -> N = 130) have every shared-memory
// offset, trip count and ownership mask as literals.
template <int TPB, int NC, int CARS>
__global__ void __launch_bounds__(TPB == 2 ? 64 : 128, TPB == 2 ? GLG_FUSED_MINBLOCKS2 : GLG_FUSED_MINBLOCKS1)
race_rollout_fused_kernel(const __grid_constant__ glg_race_params pr, const FusedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int O = PK_RAYS;
    const int N = NC ? NC : a.N, B = a.B, V = 2 * N;
    const FusedLayout lay = (NC && CARS) ? fused_layout(NC, CARS) : a.lay;
    const int P = pr.num_players;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int tslot = (TPB == 2) ? warp : 0;
    const int wt = (TPB == 2) ? 0 : warp;
    const int b = blockIdx.x * TPB + tslot;
    const int grp = lane >> 4, gl = lane & 15;
    const unsigned gmask = 0xffffu << (grp * PK_G);
    const int p = wt * 2 + grp;
    const bool track_on = b < B;
    const bool car_on = track_on && p < P;
    const int ci = wt * 2 + grp;

    // per-warp copy of the ray table (cos, sin of the 18 ray angles): per-lane indexed loads from the parameter
    // (constant) bank serialise, shared memory does not; and the cars' actions, fetched 16 steps at a time by
    // asynchronous copies (one step per lane) one block ahead - no register holds a prefetched value across a step
    __shared__ float2 s_raytab[4][PK_RAYS];
    __shared__ long long s_acts[8][2][PK_G];
    if (lane < PK_RAYS) s_raytab[warp][lane] = make_float2(pr.ray_cos[lane], pr.ray_sin[lane]);
    long long (*acts)[PK_G] = s_acts[threadIdx.x >> 4];

    unsigned char* tbase = smem_raw + (unsigned)tslot * lay.track_bytes;
    const float2* centre = reinterpret_cast<const float2*>(tbase + lay.centre_off);
    float* xs = reinterpret_cast<float*>(tbase + lay.xs_off);
    float* ys = reinterpret_cast<float*>(tbase + lay.ys_off);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tbase + lay.bar_off);
    PackedCar* car = reinterpret_cast<PackedCar*>(tbase + lay.cars_off) + ci;
    unsigned short* cq = reinterpret_cast<unsigned short*>(tbase + lay.cq_off) + (unsigned)ci * lay.list_len;
    unsigned short* wlist = reinterpret_cast<unsigned short*>(tbase + lay.wlist_off) + (unsigned)ci * 2u * (unsigned)N;

    // everything this kernel reads may have been written by the previous kernel of the stream (reset, state restore)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ---- stage the record: polyline (16N bytes) into the landing zone, centre points (8N bytes) into place ----
    if (track_on && wt == 0 && lane == 0) {
        const float2* rec = reinterpret_cast<const float2*>(a.geom) + (size_t)b * 3 * N;
        const uint32_t mb = smem_u32(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(24 * N)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(tbase + lay.cars_off)), "l"(rec), "r"((uint32_t)(16 * N)), "r"(mb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(tbase + lay.centre_off)), "l"(rec + 2 * N), "r"((uint32_t)(8 * N)), "r"(mb) : "memory");
    }
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- car state: read once ----
    const int k = b * P + p;
    const size_t PB = (size_t)P * B;
    const size_t goff = (size_t)p * B + b;                                 // element [0][p][b] of the [T,P,B] arrays
    // output cursors: [t][p][b] with keep_all (advanced by one step's worth per step), else [p][b]
    float* rw_out = a.rewards_out + goff;
    float* st_out = a.states_out + goff * (PK_RAYS + 2);
    const size_t rw_step = a.keep_all ? PB : 0, st_step = rw_step * (PK_RAYS + 2);
    bool alive = false, fin = false, ok = false;
    float2 dir = make_float2(0.f, 1.f), pos = make_float2(0.f, 0.f);
    float spd = 0.f;
    float2 ext = make_float2(INF, INF);
    int act_next = 0;
    if (car_on) {
        alive = a.st.alive[k] != 0;
        fin = a.st.finishes[k] != 0;
        dir = reinterpret_cast<const float2*>(a.st.directions)[k];
        pos = reinterpret_cast<const float2*>(a.st.positions)[k];
        spd = a.st.speeds[k];
        ok = a.valid[b] != 0;
        ext = __ldg(reinterpret_cast<const float2*>(a.extent) + b);
#if GLG_F_ACT_SMEM
        if (gl < a.T) cp_async_8(&acts[0][gl], a.actions + goff + (size_t)gl * PB);          // steps 0..15
#else
        act_next = (int)__ldg(a.actions + goff);
#endif
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int pc = min(p, GLG_MAX_PLAYERS - 1);
    const float vmax = pr.vmax[pc];
    const float Lmax = ext.y;
    const float Rc = fmaf(RC_FACTOR, Lmax, 1e-3f);
    const int npass = lay.VP >> 5;                                           // stage-1 passes: 32 vertices per group and pass
    // Bit j of a lane's stage-1 words is vertex v = 32*(j>>1) + 2*gl + (j&1), which is also the first vertex of wall v.
    // Walls owned by this lane: v <= V-2; the start line (v = N-1) is appended to the list separately.
    unsigned own = 0;
    for (int j = 0; j < 2 * npass; ++j) {
        const int v = 32 * (j >> 1) + 2 * gl + (j & 1);
        if (v <= V - 2 && v != N - 1) own |= 1u << j;
    }

    if (TPB == 2) __syncwarp();
    else __syncthreads();
    if (track_on) record_copy_wait(bar);
    {   // transpose the polyline into xs[] / ys[] (padding = copies of the last vertex: those bits are never owned)
        const float2* landing = reinterpret_cast<const float2*>(tbase + lay.cars_off);
        const int stride = (TPB == 2) ? 32 : (int)blockDim.x;
        for (int v = (TPB == 2) ? lane : (int)threadIdx.x; v < lay.VP; v += stride) {
            const float2 pt = track_on ? landing[min(v, V - 1)] : make_float2(0.f, 0.f);
            xs[v] = pt.x;
            ys[v] = pt.y;
        }
    }
    if (TPB == 2) __syncwarp();
    else __syncthreads();                // from here on the landing zone is scratch

    // arg-min window (group-uniform): points [win_lo, win_lo+FW); every other point was >= m_out away from c0
    int win_lo = 0;
    float c0x = 0.f, c0y = 0.f, m_out = -1.f;

    for (int t = 0; t < a.T; ++t) {
        const int step_no = a.first_step_no + t;
#if !GLG_F_ACT_SMEM
        int act = act_next;
        if (car_on && t + 1 < a.T) act_next = (int)__ldg(a.actions + goff + (size_t)(t + 1) * PB);
#else
        if ((t & (PK_G - 1)) == 0) {
            // the block of 16 actions this step starts was requested 16 steps ago; request the next one
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int tn = t + PK_G + gl;
            if (car_on && tn < a.T) cp_async_8(&acts[((t >> 4) + 1) & 1][gl], a.actions + goff + (size_t)tn * PB);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        int act = car_on ? (int)acts[(t >> 4) & 1][t & (PK_G - 1)] : 0;
#endif
        // ---- kinematics (uniform within a group) ----
        act = min(max(act, 0), 8);
        if (!alive || !ok) act = 0;                                           // race.py:359
        const int fs = act / 3, ft = act - 3 * fs;
        const float c = pr.turn_cos[pc][fs], s_ = pr.turn_sin[pc][fs];
        const P2 nd{xadd(xmul(dir.x, c), xmul(dir.y, s_)),                    // race.py:362-364
                    xadd(xmul(dir.x, -s_), xmul(dir.y, c))};
        const float v = xadd(spd, pr.speed_inc[pc][ft]);                      // race.py:367
        float nv = fminf(vmax, fmaxf(v, 0.f));                                // race.py:369
        const bool moving = fabsf(nv) > 1e-7f;                                // race.py:370
        const P2 op{pos.x, pos.y};
        const P2 np{xadd(pos.x, xmul(nd.x, nv)), xadd(pos.y, xmul(nd.y, nv))};   // race.py:372

        // ---- progress: FIRST arg-min of |np - centre_j| (race.py:374-376) ----
        // Window pass: 16 points, one per lane.  Every point outside the window was at least m_out away from c0 when
        // the last full pass ran, hence at least m_out - |np - c0| away now; if the window's minimum is smaller than
        // that (with room for the fp32 evaluation), no outside point can win or tie, and the arg-min - first index on
        // ties of the ROUNDED norms, as the reference takes it - is decided inside the window with the very values
        // the full pass computes.
        int idx;
        {
            const float2 cw = centre[win_lo + gl];
            const float exw = xsub(np.x, cw.x), eyw = xsub(np.y, cw.y);
            const float qw = __fmaf_rn(eyw, eyw, xmul(exw, exw));
            const float qmin_w = __uint_as_float(group_min_u32(__float_as_uint(qw)));
            const float dxc = np.x - c0x, dyc = np.y - c0y;
            const float h = m_out - fmaf(sqrt_fast(fmaf(dxc, dxc, dyc * dyc)), 1.00001f, 1e-5f);
            const bool win_ok = h > 0.f && qmin_w * 1.00002f < h * h;
            {
                const float qcut = qmin_w * 1.000001f + 1e-45f;
                unsigned nearb = (__ballot_sync(FULL, qw <= qcut) >> (grp * PK_G)) & 0xffffu;
                if (__any_sync(FULL, win_ok && __popc(nearb) > 1)) {          // near tie of the rounded norms (rare)
                    const float smin = __fsqrt_rn(qmin_w);
                    nearb = (__ballot_sync(FULL, qw <= qcut && __fsqrt_rn(qw) == smin) >> (grp * PK_G)) & 0xffffu;
                }
                idx = win_lo + __ffs(nearb) - 1;
            }
            if (__any_sync(FULL, !win_ok)) {
                // full pass (the per-step kernel's), then the new window and its certificate
                float q1 = INF, q2 = INF;
                int j1 = 0x7fffffff;
                for (int j = gl; j < N; j += PK_G) {
                    const float2 cpt = centre[j];
                    const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                    const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                    const bool less = q < q1;
                    q2 = less ? q1 : fminf(q2, q);
                    j1 = less ? j : j1;
                    q1 = less ? q : q1;
                }
                const float qmin = __uint_as_float(group_min_u32(__float_as_uint(q1)));
                const float qcut = qmin * 1.000001f + 1e-45f;
                int idx_full = (int)group_min_u32((q1 == qmin) ? (unsigned)j1 : 0x7fffffffu);
                const unsigned n1 = __ballot_sync(FULL, q1 <= qcut) & gmask, n2 = __ballot_sync(FULL, q2 <= qcut) & gmask;
                const bool tie = __popc(n1) + __popc(n2) > 1;
                if (__any_sync(FULL, tie)) {
                    const float smin = __fsqrt_rn(qmin);
                    int first = 0x7fffffff;
                    for (int j = gl; j < N && tie; j += PK_G) {
                        const float2 cpt = centre[j];
                        const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                        const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                        if (q <= qcut && __fsqrt_rn(q) == smin) { first = j; break; }
                    }
                    const int f2 = (int)group_min_u32((unsigned)first);
                    if (tie) idx_full = f2;
                }
                if (!win_ok) {
                    idx = idx_full;
                    win_lo = max(0, min(idx_full - FW_BACK, N - FW));
                    c0x = np.x; c0y = np.y;
                }
                const int wl = win_lo;                                     // (a group that kept its window changes nothing)
                float qo = INF;
                for (int j = gl; j < N; j += PK_G) {
                    const float2 cpt = centre[j];
                    const float ex = np.x - cpt.x, ey = np.y - cpt.y;
                    const float q = fmaf(ey, ey, ex * ex);
                    if (j < wl || j >= wl + FW) qo = fminf(qo, q);
                }
                qo = __uint_as_float(group_min_u32(__float_as_uint(qo)));
                if (!win_ok) m_out = (N >= FW) ? sqrt_fast(qo) * 0.99999f : -1.f;     // +inf: the window holds every point
            }
        }

        // ---- scan: preconditions, ray table, stage 1 over all vertices ----
        float reward = fin ? 0.f : pr.step_penalty;                            // race.py:382-383
        const bool upd = alive && moving && ok;                                // race.py:380
        const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
        const bool safe = d2 > 0.5f && d2 < 2.f && fabsf(np.x) + fabsf(np.y) + ext.x < 200.f;
        const bool scan_on = alive && safe;
        if (gl < 2) car->tmin[PK_RAYS + gl] = 0;
        {
            P2 d, f;
#if GLG_F_RAYTAB
            const float2 cs = s_raytab[warp][gl];
#else
            const float2 cs = make_float2(pr.ray_cos[gl], pr.ray_sin[gl]);
#endif
            ray_setup_cs(cs.x, cs.y, np, nd, d, f);
            car->ray[gl] = make_float4(d.x, d.y, f.x, f.y);
            car->tmin[gl] = 0x7f800000;
            if (gl < O - PK_G) {
#if GLG_F_RAYTAB
                const float2 cs1 = s_raytab[warp][PK_G + gl];
#else
                const float2 cs1 = make_float2(pr.ray_cos[PK_G + gl], pr.ray_sin[PK_G + gl]);
#endif
                ray_setup_cs(cs1.x, cs1.y, np, nd, d, f);
                car->ray[PK_G + gl] = make_float4(d.x, d.y, f.x, f.y);
                car->tmin[PK_G + gl] = 0x7f800000;
            }
        }
        if (gl == 0) { car->nan_mask = 0; car->qn = 0; }

        // The walls the path can touch have their first end point within colR of the car (glg_sensors.cuh); when
        // colR <= Rc they are all flagged ("closer than Rc"), and their exact test rides along in the flagged-wall loop.
        const float colR = fabsf(op.x - np.x) + fabsf(op.y - np.y) + Lmax + 1e-3f;
        const bool col_on = upd && safe;
        const bool col_merged = col_on && colR <= Rc;
        int nw = 0;
        if (__any_sync(FULL, scan_on)) {
            const float close2 = scan_on ? Rc * Rc * d2 * 1.0001f : 0.f;       // scan off: nothing is flagged
            const float Kn = scan_on ? 9.f * EPS_PERP * 1.4143f * 1.001f : 0.f;
            const float2 npx = splat2(np.x), npy = splat2(np.y), ndx = splat2(nd.x), ndy = splat2(nd.y);
            const float2 m3 = splat2(-3.f), mKn = splat2(-Kn), mclose = splat2(-close2);
            unsigned sbits = 0, fbits = 0;
            const float2* xp = reinterpret_cast<const float2*>(xs) + gl;
            const float2* yp = reinterpret_cast<const float2*>(ys) + gl;
#pragma unroll S1_UNROLL
            for (int pass = 0; pass < npass; ++pass, xp += PK_G, yp += PK_G) {
                // two consecutive vertices per lane; z = (u . nd) + i (u x nd); flags from Im z^9 (glg_sensors.cuh,
                // scan_two_stage).  im3n = -Im z^3 and the bracket of the second cubing is negated too, so that no
                // operand needs a negation: im9 keeps its sign.
                const float2 ux = sub2(*xp, npx), uy = sub2(*yp, npy);
                const float2 za = fma2(ux, ndx, mul2(uy, ndy));
                const float2 zb = sub2(mul2(ux, ndy), mul2(uy, ndx));
                const float2 a2 = mul2(za, za), b2 = mul2(zb, zb);
                const float2 r2z = add2(a2, b2);
                const float2 re3 = mul2(za, fma2(b2, m3, a2));                 // a (a^2 - 3 b^2)
                const float2 im3n = mul2(zb, fma2(a2, m3, b2));                // b (b^2 - 3 a^2) = -Im z^3
                const float2 im9 = mul2(im3n, fma2(mul2(re3, re3), m3, mul2(im3n, im3n)));   // = Im (z^3)^3
                const float2 r4 = mul2(r2z, r2z);
                const float2 near = fma2(mul2(r4, r4), mKn, make_float2(fabsf(im9.x), fabsf(im9.y)));   // < 0: within EPS_PERP of a ray line
                const float2 clo = add2(r2z, mclose);                                                   // < 0: closer than Rc
                sbits = __funnelshift_l(__float_as_uint(im9.x), sbits, 1);
                sbits = __funnelshift_l(__float_as_uint(im9.y), sbits, 1);
                fbits = __funnelshift_l(__float_as_uint(near.x) | __float_as_uint(clo.x), fbits, 1);
                fbits = __funnelshift_l(__float_as_uint(near.y) | __float_as_uint(clo.y), fbits, 1);
            }
            {   // the bits arrived oldest-first: bit j sits at position 2*npass-1-j
                const int sh = 32 - 2 * npass;
                sbits = __brev(sbits) >> sh;
                fbits = __brev(fbits) >> sh;
            }
            if (!scan_on) { fbits = 0; sbits = 0; }
            // wall (v, v+1): an even bit j has its partner in the same lane (bit j+1); an odd bit in the next lane's
            // bit j-1 (same pass) - for lane 15 that is lane 0's bit j+1 (next pass)
            unsigned sn = __shfl_down_sync(FULL, sbits, 1, PK_G), fn = __shfl_down_sync(FULL, fbits, 1, PK_G);
            const unsigned s0 = __shfl_sync(FULL, sbits, 0, PK_G), f0 = __shfl_sync(FULL, fbits, 0, PK_G);
            if (gl == PK_G - 1) { sn = s0 >> 2; fn = f0 >> 2; }
            const unsigned EVEN = 0x55555555u;
            const unsigned sx = ((sbits ^ (sbits >> 1)) & EVEN) | ((sbits ^ (sn << 1)) & ~EVEN);
            const unsigned fx = ((fbits | (fbits >> 1)) & EVEN) | ((fbits | (fn << 1)) & ~EVEN);
            unsigned wbits = (sx | fx) & own;
            {
                const int cnt = __popc(wbits);
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < PK_G; off <<= 1) {
                    const int tt = __shfl_up_sync(FULL, incl, off, PK_G);
                    if (gl >= off) incl += tt;
                }
                nw = __shfl_sync(FULL, incl, PK_G - 1, PK_G);
                int posn = incl - cnt;
                while (wbits) {
                    const int j = __ffs(wbits) - 1;
                    wbits &= wbits - 1;
                    wlist[posn++] = (unsigned short)(32 * (j >> 1) + 2 * gl + (j & 1));
                }
                if (scan_on) {
                    if (gl == 0) wlist[nw] = (unsigned short)(N - 1);          // the start line, always
                    ++nw;
                } else nw = 0;
            }
        }
        __syncwarp();

        // ---- finish line (race.py:431-432) ----
        bool done = false;
        if (upd) {
            const float flx = xs[2 * N - 1], fly = ys[2 * N - 1], frx = xs[0], fry = ys[0];   // left[N-1] -> right[N-1], race.py:169
            const float x0 = fminf(op.x, np.x) - BOX_MARGIN, x1 = fmaxf(op.x, np.x) + BOX_MARGIN;
            const float y0 = fminf(op.y, np.y) - BOX_MARGIN, y1 = fmaxf(op.y, np.y) + BOX_MARGIN;
            const bool apart = fmaxf(flx, frx) < x0 || fminf(flx, frx) > x1 || fmaxf(fly, fry) < y0 || fminf(fly, fry) > y1;
            if (!apart) done = segments_cross(P2{flx, fly}, P2{frx, fry}, op, np);
        }

        // ---- stage 2: flagged walls, one per lane: exact path test (merged collision), candidate rays, the first
        //      candidate ray evaluated on the spot (race.py:287-308, 406) ----
        bool hit = false;
        {
            const unsigned all_rays = (1u << O) - 1u;
            const float sect = (float)O * (0.5f / PI_F);
            const float m_eta = ETA_ANGLE * sect, m_eps = EPS_PERP * sect, fhalf = 0.5f * (float)O;
            const float ox = op.x - np.x, oy = op.y - np.y;
            const float bx0 = fminf(ox, 0.f) - BOX_MARGIN, bx1 = fmaxf(ox, 0.f) + BOX_MARGIN;
            const float by0 = fminf(oy, 0.f) - BOX_MARGIN, by1 = fmaxf(oy, 0.f) + BOX_MARGIN;
            const int nwmax = (int)__reduce_max_sync(FULL, (unsigned)nw);
            for (int base = gl; base < nwmax; base += PK_G) {
                if (base < nw) {
                    const int w = wlist[base];
                    const float px0 = xs[w], px1 = xs[w + 1], py0 = ys[w], py1 = ys[w + 1];
                    const float ux = px0 - np.x, uy = py0 - np.y, ux1 = px1 - np.x, uy1 = py1 - np.y;
                    const bool rev = w < N;                            // wall_by_line_index
                    const P2 pp = rev ? P2{px1, py1} : P2{px0, py0}, qq = rev ? P2{px0, py0} : P2{px1, py1};
                    if (col_merged && !(fmaxf(ux, ux1) < bx0 || fminf(ux, ux1) > bx1 || fmaxf(uy, uy1) < by0 || fminf(uy, uy1) > by1))
                        hit = hit || segments_cross(pp, qq, op, np);   // race.py:406
                    unsigned mask = wall_ray_mask(ux, uy, ux1, uy1, nd, O, sect, fhalf, m_eps, m_eta, all_rays);
                    if (mask) {
                        const int i = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float4 r = car->ray[i];
                        const float tw = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
                        fused_report(car, i, tw);
                        if (mask) {                                    // further rays of this wall: second round
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
        bool wall_hit = group_ballot(hit, grp) != 0u;
        const bool col_all = col_on ? !col_merged : (upd && !safe);           // long step, or pruning preconditions failed
        if (__any_sync(FULL, col_all)) {
            const bool h = fused_collide_all(xs, ys, N, op, np, gl, grp, safe);
            if (col_all) wall_hit = h;
        }
        __syncwarp();

        // ---- reward, score (race.py:434-456) ----
        if (upd) {
            const bool dead = wall_hit;
            reward = xadd(reward, xsub(done ? 1.f : 0.f, dead ? 1.f : 0.f));   // race.py:434
            alive = alive && !dead && !done;
            fin = fin || done;
            if (gl == 0 && (dead || done))                                     // race.py:442-447 (done wins over dead)
                a.st.scores[k] = done ? step_no : idx + pr.steps_limit + 1;
        }
        if (!alive) nv = 0.f;                                                  // race.py:449
        const float drag = xsub(1.f, xmul(xsub(1.f, ft != 0 ? 1.f : 0.f), pr.drag));   // race.py:452
        const float speed = xmul(nv, drag);                                    // race.py:455
        const bool emit = a.keep_all != 0 || t == a.T - 1;
        if (car_on && gl == 0) {
            if (emit) *rw_out = reward;
            if (a.history && b == a.record_id) {                               // race.py:492-494
                float* hrow = a.history + ((size_t)step_no * P + p) * 6;
                hrow[0] = np.x; hrow[1] = np.y; hrow[2] = nd.x; hrow[3] = nd.y; hrow[4] = (float)act; hrow[5] = alive ? 1.f : 0.f;
            }
        }

        // ---- second round: the queued rays (only cars that are still alive report readings) ----
        const int total = car->qn;                                         // > PK_QCAP: some rays were dropped -> brute force
        const bool overflow = total > PK_QCAP;
        if (alive && scan_on && !overflow) {
            for (int e = gl; e < total; e += PK_G) {
                const int code = cq[e];
                const int w = code >> 5, i = code & 31;
                P2 pp, qq;
                wall_soa(xs, ys, N, w, pp, qq);
                const float4 r = car->ray[i];
                const float tw = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
                fused_report(car, i, tw);
            }
        }
        const bool brute_s = alive && (!safe || overflow);
        if (__any_sync(FULL, brute_s)) fused_sensors_brute(xs, ys, N, pr, np, nd, gl, gmask, brute_s, car);
        __syncwarp();

        // ---- observation pack [P,B,O+2] (race.py:496-500) ----
        if (car_on && emit) {
            float* out = st_out;
            const unsigned nanm = car->nan_mask;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = r * PK_G + gl;
                if (i < O + 2) {
                    float num = 0.f, den = 1.f;
                    if (i < O) {
                        if (alive) {
                            float tv_ = __int_as_float(car->tmin[i]);
                            if (nanm & (1u << i)) tv_ = __int_as_float(0x7fc00000);
                            num = (tv_ != tv_) ? tv_ : fminf(tv_, pr.max_distance);
                            den = pr.max_distance;
                        }
                    } else if (i == O) { num = speed; den = vmax; }
                    else { num = (float)idx; den = pr.progress_div; }
                    out[i] = xdiv(num, den);
                }
            }
        }
        __syncwarp();                        // the scratch of this car is rewritten by the next step
        // ---- commit (race.py:452-456) ----
        dir = make_float2(nd.x, nd.y);
        pos = make_float2(np.x, np.y);
        spd = speed;
        rw_out += rw_step;
        st_out += st_step;
    }

    // ---- car state: written once ----
    if (car_on && gl == 0) {
        reinterpret_cast<float2*>(a.st.directions)[k] = dir;
        reinterpret_cast<float2*>(a.st.positions)[k] = pos;
        a.st.speeds[k] = spd;
        a.st.alive[k] = alive ? 1 : 0;
        a.st.finishes[k] = fin ? 1 : 0;
        if (alive && a.alive_stamp) atomicMax(&a.alive_stamp[b % GLG_ALIVE_SLOTS], a.last_seq);
    }
}

}  // namespace glg
