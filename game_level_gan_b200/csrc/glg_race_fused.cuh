// glg_race_fused.cuh - the persistent rollout kernel: ONE launch plays T steps (glg_race_rollout, GLG_ROLLOUT_FUSED).
//
// Same mapping and the same arithmetic as race_step_packed_kernel (two cars per warp, 16 lanes each; P <= 2: one
// warp per track and two tracks per CTA, else one track per CTA with ceil(P/2) warps), but a warp keeps its track
// and its cars for the whole rollout:
//   * the track record is staged in shared memory ONCE (one bulk async copy per track) and stays there,
//   * the car state lives in registers from step to step; the state arrays are read once and written once,
//   * per step the only global traffic is the car's action (8 B, fetched one step ahead), its observation
//     (80 B) and its reward (4 B) - 92 B per car-step instead of 3120 B per track + 145 B per car,
//   * no launch, no hand-over between launches, no grid-wide dependency: cars never interact, so
//     every warp runs at its own pace.
// Results are bit-identical to T calls of glg_race_step (tests/test_race_gpu.py, tests/test_fused_rollout_gpu.py).
// Included by glg_race.cu after glg_race_packed.cuh.
#pragma once
#include "glg_race_packed.cuh"

namespace glg {

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
    unsigned bar_off, cars_off, cq_off, wlist_off, list_len, track_bytes;
};

// shared memory per track: [record 3N float2][mbarrier 16 B][cars x PackedCar][cars x cq u16[LL]][cars x wlist u16[2N]]
// (unlike the per-step kernel the wall lists cannot live in the centre points: those are needed again next step)
__host__ inline void fused_layout(FusedArgs& a, int N, int cars) {
    a.bar_off = (unsigned)smem_barrier_offset(N);
    a.cars_off = pk_cars_offset(N);
    a.cq_off = pk_cq_offset(N, cars);
    a.wlist_off = pk_wlist_offset(N, cars);
    a.list_len = pk_list_len(N);
    a.track_bytes = (a.wlist_off + (unsigned)cars * 2u * (unsigned)N * 2u + 127u) & ~127u;
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

template <int TPB>
__global__ void __launch_bounds__(TPB == 2 ? 64 : 128, TPB == 2 ? GLG_FUSED_MINBLOCKS2 : GLG_FUSED_MINBLOCKS1)
race_rollout_fused_kernel(const __grid_constant__ glg_race_params pr, const FusedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int O = PK_RAYS;
    const int N = a.N, B = a.B, V = 2 * N;
    const int P = pr.num_players;
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
    const int ci = wt * 2 + grp;

    unsigned char* tbase = smem_raw + (unsigned)tslot * a.track_bytes;
    float2* pts = reinterpret_cast<float2*>(tbase);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tbase + a.bar_off);
    PackedCar* car = reinterpret_cast<PackedCar*>(tbase + a.cars_off) + ci;
    unsigned short* cq = reinterpret_cast<unsigned short*>(tbase + a.cq_off) + (unsigned)ci * a.list_len;
    unsigned short* wlist = reinterpret_cast<unsigned short*>(tbase + a.wlist_off) + (unsigned)ci * 2u * (unsigned)N;

    // everything this kernel reads may have been written by the previous kernel of the stream (reset, state restore)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t rec_bytes = (uint32_t)(3 * N * sizeof(float2));
    if (track_on && wt == 0 && lane == 0)
        record_copy_async(pts, reinterpret_cast<const float2*>(a.geom) + (size_t)b * 3 * N, rec_bytes, bar);
    asm volatile("griddepcontrol.launch_dependents;");

    // ---- car state: read once ----
    const int k = b * P + p;
    const size_t PB = (size_t)P * B;
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
        act_next = (int)__ldg(a.actions + (size_t)p * B + b);
    }
    const int pc = min(p, GLG_MAX_PLAYERS - 1);
    const float vmax = pr.vmax[pc];
    const float Lmax = ext.y;
    const float Rc = fmaf(RC_FACTOR, Lmax, 1e-3f);
    const int passes = (V + PK_G - 1) / PK_G;                              // <= 32 (the host routes N > 256 elsewhere)
    const int nown = (V - 2 - gl >= 0) ? ((V - 2 - gl) >> 4) + 1 : 0;      // walls w = 16*pass + gl <= V-2
    unsigned own = nown >= 32 ? FULL : ((1u << nown) - 1u);
    if (((N - 1) & 15) == gl) own &= ~(1u << ((N - 1) >> 4));              // the start line is appended separately

    if (TPB == 2) __syncwarp();
    else __syncthreads();
    if (track_on) record_copy_wait(bar);
    const TrackView tv{pts, pts + 2 * N, N};

    for (int t = 0; t < a.T; ++t) {
        const int step_no = a.first_step_no + t;
        int act = act_next;
        if (car_on && t + 1 < a.T) act_next = (int)__ldg(a.actions + (size_t)(t + 1) * PB + (size_t)p * B + b);
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

        // ---- progress: FIRST arg-min of |np - centre_j| (race.py:374-376), see race_step_kernel ----
        int idx;
        {
            float q1 = INF, q2 = INF;
            int j1 = 0x7fffffff;
            {
                const float2* cp = tv.centre + gl;
                const int full = N / PK_G;
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
                if (j < N) {
                    const float2 cpt = *cp;
                    const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                    const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                    const bool less = q < q1;
                    q2 = less ? q1 : fminf(q2, q);
                    j1 = less ? j : j1;
                    q1 = less ? q : q1;
                }
            }
            const float qmin = __uint_as_float(group_min_u32(__float_as_uint(q1)));
            const float qcut = qmin * 1.000001f + 1e-45f;
            idx = (int)group_min_u32((q1 == qmin) ? (unsigned)j1 : 0x7fffffffu);
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

        // ---- scan: preconditions, ray table, stage 1 over all vertices ----
        float reward = fin ? 0.f : pr.step_penalty;                            // race.py:382-383
        const bool upd = alive && moving && ok;                                // race.py:380
        const float d2 = fmaf(nd.x, nd.x, nd.y * nd.y);
        const bool safe = d2 > 0.5f && d2 < 2.f && fabsf(np.x) + fabsf(np.y) + ext.x < 200.f;
        const bool scan_on = alive && safe;
        if (gl < 2) car->tmin[PK_RAYS + gl] = 0;
        {
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

        const float close2 = scan_on ? Rc * Rc * d2 * 1.0001f : 0.f;
        const float colR = fabsf(op.x - np.x) + fabsf(op.y - np.y) + Lmax + 1e-3f;
        const bool col_on = upd && safe;
        const float col2 = col_on ? colR * colR * d2 * 1.0001f : 0.f;
        const float Kn = scan_on ? 9.f * EPS_PERP * 1.4143f * 1.001f : 0.f;
        unsigned sbits = 0, fbits = 0, cbits = 0;
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
        if (!scan_on) { fbits = 0; sbits = 0; }
        if (!col_on) cbits = 0;
        unsigned s1 = __shfl_down_sync(FULL, sbits, 1, PK_G), f1 = __shfl_down_sync(FULL, fbits, 1, PK_G);
        const unsigned s0 = __shfl_sync(FULL, sbits, 0, PK_G), f0 = __shfl_sync(FULL, fbits, 0, PK_G);
        if (gl == PK_G - 1) { s1 = s0 >> 1; f1 = f0 >> 1; }
        unsigned wbits = scan_on ? (((sbits ^ s1) | fbits | f1) & own) : 0u;
        cbits &= own;
        int nw, nc = 0;
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
#pragma unroll
            for (int r = 0; r < 3; ++r) {
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
            while (live) {
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

        // ---- finish line, reward, score (race.py:431-456) ----
        if (upd) {
            const bool dead = wall_hit;
            const float2 fl = tv.line[2 * N - 1], fr = tv.line[0];
            bool done = false;
            {
                const float x0 = fminf(op.x, np.x) - BOX_MARGIN, x1 = fmaxf(op.x, np.x) + BOX_MARGIN;
                const float y0 = fminf(op.y, np.y) - BOX_MARGIN, y1 = fmaxf(op.y, np.y) + BOX_MARGIN;
                const bool apart = fmaxf(fl.x, fr.x) < x0 || fminf(fl.x, fr.x) > x1 ||
                                   fmaxf(fl.y, fr.y) < y0 || fminf(fl.y, fr.y) > y1;
                if (!apart) done = segments_cross(P2{fl.x, fl.y}, P2{fr.x, fr.y}, op, np);
            }
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
        const size_t obase = a.keep_all ? (size_t)t * PB : 0;
        if (car_on && gl == 0) {
            if (emit) a.rewards_out[obase + (size_t)p * B + b] = reward;
            if (a.history && b == a.record_id) {                               // race.py:492-494
                float* h = a.history + ((size_t)step_no * P + p) * 6;
                h[0] = np.x; h[1] = np.y; h[2] = nd.x; h[3] = nd.y; h[4] = (float)act; h[5] = alive ? 1.f : 0.f;
            }
        }

        // ---- stage 2: candidate rays of the flagged walls; the first ray of a wall is evaluated on the spot ----
        const bool sense = alive && scan_on;                       // `alive` is post-update here: dead cars report zeros
        // (without keep_all the observation of every step is still computed, like T calls of glg_race_step; only the
        //  last one is stored)
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
                        const bool rev = w < N;
                        const P2 pp = rev ? P2{p1.x, p1.y} : P2{p0.x, p0.y}, qq = rev ? P2{p0.x, p0.y} : P2{p1.x, p1.y};
                        const float4 r = car->ray[i];
                        const float tw = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
                        if (tw != tw) atomicOr(&car->nan_mask, 1u << i);
                        else atomicMin(&car->tmin[i], __float_as_int(tw));
                        if (mask) {
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
        const int total = car->qn;
        const bool overflow = total > PK_QCAP;
        if (sense && !overflow) {
            for (int e = gl; e < total; e += PK_G) {
                const int code = cq[e];
                const int w = code >> 5, i = code & 31;
                P2 pp, qq;
                wall_by_line_index(tv, w, pp, qq);
                const float4 r = car->ray[i];
                const float tw = ray_wall_t_fast(pp, qq, np, P2{r.x, r.y}, P2{r.z, r.w});
                if (tw != tw) atomicOr(&car->nan_mask, 1u << i);
                else atomicMin(&car->tmin[i], __float_as_int(tw));
            }
        }
        const bool brute_s = alive && (!safe || overflow);
        if (__any_sync(FULL, brute_s)) packed_sensors_brute(tv, pr, np, nd, gl, gmask, brute_s, car);
        __syncwarp();

        // ---- observation pack [P,B,O+2] (race.py:496-500) ----
        if (car_on && emit) {
            float* out = a.states_out + (obase + (size_t)p * B + b) * (O + 2);
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
