// glg_race.cu - the batched Race environment step and its satellites (sm_100a).
//
//   glg_race_init     games/race.py:182-190   initial car state
//   glg_race_step     games/race.py:340-500   Race.step (IMPL_GPU semantics), one fused kernel
//   glg_race_rollout  T back-to-back steps with pre-computed actions
//   glg_race_winners  games/race.py:506-529   Race.winners
//   glg_winner_stats  train-gan.py:103-104    one_hot(winners+1).view(trials,-1,P+1).mean(0)
//
// Production kernel: race_step_packed_kernel (glg_race_packed.cuh, two cars per warp).  This file holds the
// launch logic and race_step_kernel<VARIANT>, the same step with one CTA per track and one warp per car:
// FAST (the packed kernel's algorithm; long or odd-length tracks), SCAN (single-pass pruning; other even ray
// counts) and BRUTE (the literal loops; odd ray counts, verification).  The CTA stages the track record
// {line, centre} (3*N float2, 3120 B at L=128) in shared memory with one bulk async copy (TMA); each warp then runs
// kinematics, progress arg-min, wall/finish collision, reward/score, and the ray-cast sensors for
// its car with the lanes striding over points / walls, and packs the [P,B,O+2] observation.
// There is no cross-car dependence in the reference step, so no global sync.
#include <stdlib.h>

#include "glg_common.cuh"
#include "glg_exact.cuh"
#include "glg_sensors.cuh"

namespace glg {

// action a -> throttle flag index a % 3 (0, +1, -3) and steering flag index a / 3 (0, +1, -1), games/race.py:52-71

struct PackedLayout { unsigned bar_off, cars_off, cq_off, wlist_off, list_len, track_bytes; };   // glg_race_packed.cuh

struct StepArgs {
    const float* geom;
    const int64_t* actions;
    const uint8_t* valid;
    const float* extent;
    glg_race_state st;
    float* states_out;
    float* rewards_out;
    int32_t* alive_stamp;
    float* history;
    int32_t* chain;       // [B,P] per-car launch stamps (glg_race_rollout), or nullptr
    const int32_t* base;  // device {step_no offset, launch_seq offset} (CUDA-graph replays), or nullptr
    int32_t B, N, step_no, record_id;
    int32_t seq;          // launch sequence number (unique, increasing per environment)
    int32_t chained;      // wait for chain[car] == seq-1 instead of for the whole previous grid
    int32_t early;        // publish chain[car] right after the state write-back (outputs of different steps do not alias)
    PackedLayout pk;      // filled by launch_packed
    unsigned long long* ll;                  // [B*P][6] LL hand-over words (packed kernel, glg_race_packed.cuh)
    int32_t ll_read, ll_write, arrays_write;
};

__global__ void race_init_kernel(glg_race_state st, int K, int32_t* alive_stamp)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < GLG_ALIVE_SLOTS && alive_stamp) alive_stamp[k] = 0;
    if (k >= K) return;
    reinterpret_cast<float2*>(st.positions)[k] = make_float2(0.f, 0.1f);   // race.py:182-183
    reinterpret_cast<float2*>(st.directions)[k] = make_float2(0.f, 1.f);   // race.py:184-185
    st.speeds[k] = 0.f;
    st.alive[k] = 1;
    st.finishes[k] = 0;
    st.scores[k] = 0;
}

// ---- bulk async copy (TMA, 1-D) of one track record into shared memory ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void record_copy_async(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void record_copy_wait(uint64_t* bar) {
    const uint32_t b = smem_u32(bar);
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(b) : "memory");
        if (spin > (1 << 22)) __trap();          // never hang the device on a lost copy
    }
}

}  // namespace glg

#include "glg_race_packed.cuh"

namespace glg {

// OC: number of rays known at compile time (0 = generic)
#ifndef GLG_STEP_MINBLOCKS
#define GLG_STEP_MINBLOCKS 6     // 256-thread blocks per SM the register allocator targets (6 -> 40 registers, 50 warps/SM;
                                 // measured best with chained rollouts: 4 -> 6.4e8, 5/6 -> 7.0e8, 7/8 -> 6.8e8 env-steps/s)
#endif

template <int VARIANT, int OC>
__global__ void __launch_bounds__(32 * GLG_MAX_PLAYERS, GLG_STEP_MINBLOCKS)
race_step_kernel(const __grid_constant__ glg_race_params pr, const StepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // every kernel parameter is baked into a captured graph: replays take the running step number from memory
    int step_no = a.step_no, seq = a.seq;
    if (a.base) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        step_no += __ldcg(a.base);
        seq += __ldcg(a.base + 1);
        if (step_no > __ldcg(a.base + 2)) return;     // past the time limit (Race.finished()): the step is a no-op
    }
    GLG_MARK_INIT;
    GLG_TRACE(0);
    GLG_TRACE(3);
    const int N = a.N, B = a.B;
    const int P = pr.num_players;
    const int O = OC ? OC : pr.num_rays;
    const int b = blockIdx.x;
    const int lane = lane_id();
    const int p = threadIdx.x >> 5;

    // ---- stage the track record {line[2N], centre[N]} (contiguous, 24N bytes) in shared memory ----
    float2* pts = reinterpret_cast<float2*>(smem_raw);
    const float2* rec = reinterpret_cast<const float2*>(a.geom) + (size_t)b * 3 * N;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + smem_barrier_offset(N));
    const uint32_t rec_bytes = (uint32_t)(3 * N * sizeof(float2));
    // one elected thread issues a single bulk copy when the record is 16-byte granular (N even)
    const bool bulk = VARIANT != GLG_STEP_BRUTE && (rec_bytes & 15u) == 0 && ((uintptr_t)a.geom & 15u) == 0;
    if (bulk) {
        if (threadIdx.x == 0) record_copy_async(pts, rec, rec_bytes, bar);
    } else {
        for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) pts[i] = __ldg(&rec[i]);
    }
    // Programmatic dependent launch: the geometry is never written by a step, so the copy above may
    // overlap the tail of the previous kernel in the stream; everything below reads what that kernel
    // (the previous step, or the policy that produced the actions) wrote and must wait for it.
    asm volatile("griddepcontrol.launch_dependents;");
    const int k = b * P + p;
    if (a.chained) {
        // Rollout with pre-computed actions: this car's previous step is the only thing the warp depends on
        // (glg_race_rollout).  The previous launch of the stream is that step's kernel and all of its CTAs have
        // started (the condition under which a programmatic dependent grid is launched), so the wait is bounded.
        if (lane == 0) {
            const int want = seq - 1;
            int got;
            int spin = 0;
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
    GLG_TRACE(1);

    // ---- car state and kinematics while the copy is in flight (uniform across the warp) ----
    // (state is read past L1: in a chained rollout the previous step may have run on another SM)
    bool alive = __ldcg(&a.st.alive[k]) != 0;
    bool fin = __ldcg(&a.st.finishes[k]) != 0;
    const bool ok = a.valid[b] != 0;
    float2 ext = make_float2(0.f, 0.f);
    if (VARIANT == GLG_STEP_FAST) ext = __ldg(reinterpret_cast<const float2*>(a.extent) + b);
    int act = (int)a.actions[(size_t)p * B + b];
    act = min(max(act, 0), 8);
    if (!alive || !ok) act = 0;                                           // race.py:359
    const int fs = act / 3, ft = act - 3 * fs;                            // race.py:52-71 (c_steer_idx / c_throttle_idx)
    const float2 dir = __ldcg(reinterpret_cast<const float2*>(a.st.directions) + k);
    const float2 pos = __ldcg(reinterpret_cast<const float2*>(a.st.positions) + k);
    const float c = pr.turn_cos[p][fs], s = pr.turn_sin[p][fs];
    const P2 nd{xadd(xmul(dir.x, c), xmul(dir.y, s)),                      // race.py:362-364
                xadd(xmul(dir.x, -s), xmul(dir.y, c))};
    const float v = xadd(__ldcg(&a.st.speeds[k]), pr.speed_inc[p][ft]);    // race.py:367
    float nv = fminf(pr.vmax[p], fmaxf(v, 0.f));                           // race.py:369
    const bool moving = fabsf(nv) > 1e-7f;                                 // race.py:370
    const P2 op{pos.x, pos.y};
    const P2 np{xadd(pos.x, xmul(nd.x, nv)), xadd(pos.y, xmul(nd.y, nv))}; // race.py:372

    GLG_MARK(1);
    __syncthreads();                     // barrier init / plain loads visible to every warp
    if (bulk) record_copy_wait(bar);
    GLG_MARK(2);
    const TrackView tv{pts, pts + 2 * N, N};
    SensorScratch* scratch = reinterpret_cast<SensorScratch*>(smem_raw + smem_scratch_offset(N)) + p;
    unsigned* maskbuf = reinterpret_cast<unsigned*>(smem_raw + smem_maskbuf_offset(N, P)) + (size_t)p * maskbuf_len(N);

    // ---- progress: FIRST arg-min of |np - centre_j| (race.py:374-376) ----
    // norm = sqrt_rn(q), q = fma(ey,ey,ex*ex); sqrt is monotone, so the winner is the first j whose
    // sqrt_rn(q_j) equals sqrt_rn(min q): only values within a few ulp of the minimum need the sqrt.
    int idx;
    {
        constexpr int KEEP = 5;                   // N <= 160 (L <= 158): every q stays in a register
        float qk[KEEP];
        float qmin = INF;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            const int j = lane + 32 * u;
            float q = INF;
            if (j < N) {
                const float2 cpt = tv.centre[j];
                const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                q = __fmaf_rn(ey, ey, xmul(ex, ex));
            }
            qk[u] = q;
            qmin = fminf(qmin, q);
        }
        for (int j = lane + 32 * KEEP; j < N; j += 32) {
            const float2 cpt = tv.centre[j];
            const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
            qmin = fminf(qmin, __fmaf_rn(ey, ey, xmul(ex, ex)));
        }
        qmin = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(qmin)));        // q >= 0
        const float smin = __fsqrt_rn(qmin);
        const float qcut = qmin * 1.000001f + 1e-45f;
        int first = 0x7fffffff;
        // almost always a single (u, lane) is within qcut; the correctly rounded sqrt is taken only there
        unsigned cand = 0;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) cand |= (qk[u] <= qcut) ? (1u << u) : 0u;
        while (cand) {
            const int u = __ffs(cand) - 1;
            cand &= cand - 1;
            float q = qk[0];
#pragma unroll
            for (int w = 1; w < KEEP; ++w) q = (u == w) ? qk[w] : q;
            if (__fsqrt_rn(q) == smin) { first = lane + 32 * u; break; }
        }
        if (__all_sync(FULL, first == 0x7fffffff)) {
            for (int j = lane + 32 * KEEP; j < N; j += 32) {
                const float2 cpt = tv.centre[j];
                const float ex = xsub(np.x, cpt.x), ey = xsub(np.y, cpt.y);
                const float q = __fmaf_rn(ey, ey, xmul(ex, ex));
                if (q <= qcut && __fsqrt_rn(q) == smin) { first = j; break; }
            }
        }
        idx = (int)__reduce_min_sync(FULL, (unsigned)first);
    }
    GLG_MARK(3);

    // ---- collisions (race.py:380-447) and sensor candidates in one pass over the polyline ----
    float reward = fin ? 0.f : pr.step_penalty;                            // race.py:382-383
    const bool upd = alive && moving && ok;                                // race.py:380
    ScanResult scan{false, true, 0};
    if (VARIANT == GLG_STEP_FAST) {
        if (alive) scan = scan_two_stage(tv, pr, np, nd, op, upd, scratch, reinterpret_cast<unsigned short*>(maskbuf), ext.x, ext.y);
    } else if (VARIANT == GLG_STEP_SCAN) {
        if (alive) scan = scan_fast<OC>(tv, pr, np, nd, op, upd, scratch, maskbuf);
    } else if (upd) {
        scan.wall_hit = collide_brute(tv, op, np);
    }
    if (upd) {
        const bool dead = scan.wall_hit;
        const float2 fl = tv.line[2 * N - 1], fr = tv.line[0];             // finish: left[N-1] -> right[N-1], race.py:169
        bool done = false;
        {   // box test first: the finish line is far away on almost every step
            const float x0 = fminf(op.x, np.x) - BOX_MARGIN, x1 = fmaxf(op.x, np.x) + BOX_MARGIN;
            const float y0 = fminf(op.y, np.y) - BOX_MARGIN, y1 = fmaxf(op.y, np.y) + BOX_MARGIN;
            const bool apart = fmaxf(fl.x, fr.x) < x0 || fminf(fl.x, fr.x) > x1 ||
                               fmaxf(fl.y, fr.y) < y0 || fminf(fl.y, fr.y) > y1;
            if (VARIANT == GLG_STEP_BRUTE || !apart)
                done = segments_cross(P2{fl.x, fl.y}, P2{fr.x, fr.y}, op, np);   // race.py:431-432
        }
        reward = xadd(reward, xsub(done ? 1.f : 0.f, dead ? 1.f : 0.f));   // race.py:434
        alive = alive && !dead && !done;                                   // race.py:414, 435
        fin = fin || done;                                                 // race.py:436
        if (lane == 0 && (dead || done)) {
            int sc = __ldcg(&a.st.scores[k]);
            if (dead) sc = idx + pr.steps_limit + 1;                       // race.py:442-444
            if (done) sc = step_no;                                      // race.py:446-447
            a.st.scores[k] = sc;
        }
    }
    GLG_MARK(9);
    if (!alive) nv = 0.f;                                                  // race.py:449
    const float drag = xsub(1.f, xmul(xsub(1.f, ft != 0 ? 1.f : 0.f), pr.drag));   // race.py:452
    const float speed = xmul(nv, drag);                                    // race.py:455
    if (lane == 0) {
        reinterpret_cast<float2*>(a.st.directions)[k] = make_float2(nd.x, nd.y);
        reinterpret_cast<float2*>(a.st.positions)[k] = make_float2(np.x, np.y);
        a.st.speeds[k] = speed;
        a.st.alive[k] = alive ? 1 : 0;
        a.st.finishes[k] = fin ? 1 : 0;
        a.rewards_out[(size_t)p * B + b] = reward;
        if (alive && a.alive_stamp) atomicMax(&a.alive_stamp[b % GLG_ALIVE_SLOTS], seq);   // launches may overlap
        if (a.history && b == a.record_id) {                               // race.py:492-494
            float* h = a.history + ((size_t)step_no * P + p) * 6;
            h[0] = np.x; h[1] = np.y; h[2] = nd.x; h[3] = nd.y; h[4] = (float)act; h[5] = alive ? 1.f : 0.f;
        }
        // the next step of this car needs nothing else from this one: let it start while the rays are cast
        if (a.chain && a.early)
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.chain + k), "r"(seq) : "memory");
    }

    // ---- sensors (race.py:459-489) and observation pack [P,B,O+2] (race.py:496-500) ----
    // lane i < O: clamp(t_i, max)/max; lane O: speed/vmax (:497); lane O+1: idx/3 (:376) - one division
    float num = 0.f, den = 1.f;
    if (alive) {
        const float t = (VARIANT != GLG_STEP_BRUTE) ? sensors_finish(tv, pr, np, nd, scratch, scan, O)
                                                    : sensors_brute(tv, pr, np, nd);
        num = (t != t) ? t : fminf(t, pr.max_distance);                    // NaN propagates like torch.clamp
        den = pr.max_distance;
    }
    GLG_MARK(11);
    float* out = a.states_out + ((size_t)p * B + b) * (O + 2);
    if (lane == O) { num = speed; den = pr.vmax[p]; }
    if (lane == O + 1) { num = (float)idx; den = pr.progress_div; }
    const float val = xdiv(num, den);
    if (lane < O + 2) out[lane] = val;
    if (O + 2 > 32 && lane == 0) {                                         // O in {31, 32}
        if (O == 32) out[O] = xdiv(speed, pr.vmax[p]);
        out[O + 1] = xdiv((float)idx, pr.progress_div);
    }
    if (a.chain && !a.early) {           // publish "this car's step `seq` is complete" (all lanes' stores first)
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(a.chain + k), "r"(seq) : "memory");
        }
    }
    GLG_MARK(12);
    GLG_TRACE(2);
    // (see race_step_packed_kernel: makes chained launches complete in order)
    if (a.chained) asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void race_winners_kernel(const int32_t* __restrict__ scores, const uint8_t* __restrict__ finishes,
                                    const uint8_t* __restrict__ valid, int B, int P, int steps_limit,
                                    int64_t* __restrict__ winners)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    bool anyf = false;
    for (int p = 0; p < P; ++p) anyf = anyf || finishes[b * P + p];
    int best = 0, bv = 0;
    for (int p = 0; p < P; ++p) {
        const int sc = scores[b * P + p];
        // finished boards: arg-min with non-finishers at steps_limit+1; others: arg-max (first index)
        const int val = anyf ? (finishes[b * P + p] ? sc : steps_limit + 1) : -sc;
        if (p == 0 || val < bv) { bv = val; best = p; }
    }
    winners[b] = valid[b] ? (int64_t)best : (int64_t)-1;                   // race.py:528
}

__global__ void winner_stats_kernel(const int64_t* __restrict__ winners, int trials, int boards, int P,
                                    float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= boards * (P + 1)) return;
    const int board = i / (P + 1), cls = i % (P + 1);
    float acc = 0.f;                                     // torch .float().mean(0): sequential sum / trials
    for (int t = 0; t < trials; ++t) acc += (winners[(size_t)t * boards + board] + 1 == cls) ? 1.f : 0.f;
    out[i] = acc / (float)trials;
}

static int check_step_args(const glg_race_params* pr, const float* geom, int B, int N, const void* actions,
                           const void* valid, const void* extent, int variant, const glg_race_state& st,
                           const void* so, const void* ro)
{
    GLG_REQUIRE(variant == GLG_STEP_FAST || variant == GLG_STEP_BRUTE || variant == GLG_STEP_SCAN || variant == GLG_STEP_PACKED,
                "glg_race_step: unknown variant %d", variant);
    GLG_REQUIRE((variant != GLG_STEP_FAST && variant != GLG_STEP_PACKED) || extent != nullptr || B == 0,
                "glg_race_step: GLG_STEP_PACKED / GLG_STEP_FAST need the track extents (glg_track_extent)");
    GLG_REQUIRE(pr != nullptr, "glg_race_step: params is null");
    GLG_REQUIRE(pr->num_players >= 1 && pr->num_players <= GLG_MAX_PLAYERS, "glg_race_step: num_players %d out of range", pr->num_players);
    GLG_REQUIRE(pr->num_rays >= 1 && pr->num_rays <= GLG_MAX_RAYS, "glg_race_step: num_rays %d out of range", pr->num_rays);
    GLG_REQUIRE(B >= 0 && N >= 2 && N <= 512, "glg_race_step: need B >= 0, 2 <= N <= 512 (B=%d N=%d)", B, N);
    if (B == 0) return GLG_OK;
    GLG_REQUIRE(geom && actions && valid && so && ro, "glg_race_step: null pointer");
    GLG_REQUIRE(st.positions && st.directions && st.speeds && st.alive && st.finishes && st.scores,
                "glg_race_step: null state pointer");
    return GLG_OK;
}

// Programmatic dependent launch is requested for every step launch, also while the stream is being captured (the
// capture records a programmatic dependency edge, so graph replays keep the overlap of consecutive steps);
// GLG_GRAPH_PDL=0 in the environment falls back to plain stream order inside captured graphs.
static bool pdl_allowed(cudaStream_t stream)
{
    static const int graph_pdl = [] { const char* e = getenv("GLG_GRAPH_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
    if (graph_pdl) return true;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    return cap == cudaStreamCaptureStatusNone;
}

template <int VARIANT, int OC>
static void launch_one(const glg_race_params* pr, const StepArgs& a, bool pdl, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.B);
    cfg.blockDim = dim3(32 * pr->num_players);
    cfg.dynamicSmemBytes = smem_total(a.N, pr->num_players);
    cfg.stream = stream;
    // long tracks with many cars need more than the default 48 KB of dynamic shared memory: opt in per instantiation
    static size_t smem_opted = 48 * 1024;
    if (cfg.dynamicSmemBytes > smem_opted) {
        cudaFuncSetAttribute(race_step_kernel<VARIANT, OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        smem_opted = cfg.dynamicSmemBytes;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, race_step_kernel<VARIANT, OC>, *pr, a);
}

// the variant that actually runs (PACKED covers 18 rays, N <= 256 and 16-byte granular records = even N)
static int effective_variant(const glg_race_params* pr, const float* geom, int N, int variant)
{
    if (variant == GLG_STEP_PACKED && (pr->num_rays != 18 || N > 256 || (N & 1) || ((uintptr_t)geom & 15u)))
        variant = GLG_STEP_FAST;
    if (variant == GLG_STEP_FAST && pr->num_rays != 18) variant = GLG_STEP_SCAN;    // stage 1 is written for 9 ray lines
    if (variant == GLG_STEP_SCAN && (pr->num_rays & 1)) variant = GLG_STEP_BRUTE;   // pruning pairs opposite rays
    if (variant == GLG_STEP_SCAN && (2 * N - 1 + 30) / 31 > 32) variant = GLG_STEP_BRUTE;   // 32-bit pass bitmap
    return variant;
}

template <int TPB>
static void launch_packed(const glg_race_params* pr, const StepArgs& a, bool pdl, cudaStream_t stream)
{
    const int WPT = (pr->num_players + 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((a.B + TPB - 1) / TPB);
    cfg.blockDim = dim3(32 * WPT * TPB);
    // Residency: the two-tracks-per-CTA shape (64 registers) holds 16 CTAs per SM, which is where a chained config-2
    // rollout peaks (a launch has 2048 CTAs: more slots only hold CTAs that wait for their predecessor; measured
    // with 48 registers: 14 CTAs/SM -3 %, 18 -7 %, 21 -13 %).  One track of 3..8 cars per CTA runs unconstrained.
    cfg.dynamicSmemBytes = (size_t)TPB * pk_track_bytes(a.N, 2 * WPT);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    StepArgs ap = a;
    ap.pk = pk_layout(a.N, 2 * WPT);
    cudaLaunchKernelEx(&cfg, race_step_packed_kernel<TPB>, *pr, ap);
}

// `variant` is the effective one (effective_variant); `pdl` from pdl_allowed, once per API call
static int launch_step(const glg_race_params* pr, const StepArgs& a, int variant, bool pdl, cudaStream_t stream)
{
    if (variant == GLG_STEP_PACKED) {
        if (pr->num_players <= 2) launch_packed<2>(pr, a, pdl, stream);
        else launch_packed<1>(pr, a, pdl, stream);
        return GLG_OK;
    }
    if (variant == GLG_STEP_BRUTE) launch_one<GLG_STEP_BRUTE, 0>(pr, a, pdl, stream);
    else if (variant == GLG_STEP_FAST) launch_one<GLG_STEP_FAST, 18>(pr, a, pdl, stream);
    else if (pr->num_rays == 18) launch_one<GLG_STEP_SCAN, 18>(pr, a, pdl, stream);
    else launch_one<GLG_STEP_SCAN, 0>(pr, a, pdl, stream);
    return GLG_OK;
}

}  // namespace glg

#include "glg_race_fused.cuh"

namespace glg {

// GLG_ROLLOUT_FUSED: one persistent launch for all T steps (configurations GLG_STEP_PACKED covers)
static void launch_fused_rollout(const glg_race_params* pr, const float* geom, int B, int N, const int64_t* actions, int T,
                                 const uint8_t* valid, const float* extent, const glg_race_state& st, int first_step_no,
                                 float* states_out, float* rewards_out, int keep_all, int32_t* alive_stamp,
                                 int first_launch_seq, float* history, int record_id, bool pdl, cudaStream_t stream)
{
    const int P = pr->num_players;
    const int WPT = (P + 1) / 2;
    const int TPB = P <= 2 ? 2 : 1;
    FusedArgs a{};
    a.geom = geom; a.actions = actions; a.valid = valid; a.extent = extent; a.st = st;
    a.states_out = states_out; a.rewards_out = rewards_out; a.alive_stamp = alive_stamp; a.history = history;
    a.B = B; a.N = N; a.T = T; a.first_step_no = first_step_no; a.record_id = history ? record_id : -1;
    a.last_seq = first_launch_seq + T - 1; a.keep_all = keep_all ? 1 : 0;
    a.lay = fused_layout(N, 2 * WPT);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((B + TPB - 1) / TPB);
    cfg.blockDim = dim3(32 * WPT * TPB);
    cfg.dynamicSmemBytes = (size_t)TPB * a.lay.track_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    // instantiations: the reference's own track length (N = 130) with 2 / 4 car slots, and the generic ones
    using kernel_t = void (*)(const glg_race_params, const FusedArgs);
    kernel_t kernel;
    int slot;
    if (TPB == 2 && N == 130) { kernel = race_rollout_fused_kernel<2, 130, 2>; slot = 0; }
    else if (TPB == 2) { kernel = race_rollout_fused_kernel<2, 0, 0>; slot = 1; }
    else if (N == 130 && WPT == 2) { kernel = race_rollout_fused_kernel<1, 130, 4>; slot = 2; }
    else { kernel = race_rollout_fused_kernel<1, 0, 0>; slot = 3; }
    static size_t smem_opted[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
    if (cfg.dynamicSmemBytes > smem_opted[slot]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        smem_opted[slot] = cfg.dynamicSmemBytes;
    }
    cudaLaunchKernelEx(&cfg, kernel, *pr, a);
}

}  // namespace glg

#ifdef GLG_PHASE_CLOCKS
extern "C" int glg_debug_phases(unsigned long long* out32, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, glg::g_phase, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(glg::g_phase, z, sizeof(z)); }
    return 0;
}
extern "C" int glg_debug_trace(unsigned long long* out, int n)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, glg::g_trace, sizeof(unsigned long long) * n);
    return 0;
}
#endif

extern "C" int glg_race_init(glg_race_state st, int32_t B, int32_t P, int32_t* alive_stamp, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && P >= 1 && P <= GLG_MAX_PLAYERS, "glg_race_init: bad extents B=%d P=%d", B, P);
    const int K = B * P;
    GLG_REQUIRE(K == 0 || (st.positions && st.directions && st.speeds && st.alive && st.finishes && st.scores),
                "glg_race_init: null state pointer");
    const int n = K > GLG_ALIVE_SLOTS ? K : GLG_ALIVE_SLOTS;
    race_init_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(st, K, alive_stamp);
    return launch_status("glg_race_init");
}

extern "C" int glg_race_step(const glg_race_params* params, const float* geom, int32_t B, int32_t N,
                             const int64_t* actions, const uint8_t* valid, const float* extent, glg_race_state state,
                             int32_t step_no, float* states_out, float* rewards_out,
                             int32_t* alive_stamp, int32_t launch_seq, const int32_t* base, float* history,
                             int32_t record_id, int32_t variant, glg_stream_t stream)
{
    using namespace glg;
    const int rc = check_step_args(params, geom, B, N, actions, valid, extent, variant, state, states_out, rewards_out);
    if (rc != GLG_OK || B == 0) return rc;
    StepArgs a{geom, actions, valid, extent, state, states_out, rewards_out, alive_stamp, history, nullptr, base,
               B, N, step_no, record_id, launch_seq, 0, 0, {}, nullptr, 0, 0, 1};
    launch_step(params, a, effective_variant(params, geom, N, variant), pdl_allowed((cudaStream_t)stream), (cudaStream_t)stream);
    return launch_status("glg_race_step");
}

extern "C" int glg_race_rollout(const glg_race_params* params, const float* geom, int32_t B, int32_t N,
                                const int64_t* actions, int32_t T, const uint8_t* valid, const float* extent,
                                glg_race_state state,
                                int32_t first_step_no, float* states_out, float* rewards_out, int32_t keep_all,
                                int32_t* alive_stamp, int32_t first_launch_seq, int32_t* chain,
                                float* history, int32_t record_id, int32_t mode,
                                int32_t variant, glg_stream_t stream)
{
    using namespace glg;
    const int rc = check_step_args(params, geom, B, N, actions, valid, extent, variant, state, states_out, rewards_out);
    if (rc != GLG_OK || B == 0 || T <= 0) return rc;
    GLG_REQUIRE(mode == GLG_ROLLOUT_STEPWISE || mode == GLG_ROLLOUT_CHAINED || mode == GLG_ROLLOUT_FUSED,
                "glg_race_rollout: unknown mode %d", mode);
    GLG_REQUIRE(mode != GLG_ROLLOUT_CHAINED || chain != nullptr, "glg_race_rollout: GLG_ROLLOUT_CHAINED needs the chain scratch");
    const size_t PB = (size_t)params->num_players * B;
    const size_t W = params->num_rays + 2;
    const int eff = effective_variant(params, geom, N, variant);
    const bool pdl = pdl_allowed((cudaStream_t)stream);
    if (mode == GLG_ROLLOUT_FUSED) {
        if (eff == GLG_STEP_PACKED) {
            launch_fused_rollout(params, geom, B, N, actions, T, valid, extent, state, first_step_no, states_out,
                                 rewards_out, keep_all, alive_stamp, first_launch_seq, history, record_id, pdl,
                                 (cudaStream_t)stream);
            return launch_status("glg_race_rollout");
        }
        mode = chain ? GLG_ROLLOUT_CHAINED : GLG_ROLLOUT_STEPWISE;   // configurations the fused kernel does not cover
    }
    if (mode == GLG_ROLLOUT_STEPWISE) chain = nullptr;
    // LL hand-over (packed kernel, every step has its own output buffer): see glg_race_packed.cuh
    const bool ll = chain != nullptr && keep_all && eff == GLG_STEP_PACKED;
    unsigned long long* llw = chain ? reinterpret_cast<unsigned long long*>(chain + ((PB + 3) & ~(size_t)3)) : nullptr;
    for (int t = 0; t < T; ++t) {
        StepArgs a{geom, actions + (size_t)t * PB, valid, extent, state,
                   keep_all ? states_out + (size_t)t * PB * W : states_out,
                   keep_all ? rewards_out + (size_t)t * PB : rewards_out,
                   alive_stamp, history, chain, nullptr, B, N, first_step_no + t, history ? record_id : -1, first_launch_seq + t,
                   (chain != nullptr && t > 0) ? 1 : 0, keep_all ? 1 : 0, {}, llw,
                   (ll && t > 0) ? 1 : 0, (ll && t < T - 1) ? 1 : 0, (!ll || t == T - 1) ? 1 : 0};
        launch_step(params, a, eff, pdl, (cudaStream_t)stream);
    }
    return launch_status("glg_race_rollout");
}

extern "C" int64_t glg_race_chain_bytes(int32_t B, int32_t P)
{
    const int64_t cars = (int64_t)B * P;
    return (((cars + 3) & ~(int64_t)3) + 12 * cars) * (int64_t)sizeof(int32_t);
}

extern "C" int glg_race_winners(const int32_t* scores, const uint8_t* finishes, const uint8_t* valid,
                                int32_t B, int32_t P, int32_t steps_limit, int64_t* winners, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && P >= 1 && P <= GLG_MAX_PLAYERS, "glg_race_winners: bad extents B=%d P=%d", B, P);
    if (B == 0) return GLG_OK;
    GLG_REQUIRE(scores && finishes && valid && winners, "glg_race_winners: null pointer");
    race_winners_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scores, finishes, valid, B, P, steps_limit, winners);
    return launch_status("glg_race_winners");
}

extern "C" int glg_winner_stats(const int64_t* winners, int32_t trials, int32_t boards, int32_t P,
                                float* out, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(trials >= 1 && boards >= 0 && P >= 1 && P <= GLG_MAX_PLAYERS, "glg_winner_stats: bad extents");
    if (boards == 0) return GLG_OK;
    GLG_REQUIRE(winners && out, "glg_winner_stats: null pointer");
    const int n = boards * (P + 1);
    winner_stats_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(winners, trials, boards, P, out);
    return launch_status("glg_winner_stats");
}
