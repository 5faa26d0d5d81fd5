// glg_race.cu - the batched Race environment step and its satellites (sm_100a).
//
//   glg_race_init     games/race.py:182-190   initial car state
//   glg_race_step     games/race.py:340-500   Race.step (IMPL_GPU semantics), one fused kernel
//   glg_race_rollout  T back-to-back steps with pre-computed actions
//   glg_race_winners  games/race.py:506-529   Race.winners
//   glg_winner_stats  train-gan.py:103-104    one_hot(winners+1).view(trials,-1,P+1).mean(0)
//
// Mapping: one CTA per track, one warp per car.  The CTA stages the track record
// {right, left, centre} (3*N float2, 3120 B at L=128) in shared memory once; each warp then runs
// kinematics, progress arg-min, wall/finish collision, reward/score, and the ray-cast sensors for
// its car with the lanes striding over points / walls, and packs the [P,B,O+2] observation.
// There is no cross-car dependence in the reference step, so no global sync.
#include "glg_common.cuh"
#include "glg_exact.cuh"
#include "glg_sensors.cuh"

namespace glg {

// action -> (throttle flag index, steering flag index), games/race.py:52-71
__device__ __constant__ int8_t c_throttle_idx[9] = {0, 1, 2, 0, 1, 2, 0, 1, 2};
__device__ __constant__ int8_t c_steer_idx[9] = {0, 0, 0, 1, 1, 1, 2, 2, 2};

struct StepArgs {
    const float* geom;
    const int64_t* actions;
    const uint8_t* valid;
    glg_race_state st;
    float* states_out;
    float* rewards_out;
    int32_t* alive_stamp;
    float* history;
    int32_t B, N, step_no, record_id;
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

template <int VARIANT>
__global__ void __launch_bounds__(32 * GLG_MAX_PLAYERS)
race_step_kernel(const __grid_constant__ glg_race_params pr, const StepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, B = a.B;
    const int P = pr.num_players, O = pr.num_rays;
    const int b = blockIdx.x;
    const int lane = lane_id();
    const int p = threadIdx.x >> 5;

    // ---- stage the track record (coalesced 8-byte loads; the record is contiguous in HBM) ----
    float2* pts = reinterpret_cast<float2*>(smem_raw);
    {
        const float2* rec = reinterpret_cast<const float2*>(a.geom) + (size_t)b * 3 * N;
        for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) pts[i] = __ldg(&rec[i]);
    }
    __syncthreads();
    TrackView tv{pts, pts + N, pts + 2 * N, N};
    SensorScratch* scratch = reinterpret_cast<SensorScratch*>(smem_raw + sensor_scratch_offset(N)) + p;

    // ---- car state (uniform across the warp) ----
    const int k = b * P + p;
    bool alive = a.st.alive[k] != 0;
    bool fin = a.st.finishes[k] != 0;
    const bool ok = a.valid[b] != 0;
    int act = (int)a.actions[(size_t)p * B + b];
    act = min(max(act, 0), 8);
    if (!alive || !ok) act = 0;                                           // race.py:359
    const int fs = c_steer_idx[act], ft = c_throttle_idx[act];

    const float2 dir = reinterpret_cast<const float2*>(a.st.directions)[k];
    const float2 pos = reinterpret_cast<const float2*>(a.st.positions)[k];
    const float c = pr.turn_cos[p][fs], s = pr.turn_sin[p][fs];
    const P2 nd{xadd(xmul(dir.x, c), xmul(dir.y, s)),                      // race.py:362-364
                xadd(xmul(dir.x, -s), xmul(dir.y, c))};
    const float v = xadd(a.st.speeds[k], pr.speed_inc[p][ft]);             // race.py:367
    float nv = fminf(pr.vmax[p], fmaxf(v, 0.f));                           // race.py:369
    const bool moving = fabsf(nv) > 1e-7f;                                 // race.py:370
    const P2 op{pos.x, pos.y};
    const P2 np{xadd(pos.x, xmul(nd.x, nv)), xadd(pos.y, xmul(nd.y, nv))}; // race.py:372

    // ---- progress: first arg-min of the distance to the centre points (race.py:374-376) ----
    int idx = 0;
    {
        float best = INF;
        for (int j = lane; j < N; j += 32) {
            const float2 cpt = tv.centre[j];
            const float d = norm2(xsub(np.x, cpt.x), xsub(np.y, cpt.y));
            if (d < best) { best = d; idx = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(FULL, best, off);
            const int oj = __shfl_xor_sync(FULL, idx, off);
            if (od < best || (od == best && oj < idx)) { best = od; idx = oj; }
        }
    }

    // ---- collisions with the walls and the finish line (race.py:380-447) ----
    float reward = fin ? 0.f : pr.step_penalty;                            // race.py:382-383
    const bool upd = alive && moving && ok;                                // race.py:380
    if (upd) {
        const int S = N - 1;
        bool hit = false;
        for (int j = lane; j < 2 * S + 1; j += 32) {                       // race.py:406-407
            P2 wp, wq;
            wall_points(tv, j, wp, wq);
            hit = hit || segments_cross(wp, wq, op, np);
        }
        const bool dead = __any_sync(FULL, hit);
        const float2 fl = tv.left[N - 1], fr = tv.right[N - 1];            // race.py:169, 431-432
        const bool done = segments_cross(P2{fl.x, fl.y}, P2{fr.x, fr.y}, op, np);
        reward = xadd(reward, xsub(done ? 1.f : 0.f, dead ? 1.f : 0.f));   // race.py:434
        alive = alive && !dead && !done;                                   // race.py:414, 435
        fin = fin || done;                                                 // race.py:436
        if (lane == 0 && (dead || done)) {
            int sc = a.st.scores[k];
            if (dead) sc = idx + pr.steps_limit + 1;                       // race.py:442-444
            if (done) sc = a.step_no;                                      // race.py:446-447
            a.st.scores[k] = sc;
        }
    }
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
        if (alive && a.alive_stamp) a.alive_stamp[b % GLG_ALIVE_SLOTS] = a.step_no;
        if (a.history && b == a.record_id) {                               // race.py:492-494
            float* h = a.history + ((size_t)a.step_no * P + p) * 6;
            h[0] = np.x; h[1] = np.y; h[2] = nd.x; h[3] = nd.y; h[4] = (float)act; h[5] = alive ? 1.f : 0.f;
        }
    }

    // ---- sensors (race.py:459-489): lane i ends up holding the reading of ray i ----
    float obs = 0.f;
    if (alive) {
        float t;
        if (VARIANT == GLG_STEP_BRUTE) t = sensors_brute(tv, pr, np, nd, scratch);
        else t = sensors_fast(tv, pr, np, nd, scratch);
        // clamp(max)/max_distance (race.py:489); NaN propagates like torch.clamp
        obs = (t != t) ? t : xdiv(fminf(t, pr.max_distance), pr.max_distance);
    }
    // ---- observation pack [P,B,O+2] (race.py:496-500) ----
    float* out = a.states_out + ((size_t)p * B + b) * (O + 2);
    if (lane < O) out[lane] = obs;
    else if (lane == O) out[O] = xdiv(speed, pr.vmax[p]);                  // race.py:497
    else if (lane == O + 1) out[O + 1] = xdiv((float)idx, pr.progress_div);   // race.py:376
    if (O + 2 > 32 && lane == 0) {                                         // O in {31, 32}
        if (O == 31) out[O + 1] = xdiv((float)idx, pr.progress_div);
        else { out[O] = xdiv(speed, pr.vmax[p]); out[O + 1] = xdiv((float)idx, pr.progress_div); }
    }
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
                           const void* valid, const glg_race_state& st, const void* so, const void* ro)
{
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

static int launch_step(const glg_race_params* pr, const StepArgs& a, int variant, cudaStream_t stream)
{
    const int P = pr->num_players;
    if (variant == GLG_STEP_FAST && (pr->num_rays & 1)) variant = GLG_STEP_BRUTE;   // pruning pairs opposite rays
    const size_t smem = sensor_scratch_offset(a.N) + (size_t)P * sizeof(SensorScratch);
    if (variant == GLG_STEP_BRUTE)
        race_step_kernel<GLG_STEP_BRUTE><<<a.B, 32 * P, smem, stream>>>(*pr, a);
    else
        race_step_kernel<GLG_STEP_FAST><<<a.B, 32 * P, smem, stream>>>(*pr, a);
    return GLG_OK;
}

}  // namespace glg

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
                             const int64_t* actions, const uint8_t* valid, glg_race_state state,
                             int32_t step_no, float* states_out, float* rewards_out,
                             int32_t* alive_stamp, float* history, int32_t record_id,
                             int32_t variant, glg_stream_t stream)
{
    using namespace glg;
    const int rc = check_step_args(params, geom, B, N, actions, valid, state, states_out, rewards_out);
    if (rc != GLG_OK || B == 0) return rc;
    GLG_REQUIRE(variant == GLG_STEP_FAST || variant == GLG_STEP_BRUTE, "glg_race_step: unknown variant %d", variant);
    StepArgs a{geom, actions, valid, state, states_out, rewards_out, alive_stamp, history, B, N, step_no, record_id};
    launch_step(params, a, variant, (cudaStream_t)stream);
    return launch_status("glg_race_step");
}

extern "C" int glg_race_rollout(const glg_race_params* params, const float* geom, int32_t B, int32_t N,
                                const int64_t* actions, int32_t T, const uint8_t* valid, glg_race_state state,
                                int32_t first_step_no, float* states_out, float* rewards_out, int32_t keep_all,
                                int32_t* alive_stamp, int32_t variant, glg_stream_t stream)
{
    using namespace glg;
    const int rc = check_step_args(params, geom, B, N, actions, valid, state, states_out, rewards_out);
    if (rc != GLG_OK || B == 0 || T <= 0) return rc;
    GLG_REQUIRE(variant == GLG_STEP_FAST || variant == GLG_STEP_BRUTE, "glg_race_rollout: unknown variant %d", variant);
    const size_t PB = (size_t)params->num_players * B;
    const size_t W = params->num_rays + 2;
    for (int t = 0; t < T; ++t) {
        StepArgs a{geom, actions + (size_t)t * PB, valid, state,
                   keep_all ? states_out + (size_t)t * PB * W : states_out,
                   keep_all ? rewards_out + (size_t)t * PB : rewards_out,
                   alive_stamp, nullptr, B, N, first_step_no + t, -1};
        launch_step(params, a, variant, (cudaStream_t)stream);
    }
    return launch_status("glg_race_rollout");
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
