// glg_helpers.cu - the entry points of the reference's pybind module `game_helpers` (sm_100a).
//
//   glg_collision / glg_smallest_distance / glg_is_valid   games/game_helpers.cpp:335-450 (free functions)
//   glg_game_*                                             games/game_helpers.cpp:158-327 (class Game)
//
// The reference backs intersects / intersection / distance with Boost.Geometry (system headers, no
// pinned version, absent from this image) - "parity unpinned" for those.  Semantics implemented here:
//   intersects(polyline, segment)  : some polyline segment and the probe share a point; decided with
//                                    the reference's own Boost-free test (game_helpers.cpp:49-66).
//   intersection + distance        : Euclidean distance from the ray origin to the nearest common point
//                                    of the polyline and the segment origin -> origin + 1000*d; +inf if none.
//   intersects(polyline)           : two non-adjacent segments share a point, or two adjacent segments
//                                    share more than their common end point.
// Game::update_players is Boost-free in the reference and is restated literally.
#include <stdlib.h>

#include "glg_common.cuh"
#include "glg_exact.cuh"

namespace glg {

__device__ __forceinline__ P2 ld2(const float2* p) { const float2 v = *p; return P2{v.x, v.y}; }

__device__ __forceinline__ float dist2(P2 a, P2 b) {
    const float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}

// distance from the origin s to the nearest common point of wall (p,q) and segment (s,f); +inf if disjoint
__device__ float hit_distance(P2 p, P2 q, P2 s, P2 f) {
    if (!segments_cross(p, q, s, f)) return INF;
    const float wx = q.x - p.x, wy = q.y - p.y;
    const float dx = f.x - s.x, dy = f.y - s.y;
    const float den = dy * wx - dx * wy;
    if (den != 0.f) {
        const float num = (p.y - s.y) * wx - (p.x - s.x) * wy;
        float t = num / den;
        t = fminf(fmaxf(t, 0.f), 1.f);
        return t * sqrtf(dx * dx + dy * dy);
    }
    // collinear overlap: nearest point of the shared part
    if (in_box(p, q, s)) return 0.f;
    float d = INF;
    if (in_box(s, f, p)) d = fminf(d, dist2(p, s));
    if (in_box(s, f, q)) d = fminf(d, dist2(q, s));
    return d;
}

// one warp per (track, probe)
__global__ void collision_kernel(const float2* __restrict__ tracks, const float4* __restrict__ segs,
                                 uint8_t* __restrict__ out, int b, int s, int p)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= b * p) return;
    const int lane = lane_id();
    const float2* line = tracks + (size_t)(w / p) * s;
    const float4 sg = segs[w];
    const P2 a{sg.x, sg.y}, c{sg.z, sg.w};
    bool hit = false;
    for (int j = lane; j < s - 1; j += 32) hit = hit || segments_cross(ld2(line + j), ld2(line + j + 1), a, c);
    hit = __any_sync(FULL, hit);
    if (lane == 0) out[w] = hit ? 1 : 0;
}

// point j of the Game polyline: left reversed, then right (game_helpers.cpp:127-138)
__device__ __forceinline__ P2 game_line_point(const float2* left, const float2* right, int s, int j) {
    return j < s ? ld2(left + (s - 1 - j)) : ld2(right + (j - s));
}

// one warp per (row, direction); GAME = polyline taken from the Game workspace through idx
template <bool GAME>
__global__ void distance_kernel(const float2* __restrict__ tracks, const float2* __restrict__ left,
                                const float2* __restrict__ right, const int64_t* __restrict__ idx,
                                int num_players, const float4* __restrict__ dirs, float* __restrict__ out,
                                int rows, int s, int d)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= rows * d) return;
    const int lane = lane_id();
    const int row = w / d;
    const float4 r = dirs[w];
    const P2 o{r.x, r.y};
    const P2 f{r.x + 1000.f * r.z, r.y + 1000.f * r.w};
    float best = INF;
    if (GAME) {
        const int trk = (int)(idx[row] / num_players);
        const float2* l = left + (size_t)trk * s;
        const float2* rr = right + (size_t)trk * s;
        for (int j = lane; j < 2 * s - 1; j += 32)
            best = fminf(best, hit_distance(game_line_point(l, rr, s, j), game_line_point(l, rr, s, j + 1), o, f));
    } else {
        const float2* line = tracks + (size_t)row * s;
        for (int j = lane; j < s - 1; j += 32) best = fminf(best, hit_distance(ld2(line + j), ld2(line + j + 1), o, f));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = fminf(best, __shfl_xor_sync(FULL, best, off));
    if (lane == 0) out[w] = best;
}

// self-intersection of a polyline of n points given by an accessor; one CTA per track
template <bool GAME>
__global__ void self_intersect_kernel(const float2* __restrict__ tracks, const float2* __restrict__ left,
                                      const float2* __restrict__ right, uint8_t* __restrict__ valid, int s)
{
    extern __shared__ float2 pts[];
    __shared__ int bad;
    const int b = blockIdx.x;
    const int n = GAME ? 2 * s : s;
    if (threadIdx.x == 0) bad = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        P2 v = GAME ? game_line_point(left + (size_t)b * s, right + (size_t)b * s, s, j) : ld2(tracks + (size_t)b * s + j);
        pts[j] = make_float2(v.x, v.y);
    }
    __syncthreads();
    const int m = n - 1;                               // segments
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const P2 a{pts[i].x, pts[i].y}, c{pts[i + 1].x, pts[i + 1].y};
        bool hit = false;
        if (i + 1 < m) {                               // adjacent pair: more than the shared end point
            const P2 e{pts[i + 2].x, pts[i + 2].y};
            hit = (turn(a, c, e) == 0 && in_box(a, c, e)) || (turn(c, e, a) == 0 && in_box(c, e, a));
        }
        for (int j = i + 2; j < m && !hit; ++j) {
            const P2 e{pts[j].x, pts[j].y}, g{pts[j + 1].x, pts[j + 1].y};
            hit = segments_cross(a, c, e, g);
        }
        if (hit) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) valid[b] = bad ? 0 : 1;
}

__global__ void game_init_kernel(const float2* __restrict__ left, const float2* __restrict__ right,
                                 float2* __restrict__ wl, float2* __restrict__ wr, float2* __restrict__ ppos,
                                 int32_t* __restrict__ pseg, int npts, int nplayers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) { wl[i] = left[i]; wr[i] = right[i]; }
    if (i < nplayers) { ppos[i] = make_float2(0.f, 0.1f); pseg[i] = 0; }     // game_helpers.cpp:154-155
}

// literal restatement of Game::update_players, game_helpers.cpp:209-276; one thread per row
__global__ void game_update_kernel(const float2* __restrict__ left, const float2* __restrict__ right,
                                   float2* __restrict__ ppos, int32_t* __restrict__ pseg,
                                   const int64_t* __restrict__ idx, const float* __restrict__ new_pos,
                                   int k, int row_stride, int s, int num_players,
                                   uint8_t* __restrict__ dead, uint8_t* __restrict__ finished)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int pl = (int)idx[i];
    const int trk = pl / num_players;
    const float2* L = left + (size_t)trk * s;
    const float2* R = right + (size_t)trk * s;
    const int length = s - 1;                                   // :140
    const P2 np{new_pos[(size_t)i * row_stride], new_pos[(size_t)i * row_stride + 1]};
    const float2 pp = ppos[pl];
    const P2 pos{pp.x, pp.y};
    const int seg0 = pseg[pl];
    int next = seg0;
    bool alive = true, done = false;
    // `next_seg` is an int and `track.length` a size_t (game_helpers.cpp:105, 215): both comparisons of the forward
    // part are UNSIGNED in the reference.  A car that backed out over the start line (cell -1, reported dead then)
    // therefore skips the forward walk and is reported `finished` (and not dead) by every later call, its cell
    // staying at -1 - reproduced here, pinned by tests/golden/game_update.npz (case C).
    while ((unsigned)next < (unsigned)length) {                  // :220-238 forward walk
        const P2 la = ld2(L + next), lb = ld2(L + next + 1), ra = ld2(R + next), rb = ld2(R + next + 1);
        if (segments_cross(la, lb, pos, np) || segments_cross(ra, rb, pos, np)) { alive = false; break; }
        if (turn(lb, rb, np) > 0) break;                         // stayed in this cell
        ++next;
    }
    if (alive && (unsigned)next >= (unsigned)length) done = true;    // :240-243
    if (next == seg0 && alive && !done) {                        // :245-269 backward walk
        while (next >= 0) {
            const P2 la = ld2(L + next), lb = ld2(L + next + 1), ra = ld2(R + next), rb = ld2(R + next + 1);
            if (segments_cross(la, lb, pos, np) || segments_cross(ra, rb, pos, np)) { alive = false; break; }
            if (turn(la, ra, np) < 0) break;
            --next;
        }
        if (next < 0) alive = false;
    }
    ppos[pl] = make_float2(np.x, np.y);                          // :271-272
    pseg[pl] = next;
    dead[i] = alive ? 0 : 1;
    finished[i] = done ? 1 : 0;
}

}  // namespace glg

struct glg_game {
    float2* left;
    float2* right;
    float2* ppos;
    int32_t* pseg;
    int32_t b, s, num_players;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int glg_collision(const float* tracks, const float* segments, uint8_t* out,
                             int32_t b, int32_t s, int32_t p, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(b >= 0 && s >= 2 && p >= 0, "glg_collision: bad extents b=%d s=%d p=%d", b, s, p);
    if (b * p == 0) return GLG_OK;
    GLG_REQUIRE(tracks && segments && out, "glg_collision: null pointer");
    const int warps = b * p;
    collision_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        (const float2*)tracks, (const float4*)segments, out, b, s, p);
    return launch_status("glg_collision");
}

extern "C" int glg_smallest_distance(const float* tracks, const float* directions, float* out,
                                     int32_t b, int32_t s, int32_t d, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(b >= 0 && s >= 2 && d >= 0, "glg_smallest_distance: bad extents b=%d s=%d d=%d", b, s, d);
    if (b * d == 0) return GLG_OK;
    GLG_REQUIRE(tracks && directions && out, "glg_smallest_distance: null pointer");
    const int warps = b * d;
    distance_kernel<false><<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        (const float2*)tracks, nullptr, nullptr, nullptr, 1, (const float4*)directions, out, b, s, d);
    return launch_status("glg_smallest_distance");
}

extern "C" int glg_is_valid(const float* tracks, uint8_t* out, int32_t b, int32_t s, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(b >= 0 && s >= 2 && s <= 4096, "glg_is_valid: bad extents b=%d s=%d", b, s);
    if (b == 0) return GLG_OK;
    GLG_REQUIRE(tracks && out, "glg_is_valid: null pointer");
    self_intersect_kernel<false><<<b, 256, (size_t)s * sizeof(float2), (cudaStream_t)stream>>>(
        (const float2*)tracks, nullptr, nullptr, out, s);
    return launch_status("glg_is_valid");
}

extern "C" int64_t glg_game_workspace_bytes(int32_t b, int32_t s, int32_t num_players)
{
    if (b < 0 || s < 2 || num_players < 1) return -1;
    const size_t pts = (size_t)b * s * sizeof(float2);
    const size_t pl = (size_t)b * num_players;
    return (int64_t)(2 * align256(pts) + align256(pl * sizeof(float2)) + align256(pl * sizeof(int32_t)));
}

extern "C" int glg_game_create(glg_game** out, void* workspace, int64_t workspace_bytes,
                               const float* left, const float* right, int32_t b, int32_t s,
                               int32_t num_players, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(out != nullptr, "glg_game_create: out is null");
    *out = nullptr;
    const int64_t need = glg_game_workspace_bytes(b, s, num_players);
    GLG_REQUIRE(need >= 0, "glg_game_create: bad extents b=%d s=%d players=%d", b, s, num_players);
    GLG_REQUIRE(s <= 2048, "glg_game_create: at most 2048 points per boundary");
    GLG_REQUIRE(b == 0 || (workspace && left && right), "glg_game_create: null pointer");
    GLG_REQUIRE(workspace_bytes >= need, "glg_game_create: workspace too small (%lld < %lld)",
                (long long)workspace_bytes, (long long)need);
    GLG_REQUIRE(((uintptr_t)workspace & 255) == 0, "glg_game_create: workspace must be 256-byte aligned");
    glg_game* g = (glg_game*)malloc(sizeof(glg_game));
    GLG_REQUIRE(g != nullptr, "glg_game_create: out of host memory");
    char* base = (char*)workspace;
    const size_t pts = align256((size_t)b * s * sizeof(float2));
    const size_t pl = (size_t)b * num_players;
    g->left = (float2*)base;
    g->right = (float2*)(base + pts);
    g->ppos = (float2*)(base + 2 * pts);
    g->pseg = (int32_t*)(base + 2 * pts + align256(pl * sizeof(float2)));
    g->b = b; g->s = s; g->num_players = num_players;
    if (b > 0) {
        const int npts = b * s, n = npts > (int)pl ? npts : (int)pl;
        game_init_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            (const float2*)left, (const float2*)right, g->left, g->right, g->ppos, g->pseg, npts, (int)pl);
        const int rc = launch_status("glg_game_create");
        if (rc != GLG_OK) { free(g); return rc; }
    }
    *out = g;
    return GLG_OK;
}

extern "C" void glg_game_destroy(glg_game* game) { free(game); }

extern "C" int glg_game_validate_tracks(glg_game* game, uint8_t* valid, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(game != nullptr, "glg_game_validate_tracks: null handle");
    if (game->b == 0) return GLG_OK;
    GLG_REQUIRE(valid != nullptr, "glg_game_validate_tracks: null pointer");
    self_intersect_kernel<true><<<game->b, 256, (size_t)2 * game->s * sizeof(float2), (cudaStream_t)stream>>>(
        nullptr, game->left, game->right, valid, game->s);
    return launch_status("glg_game_validate_tracks");
}

extern "C" int glg_game_update_players(glg_game* game, const int64_t* idx, const float* new_positions,
                                       int32_t k, int32_t row_stride, uint8_t* dead, uint8_t* finished,
                                       glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(game != nullptr, "glg_game_update_players: null handle");
    GLG_REQUIRE(k >= 0 && row_stride >= 2, "glg_game_update_players: bad extents k=%d row_stride=%d", k, row_stride);
    if (k == 0) return GLG_OK;
    GLG_REQUIRE(idx && new_positions && dead && finished, "glg_game_update_players: null pointer");
    game_update_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        game->left, game->right, game->ppos, game->pseg, idx, new_positions, k, row_stride, game->s,
        game->num_players, dead, finished);
    return launch_status("glg_game_update_players");
}

extern "C" int glg_game_smallest_distance(glg_game* game, const int64_t* idx, const float* directions,
                                          int32_t k, int32_t d, float* out, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(game != nullptr, "glg_game_smallest_distance: null handle");
    GLG_REQUIRE(k >= 0 && d >= 0, "glg_game_smallest_distance: bad extents k=%d d=%d", k, d);
    if (k * d == 0) return GLG_OK;
    GLG_REQUIRE(idx && directions && out, "glg_game_smallest_distance: null pointer");
    const int warps = k * d;
    distance_kernel<true><<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        nullptr, game->left, game->right, idx, game->num_players, (const float4*)directions, out, k, game->s, d);
    return launch_status("glg_game_smallest_distance");
}
