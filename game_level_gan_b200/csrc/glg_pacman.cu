// glg_pacman.cu - batched grid Pacman: step and observation build (sm_100a).
//
//   glg_pacman_step     games/pacman.py:64-105  (_not_blocked + step without the observation part)
//   glg_pacman_observe  games/pacman.py:107-109 (per-player float observation)
//
// State in HBM: grid [B,H,W,4+P] int32 (channels: empty, wall, small pellet, large pellet, one layer per
// player - the P "this is me" planes of the reference's [B,H,W,4+2P] grid are constant zeros and only
// exist in the observation), players [B*P,4] int32 rows (board, x, y, player id) in np.where order.
// Integer / byte work, HBM-bound: the observation write (P*B*H*W*(4+2P)*4 bytes) dominates.
#include "glg_common.cuh"

namespace glg {

__device__ __constant__ int c_move_dx[5] = {0, -1, 1, 0, 0};    // games/pacman.py:22-23: noop, up, down, left, right
__device__ __constant__ int c_move_dy[5] = {0, 0, 0, -1, 1};

__device__ __forceinline__ bool pacman_target(const int32_t* __restrict__ grid, const int32_t* row, int act,
                                              int H, int W, int C, int& nx, int& ny)
{
    act = min(max(act, 0), 4);
    nx = row[1] + c_move_dx[act];
    ny = row[2] + c_move_dy[act];
    if (nx < 0 || ny < 0 || nx >= H || ny >= W) return false;                       // pacman.py:65-66
    return grid[(((size_t)row[0] * H + nx) * W + ny) * C + 1] == 0;                   // pacman.py:67-69 (wall layer)
}

// the reference skips the whole update when no player of the whole batch can move (pacman.py:77)
__global__ void pacman_any_kernel(const int32_t* __restrict__ grid, const int32_t* __restrict__ players,
                                  const int32_t* __restrict__ actions, int32_t* __restrict__ flag,
                                  int K, int H, int W, int C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int nx, ny;
    if (pacman_target(grid, players + 4 * (size_t)k, actions[k], H, W, C, nx, ny)) *flag = 1;
}

// one thread per board
__global__ void pacman_apply_kernel(int32_t* __restrict__ grid, int32_t* __restrict__ players,
                                    const int32_t* __restrict__ actions, double* __restrict__ rewards,
                                    const int32_t* __restrict__ flag, int B, int H, int W, int P)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int C = 4 + P;
    if (*flag == 0) {                                                                // pacman.py:76-77
        for (int j = 0; j < P; ++j) rewards[(size_t)j * B + b] = 0.0;
        return;
    }
    int32_t* rows = players + 4 * (size_t)b * P;
    int px[GLG_MAX_PLAYERS], py[GLG_MAX_PLAYERS];
    for (int j = 0; j < P; ++j) {                                                    // pacman.py:79-83
        int nx, ny;
        const bool free_cell = pacman_target(grid, rows + 4 * j, actions[b * P + j], H, W, C, nx, ny);
        grid[(((size_t)rows[4 * j] * H + rows[4 * j + 1]) * W + rows[4 * j + 2]) * C + 4 + rows[4 * j + 3]] = 0;
        px[j] = free_cell ? nx : rows[4 * j + 1];
        py[j] = free_cell ? ny : rows[4 * j + 2];
    }
    for (int j = 0; j < P; ++j) {                                                    // pacman.py:83-87
        rows[4 * j + 1] = px[j];
        rows[4 * j + 2] = py[j];
        grid[(((size_t)rows[4 * j] * H + px[j]) * W + py[j]) * C + 4 + rows[4 * j + 3]] = 1;
    }
    for (int j = 0; j < P; ++j) {                                                    // pacman.py:89-100
        const size_t cell = (((size_t)rows[4 * j] * H + px[j]) * W + py[j]) * C;
        const double r = 0.5 * (double)grid[cell + 2] + (double)grid[cell + 3];
        int freq = 0;
        for (int i = 0; i < P; ++i) freq += (rows[4 * i] == rows[4 * j] && px[i] == px[j] && py[i] == py[j]) ? 1 : 0;
        rewards[(size_t)j * B + b] = r / (double)(float)freq;
    }
    for (int j = 0; j < P; ++j) {                                                    // pacman.py:102-104
        const size_t cell = (((size_t)rows[4 * j] * H + px[j]) * W + py[j]) * C;
        grid[cell + 2] = 0;
        grid[cell + 3] = 0;
    }
}

// one thread per grid cell: reads 4+P ints once, writes P rows of 4+2P floats.
// PC = compile-time number of players (0 = generic): with it the per-cell arrays live in registers and the row of
// an even-sized observation goes out as float4 stores.
template <int PC>
__global__ void pacman_observe_kernel(const int32_t* __restrict__ grid, float* __restrict__ obs,
                                      size_t cells, int P_rt)
{
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    if (PC > 0) {
        constexpr int C = 4 + PC, D = 4 + 2 * PC;
        float v[C];
        if ((C & 1) == 0) {                                                // 8-byte aligned rows
            const int2* g = reinterpret_cast<const int2*>(grid + cell * C);
#pragma unroll
            for (int c = 0; c < C / 2; ++c) { const int2 t = __ldg(g + c); v[2 * c] = (float)t.x; v[2 * c + 1] = (float)t.y; }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = (float)__ldg(grid + cell * C + c);
        }
#pragma unroll
        for (int p = 0; p < PC; ++p) {
            float* o = obs + ((size_t)p * cells + cell) * D;
            float row[D];
#pragma unroll
            for (int c = 0; c < D; ++c) row[c] = c < C ? v[c] : (c == C + p ? 1.f : 0.f);
            if ((D & 3) == 0) {
#pragma unroll
                for (int c = 0; c < D; c += 4)
                    __stcs(reinterpret_cast<float4*>(o + c), make_float4(row[c], row[c + 1], row[c + 2], row[c + 3]));
            } else {
#pragma unroll
                for (int c = 0; c < D; c += 2)
                    __stcs(reinterpret_cast<float2*>(o + c), make_float2(row[c], row[c + 1]));
            }
        }
    } else {
        const int P = P_rt, C = 4 + P, D = 4 + 2 * P;
        for (int p = 0; p < P; ++p) {
            float* o = obs + ((size_t)p * cells + cell) * D;
            for (int c = 0; c < D; ++c) o[c] = c < C ? (float)grid[cell * C + c] : (c == C + p ? 1.f : 0.f);
        }
    }
}

}  // namespace glg

extern "C" int glg_pacman_step(int32_t* grid, int32_t* players, const int32_t* actions, double* rewards,
                               int32_t* scratch, int32_t B, int32_t H, int32_t W, int32_t P, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && H >= 1 && W >= 1 && P >= 1 && P <= GLG_MAX_PLAYERS,
                "glg_pacman_step: bad extents B=%d H=%d W=%d P=%d", B, H, W, P);
    if (B == 0) return GLG_OK;
    GLG_REQUIRE(grid && players && actions && rewards && scratch, "glg_pacman_step: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(scratch, 0, sizeof(int32_t), s);
    const int K = B * P;
    pacman_any_kernel<<<(K + 255) / 256, 256, 0, s>>>(grid, players, actions, scratch, K, H, W, 4 + P);
    pacman_apply_kernel<<<(B + 127) / 128, 128, 0, s>>>(grid, players, actions, rewards, scratch, B, H, W, P);
    return launch_status("glg_pacman_step");
}

extern "C" int glg_pacman_observe(const int32_t* grid, float* obs, int32_t B, int32_t H, int32_t W, int32_t P,
                                  glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && H >= 1 && W >= 1 && P >= 1 && P <= GLG_MAX_PLAYERS,
                "glg_pacman_observe: bad extents B=%d H=%d W=%d P=%d", B, H, W, P);
    const size_t cells = (size_t)B * H * W;
    if (cells == 0) return GLG_OK;
    GLG_REQUIRE(grid && obs, "glg_pacman_observe: null pointer");
    const unsigned blocks = (unsigned)((cells + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (P) {
        case 1: pacman_observe_kernel<1><<<blocks, 256, 0, s>>>(grid, obs, cells, P); break;
        case 2: pacman_observe_kernel<2><<<blocks, 256, 0, s>>>(grid, obs, cells, P); break;
        case 3: pacman_observe_kernel<3><<<blocks, 256, 0, s>>>(grid, obs, cells, P); break;
        case 4: pacman_observe_kernel<4><<<blocks, 256, 0, s>>>(grid, obs, cells, P); break;
        default: pacman_observe_kernel<0><<<blocks, 256, 0, s>>>(grid, obs, cells, P); break;
    }
    return launch_status("glg_pacman_observe");
}
