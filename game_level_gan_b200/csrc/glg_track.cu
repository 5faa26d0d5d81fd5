// glg_track.cu - generator output -> track geometry, and track validity (sm_100a).
//
//   glg_track_build    replaces games/race.py:126-158  (Race.reset, geometry part)
//   glg_track_validate replaces games/race.py:326-334  (Race._is_correct) as used at :199-200
//   glg_track_extent   per-track bounding radius / longest wall (no reference counterpart; the
//                      step kernel's exact pruning needs both, csrc/glg_sensors.cuh)
//
// HBM layout of the result: one record {right[N] reversed, left[N], centre[N]} of float2 per track
// (include/glg_b200.h, glg_common.cuh TrackView).  Both kernels are "reset-time" work, amortised over hundreds of steps.
#include <math.h>

#include "glg_common.cuh"
#include "glg_exact.cuh"

namespace glg {

constexpr int BUILD_WARPS = 4;

// One warp per track.  The two prefix sums replicate ATen's CPU cumsum bit-for-bit: a sequential
// running sum in double, every prefix rounded to float - therefore one lane walks
// the 130 elements; everything else (sin/cos, normals, offsets, stores) is lane-parallel.
__global__ void __launch_bounds__(BUILD_WARPS * 32)
track_build_kernel(const float* __restrict__ tracks, const uint8_t* __restrict__ levels, int B, int L,
                   const float* __restrict__ sin_table, const float* __restrict__ cos_table,
                   int table_half, float* __restrict__ geom)
{
    extern __shared__ float smem[];
    const int N = L + 2;
    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const int b = blockIdx.x * BUILD_WARPS + wib;
    if (b >= B) return;                                   // whole warp leaves together
    const int Np = (N + 1) & ~1;                          // even, so that the float2 arrays stay 8-byte aligned
    float* head = smem + wib * 5 * Np;                    // [N] heading prefix (in arc units)
    float2* seg = reinterpret_cast<float2*>(head + Np);   // [N] segment vectors
    float2* cen = seg + Np;                               // [N] centre points
    const float2* trk = reinterpret_cast<const float2*>(tracks) + (size_t)b * L;
    // generator levels: segment j-1 is nibble (j-1)&1 of byte (j-1)>>1, arc = (level - 4) / 4, width 0
    const uint8_t* lev = levels ? levels + (size_t)b * ((L + 1) >> 1) : nullptr;

    // sentinels (race.py:136-138) + "all arcs are multiples of 1/4" test for the table path
    bool quant = sin_table != nullptr;
    for (int j = lane; j < N; j += 32) {
        float arc = 0.f;
        if (j >= 1 && j <= L)
            arc = lev ? (float)((int)((__ldg(&lev[(j - 1) >> 1]) >> (4 * ((j - 1) & 1))) & 15) - 4) * 0.25f
                      : __ldg(&trk[j - 1]).x;
        head[j] = arc;
        const float a4 = arc * 4.f;
        if (a4 != rintf(a4) || fabsf(a4) > 4.f) quant = false;
    }
    quant = __all_sync(FULL, quant);
    __syncwarp();
    if (lane == 0) {                                      // race.py:140 cumsum (double accumulator)
        double acc = 0.0;
        for (int j = 0; j < N; ++j) {
            acc += (double)head[j];
            head[j] = (float)acc;
        }
    }
    __syncwarp();
    const float rad8 = (float)0.13962634015954636;        // math.radians(8.) as an fp32 scalar
    for (int j = lane; j < N; j += 32) {                  // race.py:140-142
        const float h = head[j];
        const float ang = xmul(rad8, h);
        const int n = (int)(h * 4.f);
        float s, c;
        if (quant && n >= -table_half && n <= table_half) {
            s = __ldg(&sin_table[n + table_half]);
            c = __ldg(&cos_table[n + table_half]);
        } else {
            s = sinf(ang);
            c = cosf(ang);
        }
        seg[j] = make_float2(xmul(s, 0.2f), xmul(c, 0.2f));
    }
    __syncwarp();
    if (lane < 2) {                                       // race.py:154-156 exclusive cumsum, x and y
        double acc = 0.0;
        float* dst = reinterpret_cast<float*>(cen) + lane;
        const float* src = reinterpret_cast<const float*>(seg) + lane;
        for (int j = 0; j < N; ++j) {
            dst[2 * j] = (float)acc;
            acc += (double)src[2 * j];
        }
    }
    __syncwarp();
    float2* rec = reinterpret_cast<float2*>(geom) + (size_t)b * 3 * N;
    for (int j = lane; j < N; j += 32) {                  // race.py:144-152, 157-158
        float ox = 0.5f, oy = 0.f;
        if (j > 0) {
            const float2 s1 = seg[j], s0 = seg[j - 1];
            const float nx = xadd(s1.y, s0.y);            // perp = (y, -x)
            const float ny = xadd(-s1.x, -s0.x);
            const float len = norm2(nx, ny);
            const float wid = (!lev && j >= 2 && j - 1 <= L) ? __ldg(&trk[j - 2]).y : 0.f;
            const float w = xadd(0.5f, xmul(1.5f, wid));
            ox = xmul(xdiv(nx, len), w);
            oy = xmul(xdiv(ny, len), w);
        }
        const float2 c = cen[j];
        rec[N - 1 - j] = make_float2(xadd(c.x, ox), xadd(c.y, oy));      // right, stored reversed
        rec[N + j] = make_float2(xadd(c.x, -ox), xadd(c.y, -oy));        // left
        rec[2 * N + j] = c;                                              // centre
    }
}

// One CTA per track; the M = 2(N-1)+2 lines live in shared memory and every unordered pair is
// tested once ((o1*o2<0)&(o3*o4<0) is symmetric in the pair).  Rows i and M-1-i are folded into
// one work item so that every thread runs M-1 pair tests.
__global__ void track_validate_kernel(const float* __restrict__ geom, int B, int N,
                                      uint8_t* __restrict__ valid)
{
    extern __shared__ float4 lines[];
    __shared__ int bad_flag;
    const int b = blockIdx.x;
    const int S = N - 1;
    const int M = 2 * S + 2;
    const float2* rec = reinterpret_cast<const float2*>(geom) + (size_t)b * 3 * N;
    if (threadIdx.x == 0) bad_flag = 0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {   // race.py:166-172 + finish line :169
        float2 p, q;
        if (j < S) { p = rec[N - 1 - j]; q = rec[N - 2 - j]; }              // right[j] -> right[j+1]
        else if (j < 2 * S) { p = rec[N + j - S]; q = rec[N + j - S + 1]; }  // left[j-S] -> left[j-S+1]
        else if (j == 2 * S) { p = rec[N]; q = rec[N - 1]; }                 // start: left[0] -> right[0]
        else { p = rec[2 * N - 1]; q = rec[0]; }                             // finish: left[N-1] -> right[N-1]
        lines[j] = make_float4(p.x, p.y, q.x, q.y);
    }
    __syncthreads();
    const int half = M / 2;
    for (int w = threadIdx.x; w < half; w += blockDim.x) {
        const int first = M - 1 - w;                       // pairs in row w; the rest go to row M-1-w
        int i = w;
        float4 a = lines[i];
        for (int m = 0; m < M - 1; ++m) {
            int j;
            if (m < first) j = w + 1 + m;
            else {
                if (m == first) { i = M - 1 - w; a = lines[i]; }
                j = i + 1 + (m - first);
            }
            const float4 c = lines[j];
            const P2 ap{a.x, a.y}, aq{a.z, a.w}, cp{c.x, c.y}, cq{c.z, c.w};
            const int o1 = turn(ap, aq, cp), o2 = turn(ap, aq, cq);
            const int o3 = turn(cp, cq, ap), o4 = turn(cp, cq, aq);
            if (o1 * o2 < 0 && o3 * o4 < 0) bad_flag = 1;
            if ((m & 63) == 63 && *(volatile int*)&bad_flag) break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) valid[b] = bad_flag ? 0 : 1;
}

// Same result from the SIGN MATRIX.  The 2N lines of a track are the consecutive pairs of the cyclic polyline
// c[0..2N) = right[N-1] ... right[0], left[0] ... left[N-1] (walls; start line = (c[N-1], c[N]); finish line =
// (c[2N-1], c[0])), and the reference's o1..o4 of a pair of lines are the orientation of an END POINT of one line with
// respect to the other (games/race.py:230-238, 252-255).  End points are shared by neighbouring lines, so the
// 4 x 33 670 orientation values of the pair loop are only 2N x 2N distinct ones: S[i][v] = sign of line i (in the
// reference's p -> q order) against vertex v, the same separately rounded arithmetic.  Phase 1 evaluates them once, 32
// vertices per warp instruction, and keeps two bit rows per line (negative, positive).  Line j "straddles" line i when
// S[i][.] is strictly negative at one end point of j and strictly positive at the other - (o1*o2 < 0) - which for
// consecutive vertices is a shift and two ANDs of the bit rows; the track is invalid iff some pair straddles each other.
// Half the arithmetic of the pair loop and none of its sign conversions; results identical (tests/test_race_gpu.py).
constexpr int VB_THREADS = 128;
__host__ __device__ inline int vb_words(int N) { return (2 * N + 31) >> 5; }
__host__ __device__ inline size_t vb_smem(int N) {      // xs, ys [2N] f32; neg / pos bit rows [2N][W] u32
    return (size_t)2 * N * 8 + (size_t)2 * N * vb_words(N) * 8;
}

template <int WMAX>
__global__ void __launch_bounds__(VB_THREADS) track_validate_bits_kernel(const float* __restrict__ geom, int B, int N,
                                                                         uint8_t* __restrict__ valid)
{
    extern __shared__ __align__(16) unsigned char vb_raw[];
    __shared__ int bad_flag;
    const int b = blockIdx.x, V = 2 * N, W = vb_words(N);
    float* xs = reinterpret_cast<float*>(vb_raw);
    float* ys = xs + V;
    unsigned* neg = reinterpret_cast<unsigned*>(ys + V);
    unsigned* pos = neg + V * W;
    const float2* rec = reinterpret_cast<const float2*>(geom) + (size_t)b * 3 * N;
    if (threadIdx.x == 0) bad_flag = 0;
    for (int v = threadIdx.x; v < V; v += VB_THREADS) {
        const float2 pt = rec[v];
        xs[v] = pt.x;
        ys[v] = pt.y;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // phase 1: line i = (c[i], c[i+1 mod V]); right walls and the start line run from c[i+1] to c[i] in the reference.
    // A lane tests the same vertices (lane, lane + 32, ...) against every line: they stay in registers.
    float vx[WMAX], vy[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
        const int v = 32 * w + lane;
        vx[w] = v < V ? xs[v] : 0.f;
        vy[w] = v < V ? ys[v] : 0.f;
    }
    for (int i = warp; i < V; i += VB_THREADS / 32) {
        const int i1 = i + 1 == V ? 0 : i + 1;
        const bool rev = i < N;
        const float px = xs[rev ? i1 : i], py = ys[rev ? i1 : i], qx = xs[rev ? i : i1], qy = ys[rev ? i : i1];
        const float ax = xsub(qx, px), ay = xsub(qy, py);                      // q - p
        unsigned mine_n = 0, mine_p = 0;                                       // lane w keeps word w of the two rows
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
            const bool in = 32 * w + lane < V;
            const float val = det2(ay, xsub(vx[w], qx), ax, xsub(vy[w], qy));  // race.py:238 with r = c[32w + lane]
            const unsigned nb = __ballot_sync(FULL, in && val < 0.f), pb = __ballot_sync(FULL, in && val > 0.f);
            if (lane == w) { mine_n = nb; mine_p = pb; }
        }
        if (lane < W) { neg[i * W + lane] = mine_n; pos[i * W + lane] = mine_p; }
    }
    __syncthreads();
    // phase 2a: straddle rows (bit j of row i: line j has its end points strictly on different sides of line i),
    // written over the `neg` rows - every (row, word) is read and written by one thread, words w and w+1 of a row
    // are needed, so a row is handled by one thread
    for (int i = threadIdx.x; i < V; i += VB_THREADS) {
        unsigned* nr = neg + i * W;
        const unsigned* pr_ = pos + i * W;
        const unsigned n0 = nr[0], p0 = pr_[0];
        unsigned ncur = n0, pcur = p0;
        for (int w = 0; w < W; ++w) {
            const bool last = w == W - 1;
            const unsigned nnext = last ? n0 : nr[w + 1], pnext = last ? p0 : pr_[w + 1];
            // successor bits: vertex 32w+b+1; in the last word the successor of vertex V-1 is vertex 0
            unsigned ns = (ncur >> 1) | (nnext << 31), ps = (pcur >> 1) | (pnext << 31);
            if (last) {
                const int top = (V - 1) & 31;                                   // bit of vertex V-1 in the last word
                ns = (ncur >> 1) | ((n0 & 1u) << top);
                ps = (pcur >> 1) | ((p0 & 1u) << top);
            }
            nr[w] = (ncur & ps) | (pcur & ns);
            ncur = nnext;
            pcur = pnext;
        }
    }
    __syncthreads();
    // phase 2b: sparse - few lines straddle a given line
    for (int i = threadIdx.x; i < V; i += VB_THREADS) {
        for (int w = 0; w < W; ++w) {
            unsigned m = neg[i * W + w];
            while (m) {
                const int j = 32 * w + __ffs(m) - 1;
                m &= m - 1;
                if ((neg[j * W + (i >> 5)] >> (i & 31)) & 1u) bad_flag = 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) valid[b] = bad_flag ? 0 : 1;
}

// One warp per track: extent[b] = { max |point| over the record, max wall length of the polyline except the start line },
// both rounded UP (they are used as conservative bounds).  A record with a non-finite coordinate
// gets +inf / +inf, which makes every car of that track take the unpruned path.
__global__ void track_extent_kernel(const float* __restrict__ geom, int B, int N, float* __restrict__ extent)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = lane_id();
    const float2* rec = reinterpret_cast<const float2*>(geom) + (size_t)b * 3 * N;
    float r2 = 0.f, l2 = 0.f;
    bool finite = true;
    for (int j = lane; j < 3 * N; j += 32) {
        const float2 p = __ldg(&rec[j]);
        const float q = fmaf(p.x, p.x, p.y * p.y);
        finite = finite && (q <= 3e38f);                  // false for NaN and inf
        r2 = fmaxf(r2, q);
        if (j + 1 < 2 * N && j != N - 1) {                // wall j of the polyline; the start line (N-1) is not counted
            const float2 n = __ldg(&rec[j + 1]);
            const float dx = n.x - p.x, dy = n.y - p.y;
            l2 = fmaxf(l2, fmaf(dx, dx, dy * dy));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        r2 = fmaxf(r2, __shfl_xor_sync(FULL, r2, off));
        l2 = fmaxf(l2, __shfl_xor_sync(FULL, l2, off));
    }
    finite = __all_sync(FULL, finite);
    if (lane == 0) {
        const float up = 1.000001f;
        extent[2 * b] = finite ? __fsqrt_ru(r2) * up : INF;
        extent[2 * b + 1] = finite ? __fsqrt_ru(l2) * up : INF;
    }
}

}  // namespace glg

extern "C" int glg_track_extent(const float* geom, int32_t B, int32_t N, float* extent, glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && N >= 2 && N <= 512, "glg_track_extent: need B >= 0 and 2 <= N <= 512 (got B=%d N=%d)", B, N);
    GLG_REQUIRE((B == 0) || (geom && extent), "glg_track_extent: null pointer");
    if (B == 0) return GLG_OK;
    track_extent_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(geom, B, N, extent);
    return launch_status("glg_track_extent");
}

extern "C" int glg_track_build(const float* tracks, int32_t B, int32_t L, const float* sin_table,
                               const float* cos_table, int32_t table_half, float* geom,
                               glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && L >= 1 && L <= 510, "glg_track_build: need B >= 0 and 1 <= L <= 510 (got B=%d L=%d)", B, L);
    GLG_REQUIRE((B == 0) || (tracks && geom), "glg_track_build: null pointer");
    GLG_REQUIRE((sin_table == nullptr) == (cos_table == nullptr), "glg_track_build: give both tables or none");
    if (B == 0) return GLG_OK;
    const int N = L + 2;
    const size_t smem = (size_t)BUILD_WARPS * 5 * ((N + 1) & ~1) * sizeof(float);
    const int grid = (B + BUILD_WARPS - 1) / BUILD_WARPS;
    track_build_kernel<<<grid, BUILD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        tracks, nullptr, B, L, sin_table, cos_table, table_half, geom);
    return launch_status("glg_track_build");
}

extern "C" int glg_track_build_levels(const uint8_t* levels, int32_t B, int32_t L, const float* sin_table,
                                      const float* cos_table, int32_t table_half, float* geom,
                                      glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && L >= 1 && L <= 510, "glg_track_build_levels: need B >= 0 and 1 <= L <= 510 (got B=%d L=%d)", B, L);
    GLG_REQUIRE((B == 0) || (levels && geom), "glg_track_build_levels: null pointer");
    GLG_REQUIRE((sin_table == nullptr) == (cos_table == nullptr), "glg_track_build_levels: give both tables or none");
    if (B == 0) return GLG_OK;
    const int N = L + 2;
    const size_t smem = (size_t)BUILD_WARPS * 5 * ((N + 1) & ~1) * sizeof(float);
    const int grid = (B + BUILD_WARPS - 1) / BUILD_WARPS;
    track_build_kernel<<<grid, BUILD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        nullptr, levels, B, L, sin_table, cos_table, table_half, geom);
    return launch_status("glg_track_build_levels");
}

extern "C" int glg_track_validate(const float* geom, int32_t B, int32_t N, uint8_t* valid,
                                  glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && N >= 2 && N <= 512, "glg_track_validate: need B >= 0 and 2 <= N <= 512 (got B=%d N=%d)", B, N);
    GLG_REQUIRE((B == 0) || (geom && valid), "glg_track_validate: null pointer");
    if (B == 0) return GLG_OK;
    if (vb_smem(N) <= 48 * 1024) {                       // N <= 210: the sign-matrix kernel
        if (vb_words(N) <= 9)                            // N <= 144 (the reference's N = 130: 9 words per bit row)
            track_validate_bits_kernel<9><<<B, VB_THREADS, vb_smem(N), (cudaStream_t)stream>>>(geom, B, N, valid);
        else
            track_validate_bits_kernel<14><<<B, VB_THREADS, vb_smem(N), (cudaStream_t)stream>>>(geom, B, N, valid);
        return launch_status("glg_track_validate");
    }
    return glg_track_validate_pairs(geom, B, N, valid, stream);
}

extern "C" int glg_track_validate_pairs(const float* geom, int32_t B, int32_t N, uint8_t* valid,
                                        glg_stream_t stream)
{
    using namespace glg;
    GLG_REQUIRE(B >= 0 && N >= 2 && N <= 512, "glg_track_validate_pairs: need B >= 0 and 2 <= N <= 512 (got B=%d N=%d)", B, N);
    GLG_REQUIRE((B == 0) || (geom && valid), "glg_track_validate_pairs: null pointer");
    if (B == 0) return GLG_OK;
    const int M = 2 * (N - 1) + 2;
    int threads = ((M / 2 + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    track_validate_kernel<<<B, threads, (size_t)M * sizeof(float4), (cudaStream_t)stream>>>(geom, B, N, valid);
    return launch_status("glg_track_validate_pairs");
}
