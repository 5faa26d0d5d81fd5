/*
 * race_oracle.c - plain C restatement of the reference's Race environment (torch IMPL_GPU path).
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/librace_oracle.so by oracle/Makefile and
 * loaded (ctypes) by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product never links or calls it.
 *
 * Parity status: PINNED - tests/test_oracle_c.py replays the reference-generated fixtures in
 * tests/golden/ through this file and requires bit-identical state, observations and rewards
 * (geometry for tracks whose arcs are multiples of 0.25, where the sin/cos table applies).
 *
 * Scalar fp32 model of the ATen CPU kernels the reference runs (re-verified):
 *   - a*b - c*d      : three roundings (mul, mul, sub), never fused     -> compile with -ffp-contract=off
 *   - v @ R (2x2)    : x*c + y*s, x*(-s) + y*c, three roundings each
 *   - norm(dim=-1)   : sqrtf(fmaf(y, y, x*x))
 *   - cumsum         : running sum in double, each prefix rounded to float
 * sin/cos of per-step angles come from host tables (glg_race_params); sin/cos of the track headings
 * come from the optional table (exact for arcs in 0.25 steps) or libm (last-ulp differences).
 *
 * All citations are file:line of /root/reference/games/race.py unless stated otherwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/glg_b200.h"

static inline int sgn(float v) { return (v > 0.f) - (v < 0.f); }

/* :230-238  orientation of r with respect to p->q */
static inline int turn(float px, float py, float qx, float qy, float rx, float ry)
{
    float ax = qx - px, ay = qy - py;
    float bx = rx - qx, by = ry - qy;
    float m1 = ay * bx;
    float m2 = ax * by;
    return sgn(m1 - m2);
}

/* :240-245  r inside the bounding box of p,q */
static inline int in_box(float px, float py, float qx, float qy, float rx, float ry)
{
    return rx <= fmaxf(px, qx) && rx >= fminf(px, qx) && ry <= fmaxf(py, qy) && ry >= fminf(py, qy);
}

/* :248-269  wall (p,q) against probe (a,b); *start_on = probe start lies on the wall */
static inline int cross_general(float px, float py, float qx, float qy,
                                float ax, float ay, float bx, float by, int* start_on)
{
    int o1 = turn(px, py, qx, qy, ax, ay);
    int o2 = turn(px, py, qx, qy, bx, by);
    int o3 = turn(ax, ay, bx, by, px, py);
    int o4 = turn(ax, ay, bx, by, qx, qy);
    int hit = (o1 != o2) && (o3 != o4);
    *start_on = (o1 == 0) && in_box(px, py, qx, qy, ax, ay);
    hit |= (o2 == 0) && in_box(px, py, qx, qy, bx, by);
    hit |= (o3 == 0) && in_box(ax, ay, bx, by, px, py);
    hit |= (o4 == 0) && in_box(ax, ay, bx, by, qx, qy);
    return hit;
}

/* wall j of track record g (N points per polyline): order right 0..N-2, left 0..N-2, start (:166-172);
 * j == 2(N-1)+1 is the finish line (:169). */
static inline void wall_of(const float* g, int N, int j, float* w)
{
    const float* right = g;
    const float* left = g + 2 * N;
    int S = N - 1;
    if (j < S) { memcpy(w, right + 2 * j, 4 * sizeof(float)); }
    else if (j < 2 * S) { memcpy(w, left + 2 * (j - S), 4 * sizeof(float)); }
    else if (j == 2 * S) { w[0] = left[0]; w[1] = left[1]; w[2] = right[0]; w[3] = right[1]; }
    else { w[0] = left[2 * (N - 1)]; w[1] = left[2 * (N - 1) + 1];
           w[2] = right[2 * (N - 1)]; w[3] = right[2 * (N - 1) + 1]; }
}

/* ---------------------------------------------------------------------------------------- */
/* :126-158 geometry.  geom [B,3,N,2] = right, left, centre.  Returns 0.                     */
int ro_track_build(const float* tracks, int B, int L, const float* sin_table,
                   const float* cos_table, int table_half, float* geom)
{
    const int N = L + 2;
    const float rad = (float)0.13962634015954636;   /* math.radians(8.) as fp32 scalar, :140 */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float* t = tracks + (size_t)b * L * 2;
        float* right = geom + (size_t)b * 6 * N;
        float* left = right + 2 * N;
        float* centre = left + 2 * N;
        float* sx = (float*)malloc(sizeof(float) * 4 * N);
        float* sy = sx + N; float* arc = sy + N; float* wid = arc + N;
        int quant = sin_table != NULL;
        for (int j = 0; j < N; ++j) {                 /* :136-138 sentinels */
            arc[j] = (j >= 1 && j <= L) ? t[2 * (j - 1)] : 0.f;
            wid[j] = (j >= 1 && j <= L) ? t[2 * (j - 1) + 1] : 0.f;
            float a4 = arc[j] * 4.f;
            if (a4 != rintf(a4) || fabsf(a4) > 4.f) quant = 0;
        }
        double acc = 0.0;
        for (int j = 0; j < N; ++j) {                 /* :140-142 */
            acc += (double)arc[j];
            float h = (float)acc;
            float ang = rad * h;
            float s, c;
            int n = (int)(h * 4.f);
            if (quant && n >= -table_half && n <= table_half) { s = sin_table[n + table_half]; c = cos_table[n + table_half]; }
            else { s = sinf(ang); c = cosf(ang); }
            sx[j] = s * 0.2f;
            sy[j] = c * 0.2f;
        }
        double cx = 0.0, cy = 0.0;
        for (int j = 0; j < N; ++j) {                 /* :154-156 exclusive cumsum */
            centre[2 * j] = (j == 0) ? 0.f : (float)cx;
            centre[2 * j + 1] = (j == 0) ? 0.f : (float)cy;
            cx += (double)sx[j];
            cy += (double)sy[j];
        }
        for (int j = 0; j < N; ++j) {                 /* :144-152, 157-158 */
            float ox, oy;
            if (j == 0) { ox = 0.5f; oy = 0.f; }
            else {
                float nx = sy[j] + sy[j - 1];         /* perp = (y, -x) */
                float ny = (-sx[j]) + (-sx[j - 1]);
                float len = sqrtf(fmaf(ny, ny, nx * nx));
                float w = 0.5f + 1.5f * wid[j - 1];
                ox = (nx / len) * w;
                oy = (ny / len) * w;
            }
            right[2 * j] = centre[2 * j] + ox;
            right[2 * j + 1] = centre[2 * j + 1] + oy;
            left[2 * j] = centre[2 * j] + (-ox);
            left[2 * j + 1] = centre[2 * j + 1] + (-oy);
        }
        free(sx);
    }
    return 0;
}

/* :326-334 validity over walls + start + finish (2(N-1)+2 lines) */
int ro_track_validate(const float* geom, int B, int N, uint8_t* valid)
{
    const int M = 2 * (N - 1) + 2;
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
        const float* g = geom + (size_t)b * 6 * N;
        float* w = (float*)malloc(sizeof(float) * 4 * M);
        for (int j = 0; j < M; ++j) wall_of(g, N, j, w + 4 * j);
        int ok = 1;
        for (int i = 0; i < M && ok; ++i) {
            const float* a = w + 4 * i;
            for (int j = i + 1; j < M; ++j) {
                const float* c = w + 4 * j;
                int o1 = turn(a[0], a[1], a[2], a[3], c[0], c[1]);
                int o2 = turn(a[0], a[1], a[2], a[3], c[2], c[3]);
                int o3 = turn(c[0], c[1], c[2], c[3], a[0], a[1]);
                int o4 = turn(c[0], c[1], c[2], c[3], a[2], a[3]);
                if (o1 * o2 < 0 && o3 * o4 < 0) { ok = 0; break; }
            }
        }
        valid[b] = (uint8_t)ok;
        free(w);
    }
    return 0;
}

/* :182-190 */
int ro_race_init(glg_race_state st, int B, int P)
{
    for (int k = 0; k < B * P; ++k) {
        st.positions[2 * k] = 0.f; st.positions[2 * k + 1] = 0.1f;
        st.directions[2 * k] = 0.f; st.directions[2 * k + 1] = 1.f;
        st.speeds[k] = 0.f; st.alive[k] = 1; st.finishes[k] = 0; st.scores[k] = 0;
    }
    return 0;
}

/* :340-500 one step for all cars.  Returns the number of cars alive afterwards.
 * The "nobody alive" early-out (:353-356) is the caller's job, as in the CUDA path.          */
int ro_race_step(const glg_race_params* pr, const float* geom, int B, int N,
                 const int64_t* actions, const uint8_t* valid, glg_race_state st, int step_no,
                 float* states_out, float* rewards_out)
{
    static const int thr_idx[9] = {0, 1, 2, 0, 1, 2, 0, 1, 2};   /* action_speed 0,+1,-3  (:52-61) */
    static const int str_idx[9] = {0, 0, 0, 1, 1, 1, 2, 2, 2};   /* action_dirs 0,+1,-1   (:62-71) */
    const int P = pr->num_players, O = pr->num_rays, W = O + 2;
    const int S2 = 2 * (N - 1) + 1;                                /* walls incl. start */
    int alive_after = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : alive_after)
    for (int k = 0; k < B * P; ++k) {
        const int b = k / P, p = k % P;
        const float* g = geom + (size_t)b * 6 * N;
        const float* centre = g + 4 * N;
        int alive = st.alive[k], fin = st.finishes[k], ok = valid[b];
        int a = (int)actions[(size_t)p * B + b];
        if (!alive || !ok) a = 0;                                  /* :359 */
        const int fs = str_idx[a], ft = thr_idx[a];
        /* :362-364 heading */
        float dx = st.directions[2 * k], dy = st.directions[2 * k + 1];
        float c = pr->turn_cos[p][fs], s = pr->turn_sin[p][fs];
        float ndx = dx * c + dy * s;
        float ndy = dx * (-s) + dy * c;
        /* :366-370 speed */
        float v = st.speeds[k] + pr->speed_inc[p][ft];
        float nv = fminf(pr->vmax[p], fmaxf(v, 0.f));
        int moving = fabsf(nv) > 1e-7f;
        /* :372 */
        float px = st.positions[2 * k], py = st.positions[2 * k + 1];
        float nx = px + ndx * nv, ny = py + ndy * nv;
        /* :374-376 progress */
        int idx = 0; float best = INFINITY;
        for (int j = 0; j < N; ++j) {
            float ex = nx - centre[2 * j], ey = ny - centre[2 * j + 1];
            float d = sqrtf(fmaf(ey, ey, ex * ex));
            if (d < best) { best = d; idx = j; }
        }
        float reward = fin ? 0.f : pr->step_penalty;               /* :382-383 */
        int upd = alive && moving && ok;                           /* :380 */
        if (upd) {
            int dead = 0, done = 0, so;
            float w[4];
            for (int j = 0; j < S2 && !dead; ++j) {                /* :406-407 */
                wall_of(g, N, j, w);
                dead = cross_general(w[0], w[1], w[2], w[3], px, py, nx, ny, &so) | so;
            }
            wall_of(g, N, S2, w);                                  /* :431-432 finish line */
            done = cross_general(w[0], w[1], w[2], w[3], px, py, nx, ny, &so) | so;
            reward = reward + ((float)done - (float)dead);         /* :434 */
            alive = alive && !dead && !done;
            fin = fin || done;
            int sc = st.scores[k];
            if (dead) sc = idx + pr->steps_limit + 1;              /* :442-444 */
            if (done) sc = step_no;                                /* :446-447 */
            st.scores[k] = sc;
        }
        if (!alive) nv = 0.f;                                      /* :449 */
        float drag = 1.f - (1.f - (ft != 0 ? 1.f : 0.f)) * pr->drag;   /* :452 */
        float speed = nv * drag;
        st.directions[2 * k] = ndx; st.directions[2 * k + 1] = ndy;
        st.speeds[k] = speed;
        st.positions[2 * k] = nx; st.positions[2 * k + 1] = ny;
        st.alive[k] = (uint8_t)alive; st.finishes[k] = (uint8_t)fin;
        alive_after += alive;
        /* :459-489 sensors */
        float* out = states_out + ((size_t)p * B + b) * W;
        for (int i = 0; i < O; ++i) {
            float obs = 0.f;
            if (alive) {
                float rc = pr->ray_cos[i], rs = pr->ray_sin[i];
                float rdx = ndx * rc + ndy * rs;
                float rdy = ndx * (-rs) + ndy * rc;
                float fx = nx + 1000.f * rdx, fy = ny + 1000.f * rdy;     /* :289 */
                float tmin = INFINITY; int nan = 0;
                float w[4];
                for (int j = 0; j < S2; ++j) {
                    wall_of(g, N, j, w);
                    int so;
                    int hit = cross_general(w[0], w[1], w[2], w[3], nx, ny, fx, fy, &so);
                    hit = hit && !so;                                      /* :292 */
                    float wqx = w[2] - w[0], wqy = w[3] - w[1];
                    float psx = w[0] - nx, psy = w[1] - ny;
                    float m1 = psy * wqx, m2 = psx * wqy;
                    float num = m1 - m2;                                   /* :300 */
                    float m3 = rdy * wqx, m4 = rdx * wqy;
                    float den = m3 - m4;                                   /* :301 */
                    float t;
                    if (so) t = 0.f;                                       /* :303 */
                    else if (hit) t = num / den;                           /* :304 */
                    else t = INFINITY;                                     /* :305 */
                    if (t < 0.f) t = INFINITY;                             /* :306 */
                    if (t != t) nan = 1;
                    else if (t < tmin) tmin = t;                           /* :308 */
                }
                if (nan) tmin = NAN;
                obs = fminf(tmin, pr->max_distance);                       /* :489 clamp(max) */
                if (nan) obs = NAN;
                obs = obs / pr->max_distance;
            }
            out[i] = obs;
        }
        out[O] = speed / pr->vmax[p];                              /* :497 */
        out[O + 1] = (float)idx / pr->progress_div;                /* :376, 498 */
        rewards_out[(size_t)p * B + b] = reward;
    }
    return alive_after;
}

/* :506-529 */
int ro_race_winners(const int32_t* scores, const uint8_t* finishes, const uint8_t* valid,
                    int B, int P, int steps_limit, int64_t* winners)
{
    for (int b = 0; b < B; ++b) {
        int anyf = 0;
        for (int p = 0; p < P; ++p) anyf |= finishes[b * P + p];
        int best = 0;
        if (anyf) {                                                /* argmin, first index */
            int bv = 0;
            for (int p = 0; p < P; ++p) {
                int v = finishes[b * P + p] ? scores[b * P + p] : steps_limit + 1;
                if (p == 0 || v < bv) { bv = v; best = p; }
            }
        } else {                                                   /* argmax, first index */
            int bv = 0;
            for (int p = 0; p < P; ++p) {
                int v = scores[b * P + p];
                if (p == 0 || v > bv) { bv = v; best = p; }
            }
        }
        winners[b] = valid[b] ? best : -1;
    }
    return 0;
}

int ro_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
