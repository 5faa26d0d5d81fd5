/*
 * glg_b200.h - C ABI of the B200-native batched Race / Pacman environment kernels.
 *
 * This is the drop-in boundary of the hot path: every entry point takes
 * plain DEVICE pointers, extents and a CUDA stream; nothing here allocates, synchronises or
 * touches the host side of a tensor.  The caller (PyTorch, through ctypes - see
 * game_level_gan_b200/_lib.py and INTEGRATION.md) owns every buffer.
 *
 * All functions return GLG_OK (0) or a negative error code; glg_last_error() gives the text.
 * All functions are re-entrant; there is no global state besides the thread-local error string.
 *
 * File:line citations name the interface of the reference (Grzego/game-level-gan) that an entry
 * point replaces; they are relative to the reference's repository root.
 */
#ifndef GLG_B200_H
#define GLG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLG_OK            0
#define GLG_ERR_ARG      -1   /* bad argument (null pointer, extent out of range, misalignment) */
#define GLG_ERR_LAUNCH   -2   /* cudaGetLastError() after a launch was not cudaSuccess        */
#define GLG_ERR_UNSUPPORTED -3

#define GLG_MAX_PLAYERS   8
#define GLG_MAX_RAYS      32
#define GLG_ALIVE_SLOTS   1024 /* int32 slots of the "somebody is alive" step stamp (many, so that
                                 the atomic max of a launch does not pile up on a few addresses)  */

/* step kernel variants.  BRUTE is the literal reference loop.  The pruned variants (FAST, SCAN, PACKED and the fused
 * rollout) evaluate every (ray, wall) / (path, wall) pair they keep with BRUTE's arithmetic and drop only pairs that
 * cannot fire - with ONE stated caveat: the path/wall and path/finish-line tests are skipped when the two bounding
 * boxes, each widened by 1e-4, do not meet.  The reference's general case (o1 != o2 && o3 != o4, games/race.py:248)
 * has no box test, so for a path and a wall that are collinear to ~1e-5 with DISJOINT boxes, where the fp32
 * orientation signs are rounding noise, it could report a crossing that the pruned path does not.  No such case has
 * been observed: 0 mismatching car-steps in 2e9 fuzzed car-steps against BRUTE (tools/fuzz_pruned_vs_brute.py,
 * profiles/r02k_fuzz_pruned_vs_brute.json) and on every reference fixture; bench.py counts mismatches on every run. */
#define GLG_STEP_FAST     0   /* two-stage exact pruning of the ray cast, one warp per car (18 rays) */
#define GLG_STEP_BRUTE    1   /* every ray x every wall, the literal reference loop            */
#define GLG_STEP_SCAN     2   /* single-pass exact angular pruning (any even number of rays)   */
#define GLG_STEP_PACKED   3   /* production: GLG_STEP_FAST's algorithm with two cars per warp; falls back to
                                 GLG_STEP_FAST / SCAN / BRUTE for ray counts and track lengths it does not cover */

typedef void* glg_stream_t;   /* cudaStream_t */

const char* glg_last_error(void);
int glg_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Simulation constants of one Race instance.  Replaces the tensors built in
 * games/race.py:26-82 (Race.__init__: car tables, action tables, observation setup).
 *
 * The trigonometric entries are computed ON THE HOST WITH TORCH by the caller, using the same
 * ops the reference uses per step (games/race.py:310-324, 362-363, 462-466), so that headings and
 * ray directions are bit-identical to the reference; the kernels never call sinf/cosf per step.
 * Index f of the steering tables: 0 = straight, 1 = action_dirs +1 (right), 2 = action_dirs -1.
 * Index f of speed_inc: 0 = action_speed 0, 1 = +1 (forward), 2 = -3 (brake).
 * ------------------------------------------------------------------------------------------ */
typedef struct glg_race_params {
    int32_t num_players;                       /* P, 1..GLG_MAX_PLAYERS                        */
    int32_t num_rays;                          /* observation_size, 1..GLG_MAX_RAYS            */
    int32_t steps_limit;                       /* int(timeout // framerate), race.py:47        */
    float   max_distance;                      /* race.py:26 (10.)                             */
    float   step_penalty;                      /* race.py:74 (-0.01)                           */
    float   drag;                              /* race.py:346 (0.05)                           */
    float   progress_div;                      /* race.py:376: bounds.size(2) - 1 == 3         */
    float   vmax[GLG_MAX_PLAYERS];             /* race.py:34                                   */
    float   speed_inc[GLG_MAX_PLAYERS][3];     /* fl(fl(framerate*flag)*accel), race.py:367    */
    float   turn_cos[GLG_MAX_PLAYERS][3];      /* cos/sin of framerate*flag*angle, race.py:363 */
    float   turn_sin[GLG_MAX_PLAYERS][3];
    float   ray_cos[GLG_MAX_RAYS];             /* cos/sin of torch.linspace(...), race.py:462  */
    float   ray_sin[GLG_MAX_RAYS];
} glg_race_params;

/* Per-car state, structure of arrays, each array in the reference's own tensor layout
 * (games/race.py:182-190), so the host wrapper exposes them as Race.positions, .alive, ...  */
typedef struct glg_race_state {
    float*   positions;    /* [B,P,2] f32 */
    float*   directions;   /* [B,P,2] f32 */
    float*   speeds;       /* [B,P]   f32 */
    uint8_t* alive;        /* [B,P]   bool */
    uint8_t* finishes;     /* [B,P]   bool */
    int32_t* scores;       /* [B,P]   i32 */
} glg_race_state;

/* ------------------------------------------------------------------------------------------
 * Track geometry in HBM: one contiguous record per track,
 *     geom[b] = { right[N] REVERSED , left[N] , centre[N] }  each point (x,y) f32,   N = L + 2,
 * i.e. a [B,3,N,2] f32 tensor whose first 2N points form the polyline right-end .. start line ..
 * left-end (the order of the reference's own `line_bounds`, games/race.py:175), followed by the
 * centre points.  3*N*8 bytes per track = 3120 B at L = 128, 16-byte granular, so the step kernel
 * stages a record with ONE bulk async copy (TMA).  Walls, start and finish lines are consecutive
 * point pairs; the reference's per-player duplicated `bounds [B*P,2L+3,4]` (race.py:172-173) is
 * never materialised on the hot path.
 * ------------------------------------------------------------------------------------------ */

/* Generator output -> geometry.  Replaces games/race.py:126-158 (Race.reset, geometry part).
 *   tracks     [B,L,2] f32 (arc, width)
 *   sin_table / cos_table: optional [2*table_half+1] host-torch values of sin/cos(fl32(rad 8deg)*(n/4)),
 *              n = -table_half..table_half.  Tracks whose arcs are all exact multiples of 0.25
 *              (every generator-produced track) then get headings bit-identical to the reference;
 *              other tracks (or NULL tables) use sinf/cosf (<= 2 ulp from the reference's SLEEF).
 *   geom       [B,3,N,2] f32 out                                                             */
int glg_track_build(const float* tracks, int32_t B, int32_t L,
                    const float* sin_table, const float* cos_table, int32_t table_half,
                    float* geom, glg_stream_t stream);

/* The same geometry straight from the generator's discrete output.  The reference's
 * GeneratorNetworkConvDiscrete emits, per segment, one of the 9 arc levels linspace(-1, 1, 9) and width 0
 * (generators/race_track_generator.py:250-261); `levels` holds the level index 0..8 of every segment, two per
 * byte (segment s in bits 4*(s&1).. of byte s>>1 of the track's ceil(L/2) bytes): 64 B per track at L = 128
 * instead of 1 KB.  Result identical to glg_track_build on tracks[..., 0] = (level - 4) / 4, tracks[..., 1] = 0. */
int glg_track_build_levels(const uint8_t* levels, int32_t B, int32_t L,
                           const float* sin_table, const float* cos_table, int32_t table_half,
                           float* geom, glg_stream_t stream);

/* Track validity = no proper crossing among the 2(L+1)+2 lines (walls, start, finish).
 * Replaces Race._is_correct, games/race.py:326-334 (IMPL_GPU branch of reset, :199-200).
 *   valid [B] u8 out                                                                         */
int glg_track_validate(const float* geom, int32_t B, int32_t N, uint8_t* valid, glg_stream_t stream);
/* The same through the literal loop over all pairs of lines (the reference's own shape of the computation,
 * games/race.py:252-255 on every pair); glg_track_validate uses it for N > 210 and otherwise evaluates each
 * (line, end point) orientation once (neighbouring lines share end points) - identical results, kept as a cross-check. */
int glg_track_validate_pairs(const float* geom, int32_t B, int32_t N, uint8_t* valid, glg_stream_t stream);

/* Conservative per-track bounds the production step kernel prunes with (no reference counterpart):
 *   extent [B,2] f32 out = { max |point| over the record, longest wall of the polyline (the start
 *   line excepted) }, rounded up;
 *   +inf for a record with non-finite coordinates (its cars then take the unpruned path).      */
int glg_track_extent(const float* geom, int32_t B, int32_t N, float* extent, glg_stream_t stream);

/* Initial car state (games/race.py:182-190) and a cleared alive stamp.                        */
int glg_race_init(glg_race_state state, int32_t B, int32_t P, int32_t* alive_stamp, glg_stream_t stream);

/* One environment step for every car.  Replaces Race.step, games/race.py:340-500 (IMPL_GPU):
 * action masking, kinematics, progress, wall/finish collision, reward, score, drag, 18-ray
 * sensors, observation pack.
 *   actions    [P,B] i64 (values 0..8), not modified
 *   valid      [B] u8 (per track)
 *   extent     [B,2] f32 from glg_track_extent (required by GLG_STEP_PACKED / FAST, else may be NULL)
 *   step_no    value of Race.steps AFTER the increment of this step (race.py:349)
 *   states_out [P,B,num_rays+2] f32, rewards_out [P,B] f32
 *   alive_stamp [GLG_ALIVE_SLOTS] i32 or NULL: slot (b % GLG_ALIVE_SLOTS) := max(slot, launch_seq) if track b still
 *              has an alive car after this step (host reads it for Race.finished(), race.py:502-504)
 *   launch_seq a number the caller increases with every launch on this environment (> 0)
 *   base       NULL, or a device pointer to {step_no offset, launch_seq offset, last step} (3 x i32): the kernel
 *              adds the first two to step_no / launch_seq - a launch captured in a CUDA graph is replayed with
 *              the running numbers kept in memory - and does nothing if the resulting step number exceeds
 *              the third, so a graph may run past the end of an episode (steps_limit + 1 is the last step
 *              before Race.finished() reports the time-out, race.py:502-504; INT32_MAX = no limit, like
 *              Race.step itself).
 *   history    optional [>= step_no+1, P, 6] f32 ring written for track `record_id`
 *              (x, y, dx, dy, masked action, alive) at row step_no (race.py:492-494), or NULL
 *   variant    GLG_STEP_PACKED, GLG_STEP_FAST, GLG_STEP_SCAN or GLG_STEP_BRUTE (identical results)  */
int glg_race_step(const glg_race_params* params, const float* geom, int32_t B, int32_t N,
                  const int64_t* actions, const uint8_t* valid, const float* extent, glg_race_state state,
                  int32_t step_no, float* states_out, float* rewards_out,
                  int32_t* alive_stamp, int32_t launch_seq, const int32_t* base, float* history,
                  int32_t record_id, int32_t variant, glg_stream_t stream);

/* T consecutive steps with pre-computed actions [T,P,B] (random-action rollouts, replay): the loop of
 * train-gan.py:86-93 without a policy in it.  states_out/rewards_out hold the LAST step's outputs, or all steps if
 * `keep_all` (then [T,P,B,W] / [T,P,B]); every step's observation is computed either way.  first_step_no as in
 * glg_race_step; step t uses launch_seq = first_launch_seq + t.  history / record_id as in glg_race_step (row
 * first_step_no + t).
 *   mode   GLG_ROLLOUT_FUSED    ONE persistent launch plays all T steps: a warp keeps its track record in shared
 *                               memory and its cars' state in registers; per step only the action is read and the
 *                               observation and reward are written.  Configurations GLG_STEP_PACKED covers (18 rays,
 *                               N <= 256 even); others fall back to CHAINED (if `chain`) or STEPWISE.
 *          GLG_ROLLOUT_STEPWISE T step kernels back to back in plain stream order.
 *          GLG_ROLLOUT_CHAINED  T step kernels; launches 1..T-1 do not wait for the whole previous grid but, warp by
 *                               warp, for the previous step of their own car (chain[car] == launch_seq - 1, published
 *                               with release/acquire), so consecutive steps overlap.  With keep_all the stamp is
 *                               published right after the car state is written back (before the ray cast), without it
 *                               after the outputs (the steps share one output buffer); the production kernel with
 *                               keep_all hands the car state from launch to launch in six self-validating 64-bit words
 *                               {launch number : value} of `chain` (no fences, no second round trip) and only the
 *                               last launch writes the state arrays.
 *   chain  scratch of glg_race_chain_bytes(B, P) bytes (16-byte aligned, zero-filled once), required by CHAINED,
 *          else may be NULL.  The numbers first_launch_seq .. first_launch_seq+T-1 must be larger than anything
 *          stored in `chain` before.
 * All modes give identical results.                                                              */
#define GLG_ROLLOUT_STEPWISE 0
#define GLG_ROLLOUT_CHAINED  1
#define GLG_ROLLOUT_FUSED    2
int glg_race_rollout(const glg_race_params* params, const float* geom, int32_t B, int32_t N,
                     const int64_t* actions, int32_t T, const uint8_t* valid, const float* extent,
                     glg_race_state state, int32_t first_step_no, float* states_out, float* rewards_out, int32_t keep_all,
                     int32_t* alive_stamp, int32_t first_launch_seq, int32_t* chain,
                     float* history, int32_t record_id, int32_t mode,
                     int32_t variant, glg_stream_t stream);

int64_t glg_race_chain_bytes(int32_t B, int32_t P);

/* Winner per track.  Replaces Race.winners, games/race.py:506-529.  winners [B] i64 out.      */
int glg_race_winners(const int32_t* scores, const uint8_t* finishes, const uint8_t* valid,
                     int32_t B, int32_t P, int32_t steps_limit, int64_t* winners, glg_stream_t stream);

/* Winner statistics for the winner discriminator (train-gan.py:103-104): tracks are laid out
 * trial-major [trials, boards]; out[board, c] = mean over trials of one_hot(winner+1, P+1).
 *   winners [trials*boards] i64, out [boards, P+1] f32                                        */
int glg_winner_stats(const int64_t* winners, int32_t trials, int32_t boards, int32_t P,
                     float* out, glg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stateless helpers.  Replace the three free functions of the pybind module `game_helpers`
 * (games/game_helpers.cpp:335-371, 373-418, 421-450; bound at :455-457).  The reference backs
 * them with Boost.Geometry (not vendored, unpinned); semantics here: intersection including
 * touches, Euclidean distance from the ray origin to the nearest intersection point (+inf if
 * none), validity = the polyline does not intersect itself.
 * ------------------------------------------------------------------------------------------ */
/* tracks [b,s,2] f32 polyline, segments [b,p,4] f32, out [b,p] u8 */
int glg_collision(const float* tracks, const float* segments, uint8_t* out,
                  int32_t b, int32_t s, int32_t p, glg_stream_t stream);
/* tracks [b,s,2], directions [b,d,4] = (x,y,dx,dy), out [b,d] f32 */
int glg_smallest_distance(const float* tracks, const float* directions, float* out,
                          int32_t b, int32_t s, int32_t d, glg_stream_t stream);
/* tracks [b,s,2], out [b] u8 */
int glg_is_valid(const float* tracks, uint8_t* out, int32_t b, int32_t s, glg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stateful helper.  Replaces class `Game` of the pybind module (games/game_helpers.cpp:158-327,
 * bound at :459-463): per-player cell tracking along the track.
 * The handle is a small host object; all device storage lives in a caller-provided workspace of
 * glg_game_workspace_bytes(b, s, num_players) bytes (256-byte aligned device memory).
 * ------------------------------------------------------------------------------------------ */
typedef struct glg_game glg_game;

int64_t glg_game_workspace_bytes(int32_t b, int32_t s, int32_t num_players);
/* left,right [b,s,2] f32 device.  Copies them into the workspace in the order of
 * create_racetracks (game_helpers.cpp:109-144) and initialises players (:146-156, 160-174).   */
int glg_game_create(glg_game** out, void* workspace, int64_t workspace_bytes,
                    const float* left, const float* right, int32_t b, int32_t s,
                    int32_t num_players, glg_stream_t stream);
void glg_game_destroy(glg_game* game);
/* game_helpers.cpp:176-189.  valid [b] u8 out. */
int glg_game_validate_tracks(glg_game* game, uint8_t* valid, glg_stream_t stream);
/* game_helpers.cpp:191-279.  idx [k] i64, new_positions [k,row_stride] f32 (columns 0,1 used),
 * dead/finished [k] u8 out.  Mutates the players' position/cell.                              */
int glg_game_update_players(glg_game* game, const int64_t* idx, const float* new_positions,
                            int32_t k, int32_t row_stride, uint8_t* dead, uint8_t* finished,
                            glg_stream_t stream);
/* game_helpers.cpp:281-322.  idx [k] i64, directions [k,d,4] f32, out [k,d] f32.               */
int glg_game_smallest_distance(glg_game* game, const int64_t* idx, const float* directions,
                               int32_t k, int32_t d, float* out, glg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Pacman grid environment.  Replaces Pacman.step / observation build, games/pacman.py:64-109.
 *   grid    [B,H,W,4+P] i32 (the 4+P "real" channels of the reference's [B,H,W,4+2P] grid)
 *   players [B*P,4] i32 rows (b,x,y,p) in the reference's np.where order (pacman.py:59-61)
 *   actions [B,P] i32 (0..4), rewards [P,B] f64 out (the reference returns float64 rewards)
 *   scratch [1] i32 device word (the batch-wide "somebody can move" flag of pacman.py:77)
 * ------------------------------------------------------------------------------------------ */
int glg_pacman_step(int32_t* grid, int32_t* players, const int32_t* actions, double* rewards,
                    int32_t* scratch, int32_t B, int32_t H, int32_t W, int32_t P, glg_stream_t stream);
/* obs [P,B,H,W,4+2P] f32 out: grid channels, zeros, and 1.0 in channel 4+P+p (pacman.py:107-109) */
int glg_pacman_observe(const int32_t* grid, float* obs, int32_t B, int32_t H, int32_t W, int32_t P,
                       glg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GLG_B200_H */
