/* ncg_b200.h -- C ABI of the B200-native batched CarEnv stepping engine.
 *
 * Drop-in boundary for ONE path of heihachi78/NascarGymnasium: CarEnv.reset()/step()
 * for E environments of C cars (1..10), replacing, per call:
 *
 *   ncg_create / ncg_upload_tracks   CarEnv.__init__ + CarPhysics.__init__ + wall construction
 *                                    (/root/reference/src/car_env.py:79-241, src/car_physics.py:74-339)
 *   ncg_reset                        CarEnv.reset                  (src/car_env.py:316-535)
 *   ncg_step                         CarEnv.step -> _step_multi_car (src/car_env.py:678-803), i.e.
 *                                    CarPhysics.step (src/car_physics.py:341-384), Car.update_physics
 *                                    (src/car.py:329-387), TyreManager.update (src/tyre_manager.py:78),
 *                                    b2World.Step/RayCast (box2d-py 2.3.8), LapTimer.update
 *                                    (src/lap_timer.py:95), DistanceSensor.get_sensor_distances
 *                                    (src/distance_sensor.py:71), reward/termination (src/car_env.py:980-1158)
 *   ncg_rollout                      the demo/random_demo.py loop shape (random actions, T steps) kept on device
 *   ncg_get_state / ncg_set_state    (no reference equivalent: Box2D state is not serialisable; used for
 *                                    teacher-forced parity tests and env checkpoints)
 *   (info dict)                      built lazily on the host from ncg_get_state records (src/car_env.py:1160-1227)
 *
 * Conventions: every pointer named d_* is a DEVICE pointer owned by the caller (PyTorch tensor
 * storage), contiguous, on the handle's device; h_* are HOST pointers.  All work is enqueued on the
 * `stream` argument (a cudaStream_t passed as void*); nothing synchronises implicitly except the
 * host-buffer calls (ncg_step_host / ncg_step_pinned / ncg_step_mapped / ncg_reset_host), which return
 * when the results are in the host buffers.  Every call returns 0 on success or a negative NCG_E_* code;
 * ncg_last_error() returns the message for the calling thread.  A handle is not re-entrant.
 * There is no CPU fallback: ncg_create fails if no CUDA device is usable.
 */
#ifndef NCG_B200_H
#define NCG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCG_OBS_DIM 38
#define NCG_NUM_SENSORS 16
#define NCG_MAX_CARS 10
#define NCG_RECORD_WORDS 128   /* 32-bit words of persistent state per car */
#define NCG_MAX_CONTACTS 12    /* broad-phase contacts kept per car (Box2D keeps an unbounded list) */
#define NCG_MAX_TOUCHING 4     /* contacts with manifold points kept per car */
#define NCG_MAX_ACTIVE 4       /* CarCollisionListener.active_collisions entries kept per car */
#define NCG_VEL_HISTORY 600    /* Car.velocity_history: deque(maxlen=VELOCITY_HISTORY_SIZE), src/constants/environment.py:23 */

enum {
    NCG_OK = 0,
    NCG_E_INVALID = -1,   /* bad argument (ValueError in the Python mirror) */
    NCG_E_CUDA = -2,      /* CUDA runtime error */
    NCG_E_STATE = -3,     /* call out of order, e.g. step before reset (RuntimeError) */
    NCG_E_NOMEM = -4
};

/* Per-car state record: NCG_RECORD_WORDS 32-bit words, car-major (record[car][word]).  Words are
 * float32 unless marked u32.  A CTA moves its (at most 32) records between HBM and shared memory as
 * coalesced 16-byte accesses and keeps them there for the whole launch.  Field list follows SURVEY.md App. C. */
enum NcgRecordField {
    NCG_R_X = 0, NCG_R_Y, NCG_R_ANGLE, NCG_R_VX, NCG_R_VY, NCG_R_OMEGA,   /* b2Body sweep.c, sweep.a, velocity */
    NCG_R_SLEEP,              /* b2Body::m_sleepTime */
    NCG_R_FLAGS,              /* u32 NCG_F_* bits */
    NCG_R_FAT_LX, NCG_R_FAT_LY, NCG_R_FAT_UX, NCG_R_FAT_UY,               /* car proxy fat AABB */
    NCG_R_INV_DT0,            /* b2World::m_inv_dt0 */
    NCG_R_NCONTACT,           /* u32: bits 0-7 contacts, 8-15 active collisions, 16-27 touching mask (by contact slot) */
    NCG_R_CONTACT_WALL,       /* u32 x6: NCG_MAX_CONTACTS wall indices, 16 bit each, slot 0 = newest */
    NCG_R_MANIFOLD_PC = NCG_R_CONTACT_WALL + 6,   /* u32: point count (2 bits each) of the k-th touching contact */
    NCG_R_MANIFOLD,           /* NCG_MAX_TOUCHING x {key0 u32, key1 u32, ni0, ti0, ni1, ti1} */
    NCG_R_IMPULSE = NCG_R_MANIFOLD + 6 * NCG_MAX_TOUCHING,   /* listener max normal impulse this step */
    NCG_R_ACTIVE,             /* NCG_MAX_ACTIVE x {wall u32, nx, ny} in insertion order */
    NCG_R_RPM = NCG_R_ACTIVE + 3 * NCG_MAX_ACTIVE,
    NCG_R_PREV_VX, NCG_R_PREV_VY,
    NCG_R_ACC_N,              /* u32: samples held in the 10-deep acceleration window (oldest first) */
    NCG_R_ACC,                /* 10 x {long, lat} */
    NCG_R_SLIP = NCG_R_ACC + 20, NCG_R_FLAT, NCG_R_BANK,
    NCG_R_TYRE_TEMP, NCG_R_TYRE_WEAR = NCG_R_TYRE_TEMP + 4, NCG_R_TYRE_LOAD = NCG_R_TYRE_WEAR + 4,
    NCG_R_CUM_IMPACT = NCG_R_TYRE_LOAD + 4,
    NCG_R_STUCK_STEPS,        /* u32: consecutive steps below 0.5 m/s */
    NCG_R_STUCK_X, NCG_R_STUCK_Y,
    NCG_R_BACK, NCG_R_BACK_PREV, NCG_R_PROGRESS_PREV,
    NCG_R_PREV_X, NCG_R_PREV_Y,   /* reward odometer anchor (_previous_car_position) */
    NCG_R_CUM_REWARD,
    NCG_R_LAP_START,          /* u32: step index of current_lap_start_time */
    NCG_R_LAST_LAP, NCG_R_BEST_LAP,   /* seconds */
    NCG_R_LAP_COUNT,          /* u32 */
    NCG_R_ODO,
    NCG_R_LAP_X, NCG_R_LAP_Y, /* LapTimer.last_car_position */
    NCG_R_STEP,               /* u32: env steps since reset (simulation_time = sum of STEP additions of 1/60) */
    NCG_R_TRACK,              /* u32: track id of the env */
    NCG_R_MAX_SPEED,          /* info: max speed seen this episode */
    NCG_R_USED
};

enum NcgFlagBits {
    NCG_F_AWAKE = 1u << 0,
    NCG_F_PROXY_MOVED = 1u << 1,
    NCG_F_NEW_FIXTURE = 1u << 2,
    NCG_F_HAS_KEY = 1u << 3,        /* car_collision_impulses has this car's key */
    NCG_F_DISABLED = 1u << 4,
    NCG_F_FIRST_STEP = 1u << 5,
    NCG_F_STUCK_POS = 1u << 6,
    NCG_F_BACK_ACTIVE = 1u << 7,
    NCG_F_CROSSED = 1u << 8,        /* has_crossed_startline (== is_timing) */
    NCG_F_HAS_LAST = 1u << 9,
    NCG_F_HAS_BEST = 1u << 10,
    NCG_F_HAS_POS = 1u << 11,       /* LapTimer.last_car_position is not None */
    NCG_F_OVERFLOW = 1u << 12,      /* a per-car cap (contacts/touching/active) was exceeded at least once */
    NCG_F_ON_TRACK = 1u << 13       /* info: is_car_on_track() after the last step */
};

typedef struct NcgConfig {
    int32_t device;          /* CUDA device ordinal */
    int32_t num_envs;        /* E */
    int32_t cars_per_env;    /* C in 1..NCG_MAX_CARS */
    int32_t discrete;        /* 0: actions float32 (E,C,2) in [-1,1]; 1: int32 (E,C) in 0..4 */
    int32_t reset_on_lap;    /* CarEnv(reset_on_lap=...) */
    int32_t auto_reset;      /* 1: envs that finish are reset inside the same step (VecEnv semantics) */
    int32_t contacts;        /* 1: car-wall contact solver + TOI enabled (default), b2CollidePolygons as in Box2D 2.3.1+;
                                2: the same with b2CollidePolygons as in Box2D 2.3.0 (edge-walk b2FindMaxSeparation,
                                k_relativeTol/k_absoluteTol); 0: contact-free integrator */
    int32_t track_info;      /* 1: also keep what only the info dict needs: is_car_on_track() every step (NCG_F_ON_TRACK)
                                and the 600-sample velocity history behind validate_performance (src/car.py:1060-1098) */
    float start_x, start_y;  /* CarEnv(start_position=...) (src/car_env.py:114, 391, 398); (0, 0) = the GRID segment's start */
    float start_angle;       /* CarEnv(start_angle=...), radians */
    int32_t car_contacts;    /* 0 (default): every car has its own world, as in the reference (src/car_env.py:389-394).  1: the
                                cars of an env share ONE Box2D world and collide with each other (no reference
                                behaviour; Box2D semantics, restated in oracle/b2lite.h SharedWorld) */
    float grid_dx, grid_dy;  /* car_contacts only: start grid, car k at (-(k / 2) * grid_dx, +-grid_dy / 2) in the start frame
                                (e.g. 8 m, 3 m; the fat AABBs of neighbours must not overlap at rest) */
} NcgConfig;

typedef struct NcgHandle NcgHandle;

/* Counters accumulated on the device by ncg_step / ncg_rollout since the last ncg_read_stats(reset=1). */
typedef struct NcgStats {
    uint64_t car_steps;          /* cars advanced (disabled cars included, as the reference steps them) */
    uint64_t episodes;           /* envs that terminated or truncated */
    uint64_t laps;               /* laps completed */
    uint64_t ray_tests;          /* ray-vs-wall box tests */
    uint64_t contact_steps;      /* car-steps with at least one touching contact */
    uint64_t toi_events;         /* TOI sub-steps solved */
    uint64_t overflow;           /* car-steps that hit a per-car cap */
    double   return_sum;         /* sum of per-car episode returns of finished episodes */
} NcgStats;

const char* ncg_last_error(void);
int ncg_version(void);

int ncg_create(const NcgConfig* cfg, NcgHandle** out);
int ncg_destroy(NcgHandle* h);

/* Upload n_tracks track tables built by nascargymnasium_b200/track.py (one float32 blob per track,
 * concatenated; h_offsets[n_tracks+1] are word offsets, each a multiple of 4 words). */
int ncg_upload_tracks(NcgHandle* h, const float* h_blob, const int64_t* h_offsets, int32_t n_tracks);

/* Reset envs whose d_env_mask byte is non-zero (NULL = all).  d_track_id (int32[E], NULL = keep) selects
 * each reset env's track (the ids are copied back and range-checked before anything is launched: one synchronisation
 * of `stream`; an id outside [0, n_tracks) fails with NCG_E_INVALID and changes nothing).  fresh=1 is a brand-new Box2D world (first reset / track change); fresh=0 is
 * CarPhysics.reset_car on the existing world.  Writes the initial observations of reset envs to d_obs
 * (float32[E*C*38], may be NULL). */
int ncg_reset(NcgHandle* h, const uint8_t* d_env_mask, const int32_t* d_track_id, int32_t fresh, float* d_obs,
              void* stream);

/* One CarEnv.step for every env.  d_actions: float32[E*C*2] or int32[E*C].  Outputs: d_obs float32[E*C*38],
 * d_reward float32[E*C], d_terminated/d_truncated uint8[E].  With auto_reset, finished envs are reset
 * (reset_car semantics) in the same call, d_obs holds the post-reset observation and d_final_obs
 * (float32[E*C*38], may be NULL) receives the terminal observation of finished envs. */
int ncg_step(NcgHandle* h, const void* d_actions, float* d_obs, float* d_reward, uint8_t* d_terminated,
             uint8_t* d_truncated, float* d_final_obs, void* stream);

/* T steps with actions drawn on the device (Philox4x32-10, key=seed, counter=(car, step)):
 * mode 0 = action_space.sample() (continuous U[-1,1]^2 / discrete U{0..4}); mode 1 = "driving"
 * distribution tb~U[0.2,1], steer~U[-0.2,0.6].  If d_obs_rollout is non-NULL it receives every step's
 * observations, float32[T][E*C][38]; d_reward_rollout float32[T][E*C], d_done_rollout uint8[T][E] likewise.
 * Otherwise only the last step's observation is written to d_obs_last (may be NULL).  Auto-reset is forced on. */
int ncg_rollout(NcgHandle* h, int32_t steps, uint64_t seed, int32_t mode, float* d_obs_rollout,
                float* d_reward_rollout, uint8_t* d_done_rollout, float* d_obs_last, void* stream);

/* Host-buffer convenience used by the CarEnv mirror and the e2e benchmark: copies actions host->device,
 * steps, copies results device->host through pinned staging, returns after the results are readable. */
int ncg_step_host(NcgHandle* h, const void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated,
                  uint8_t* h_truncated, float* h_final_obs);
int ncg_reset_host(NcgHandle* h, const uint8_t* h_env_mask, const int32_t* h_track_id, int32_t fresh, float* h_obs);

/* Zero-copy variant of ncg_step_host: ncg_host_buffers hands out the library's page-locked staging buffers (valid
 * until ncg_destroy; actions float32[E*C*2] or int32[E*C], obs float32[E*C*38], reward float32[E*C], terminated and
 * truncated uint8[E], final_obs float32[E*C*38]).  The caller writes actions into `actions`, calls ncg_step_pinned and
 * reads the results in place; *any_done != 0 means at least one env finished (and, with want_final and auto_reset,
 * final_obs holds the terminal observations of the finished envs). */
int ncg_host_buffers(NcgHandle* h, void** actions, float** obs, float** reward, uint8_t** terminated, uint8_t** truncated,
                     float** final_obs);
int ncg_step_pinned(NcgHandle* h, int32_t want_final, int32_t* any_done);

/* Zero-staging variant: every buffer is page-locked, device-mapped host memory from ncg_host_alloc, and the kernel
 * itself reads the actions and writes observations / rewards / flags across PCIe -- a step is one kernel launch and one
 * stream synchronise.  Buffers are the caller's, so a binding can rotate result buffers and hand them out without
 * copying.  For envs that finish in this step (auto_reset on): h_final_obs rows (float32[E*C*38], may be NULL) get the
 * terminal observation, h_ep_return[car] (float32[E*C], may be NULL) the episode return CarEnv accumulates in
 * cumulative_rewards (src/car_env.py:785-789) and h_ep_length[env] (int32[E], may be NULL) the episode's step count;
 * rows of envs that did not finish are left untouched.  *any_done != 0 when at least one env finished. */
int ncg_host_alloc(size_t bytes, void** out);
int ncg_host_free(void* p);
int ncg_step_mapped(NcgHandle* h, const void* h_actions, float* h_obs, float* h_reward, uint8_t* h_terminated,
                    uint8_t* h_truncated, float* h_final_obs, float* h_ep_return, int32_t* h_ep_length, int32_t* any_done);

/* ncg_step_mapped in two halves, for a binding that has per-step work of its own (choosing the next result buffer, building
 * the objects it returns) to do while the GPU steps: ncg_step_mapped_post stages and checks src_actions like
 * ncg_step_mapped_from (src_actions may be NULL: the actions are already in b->actions) and hands the step to the GPU,
 * ncg_step_mapped_wait returns when its results are readable.  Exactly one wait per post; any other entry point called in
 * between completes the posted step first.  The buffers are those of ncg_step_mapped (final_obs, ep_return, ep_length may be NULL). */
typedef struct NcgMappedBuffers {
    void* actions; float* obs; float* reward; uint8_t* terminated; uint8_t* truncated;
    float* final_obs; float* ep_return; int32_t* ep_length;
} NcgMappedBuffers;
int ncg_step_mapped_post(NcgHandle* h, const void* src_actions, int32_t validate, const NcgMappedBuffers* b);
int ncg_step_mapped_wait(NcgHandle* h, int32_t* any_done);

/* ncg_step_mapped for a binding whose caller owns the action array: src_actions (pageable host memory, float32[E*C*2] or
 * int32[E*C]) is copied into the mapped buffer h_actions and, with validate != 0, checked in the same pass the way CarEnv.step
 * asserts action_space.contains(action) (src/car_env.py:694): continuous values in [-1, 1] and not NaN, discrete values in 0..4.
 * A failed check returns NCG_E_INVALID ("Invalid action") and steps nothing. */
int ncg_step_mapped_from(NcgHandle* h, const void* src_actions, int32_t validate, void* h_actions, float* h_obs, float* h_reward,
                         uint8_t* h_terminated, uint8_t* h_truncated, float* h_final_obs, float* h_ep_return, int32_t* h_ep_length,
                         int32_t* any_done);

/* Raw records, NCG_RECORD_WORDS words per car, car-major; d_records float32[n_cars*128].  ncg_set_state* read the
 * env -> track map out of the records (word NCG_R_TRACK of each env's first car), reject ids that were not uploaded
 * (nothing is written then) and re-plan the launch; the device variant synchronises `stream` once to do so. */
int ncg_get_state(NcgHandle* h, float* d_records, void* stream);
int ncg_set_state(NcgHandle* h, const float* d_records, void* stream);
int ncg_get_state_host(NcgHandle* h, float* h_records);
int ncg_set_state_host(NcgHandle* h, const float* h_records);
/* NcgConfig.car_contacts only: the rest of an env's state, the car-car contact table of its shared world (which pairs of cars
 * have a contact, touching or not, feature ids and warm-start impulses), NCG_CAR_PAIR_WORDS 32-bit words per env.  A checkpoint
 * of a shared-world engine is ncg_get_state_host + ncg_get_car_pairs_host; NCG_E_STATE without car_contacts. */
#define NCG_CAR_PAIR_WORDS 368
int ncg_get_car_pairs_host(NcgHandle* h, float* h_pairs);
int ncg_set_car_pairs_host(NcgHandle* h, const float* h_pairs);

/* Car.velocity_history (src/car.py:173, 384-386, 1058): the velocity (vx, vy) update_physics saw on each of the last
 * NCG_VEL_HISTORY steps of the running episode, a ring indexed by (episode step mod NCG_VEL_HISTORY);
 * h_out float32[n_cars][NCG_VEL_HISTORY][2].  Kept only with NcgConfig.track_info = 1 (NCG_E_STATE otherwise). */
int ncg_get_velocity_history_host(NcgHandle* h, float* h_out);

int ncg_read_stats(NcgHandle* h, NcgStats* out, int32_t reset);

/* The launch plan of ncg_step / ncg_rollout, exposed for inspection and tests (pure host code, no CUDA call): envs are
 * cut into CTAs of whole envs, at most 32 car slots each, never across a change of track id, and small batches are
 * spread over num_sms SMs.  Writes up to `capacity` entries and returns the number of CTAs. */
int32_t ncg_plan_ctas(const int32_t* h_env_track, int32_t num_envs, int32_t cars_per_env, int32_t num_sms,
                      int32_t* h_first_env, int32_t* h_num_envs, int32_t capacity);

/* CarEnv(track_file=None) for a batch (/root/reference/src/car_env.py:264-303, learn/ppo.py:65-77): with the redraw
 * enabled, an env that finishes inside ncg_step / ncg_step_mapped (auto_reset on) does not restart on its own track: it
 * moves to another of the uploaded tracks, drawn uniformly among the others (Philox keyed by `seed`, counter = (env, step
 * index)), gets brand-new physics worlds there (fresh reset) and returns that track's reset observation.  The host
 * re-groups the envs by track before the next launch; on the device-tensor path that costs one synchronisation of the
 * caller's stream per step (so a redrawing engine cannot be captured in a CUDA graph).  ncg_rollout never redraws.
 * ncg_get_env_tracks returns the current env -> track map (int32[E], host). */
int ncg_set_track_redraw(NcgHandle* h, int32_t enable, uint64_t seed);
int ncg_get_env_tracks(NcgHandle* h, int32_t* h_out);

/* Offsets of ncg_rollout's synthetic action stream: the Philox counter of local car c at launch step t is
 * (car_base + c, step_base + t).  R ranks that own disjoint env slices (rank r: car_base = r * E * C) draw the streams
 * of one R*E-env job; a rank's slice is then bit-identical to the same envs inside a single larger engine.
 * step_base otherwise counts the steps rolled out on the handle so far. */
int ncg_set_rollout_base(NcgHandle* h, uint32_t car_base, uint32_t step_base);

/* Monitor-style episode statistics on the device path (stable-baselines3 Monitor: /root/reference/learn/ppo.py:69):
 * once set (pointers may be NULL to unset), every ncg_step writes, for envs that finish in that step,
 * d_ep_return[car] = the episode return (CarEnv.cumulative_rewards, src/car_env.py:785-789) and d_ep_length[env] = the
 * episode's step count, and sets *d_any_done = 1; rows of envs that did not finish are left untouched (the caller
 * clears *d_any_done).  float32[E*C], int32[E], int32[1], device memory owned by the caller. */
int ncg_set_episode_outputs(NcgHandle* h, float* d_ep_return, int32_t* d_ep_length, int32_t* d_any_done);

/* Number of kernels this library has launched on the handle (for bench.py's gpu_launches). */
int64_t ncg_launch_count(NcgHandle* h);

/* ncg_step_mapped's resident mode: for a host-driven step loop the library keeps ONE launch of the step kernel on the SMs
 * (records and track table in shared memory) and feeds it a command per step through a mailbox in page-locked memory, instead of
 * launching per step; the kernel leaves by itself when no step arrives for NCG_RESIDENT_IDLE_US (default 1000; a caller that
 * lets that happen three times in a row gets per-step launches for a while, doubling) and every other entry point of this header
 * ends it first, so it is invisible except in time.  Batches without a resident kernel (more CTAs than
 * fit on the SMs at once, car_contacts, track redraw) and NCG_RESIDENT=0 take per-step launches.  While it is resident, CUDA calls
 * that synchronise the device (cudaMalloc, cudaHostAlloc, cudaDeviceSynchronize ...) wait for that idle time: a caller about to
 * make one can end the launch at once with ncg_resident_pause (the next ncg_step_mapped starts it again). */
int ncg_resident_pause(NcgHandle* h);

/* Diagnostics of ncg_step_mapped's resident mode (ends a running resident launch): out4 = {nanoseconds spent inside
 * ncg_step_mapped from entry to the kernel's done word, steps taken through the mailbox, nanoseconds on the device from "command
 * seen" to "done word raised", steps counted there}. */
int ncg_debug_resident(NcgHandle* h, unsigned long long* out4);

#ifdef __cplusplus
}
#endif
#endif /* NCG_B200_H */
