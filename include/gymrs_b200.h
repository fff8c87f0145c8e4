/*
 * gymrs_b200.h -- C ABI of the B200-native batched classic-control env stepper.
 *
 * This is the drop-in boundary under gym-rs's `Env` / `EnvProperties` traits
 * (reference: src/core.rs:25-57 and :60-90).  The reference has no FFI of its
 * own -- its step path is a scalar Rust method on one env object -- so each
 * entry point below names the Rust item it replaces.  One handle = one batch
 * of independent env instances living in HBM on one GPU, f32 struct-of-arrays.
 *
 * Conventions
 *   - every function returns a gymrs_status (0 = ok); nothing unwinds or aborts.
 *     gymrs_last_error() gives a thread-local message for the last failure.
 *   - a handle is NOT thread-safe (the reference's step/reset take &mut self,
 *     core.rs:42,45).  Different handles may be driven from different threads.
 *   - device work is enqueued on the handle's CUDA stream and is asynchronous
 *     to the host unless a function says it synchronises.
 *   - "device pointer" arguments must be readable from the handle's device;
 *     "host pointer" arguments are plain memory (pinned memory from
 *     gymrs_host_alloc makes the copies asynchronous and full speed).
 *   - SoA layout: a [rows][num_envs] array is `rows` contiguous runs of
 *     num_envs elements; handle-owned arrays use a row stride `ld` >= num_envs
 *     (see gymrs_buffers).
 *   - no CPU fallback exists: without a CUDA device gymrs_create fails with
 *     GYMRS_ERR_NO_DEVICE.
 */
#ifndef GYMRS_B200_H
#define GYMRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYMRS_ABI_VERSION 1

typedef struct gymrs_env gymrs_env; /* opaque handle */

typedef enum gymrs_kind {
    GYMRS_CARTPOLE = 0,     /* src/envs/classical_control/cartpole.rs     */
    GYMRS_MOUNTAIN_CAR = 1, /* src/envs/classical_control/mountain_car.rs */
    GYMRS_PENDULUM = 2      /* not in the reference; upstream Gym Pendulum-v1 */
} gymrs_kind;

typedef enum gymrs_status {
    GYMRS_OK = 0,
    GYMRS_ERR_INVALID_ACTION = 1, /* reference: assert! panic, cartpole.rs:402-406, mountain_car.rs:402-406 */
    GYMRS_ERR_BAD_ARG = 2,
    GYMRS_ERR_CUDA = 3,
    GYMRS_ERR_NO_DEVICE = 4,
    GYMRS_ERR_ALLOC = 5,
    GYMRS_ERR_UNSUPPORTED = 6
} gymrs_status;

/* gymrs_create flags */
#define GYMRS_FLAG_TIME_LIMIT 0x1u /* keep a per-env step counter and raise `truncated`
                                      at max_episode_steps.  OFF by default: the reference
                                      documents truncation (cartpole.rs:50, mountain_car.rs:45)
                                      but never implements it (truncated: false, cartpole.rs:480,
                                      mountain_car.rs:432). */

/* step flags */
#define GYMRS_STEP_AUTORESET 0x1u  /* envs whose step ended the episode are re-sampled in the
                                      same launch (what examples/cartpole.rs:23-28 does by hand);
                                      the observation returned for them is the fresh one. */

/* ---- physics parameters (the reference's `pub` fields; f64 like the reference,
 *      rounded to f32 once on the host when the handle is created/updated) ---- */

typedef struct gymrs_cartpole_params { /* cartpole.rs:63-80, defaults :94-103 */
    double gravity, masscart, masspole, length, force_mag, tau;
    double theta_threshold_radians, x_threshold;
    int32_t kinematics_integrator; /* 0 Euler (default), 1 semi-implicit (:380-387) */
    int32_t max_episode_steps;     /* used only with GYMRS_FLAG_TIME_LIMIT; default 500 */
} gymrs_cartpole_params;

typedef struct gymrs_mountain_car_params { /* mountain_car.rs:49-63, defaults :344-351 */
    double min_position, max_position, max_speed, goal_position, goal_velocity;
    double force, gravity;
    int32_t max_episode_steps; /* default 200 */
    int32_t _pad;
} gymrs_mountain_car_params;

typedef struct gymrs_pendulum_params { /* upstream Gym pendulum.py */
    double max_speed, max_torque, dt, g, m, l;
    int32_t max_episode_steps; /* default 200 */
    int32_t _pad;
} gymrs_pendulum_params;

/* Device pointers into the handle's SoA arrays (valid until gymrs_destroy). */
typedef struct gymrs_buffers {
    uint64_t num_envs;
    uint64_t ld;         /* row stride, in elements, of state/obs */
    uint32_t state_dim;  /* CartPole 4 (x, x_dot, theta, theta_dot); MountainCar 2 (position, velocity);
                            Pendulum 2 (theta, theta_dot) */
    uint32_t obs_dim;    /* CartPole 4, MountainCar 2 (observation == state), Pendulum 3 (cos, sin, theta_dot) */
    float *state;        /* [state_dim][ld] */
    float *obs;          /* [obs_dim][ld]; same pointer as state when observation == state */
    float *reward;       /* [num_envs]  ActionReward.reward   (core.rs:99) */
    uint8_t *done;       /* [num_envs]  ActionReward.done     (core.rs:101) */
    uint8_t *truncated;  /* [num_envs]  ActionReward.truncated(core.rs:103); all zero without TIME_LIMIT */
    int32_t *steps_beyond_terminated; /* CartPole only: -1 = None, k = Some(k) (cartpole.rs:81); else NULL */
    uint32_t *elapsed_steps;          /* TIME_LIMIT only, else NULL */
} gymrs_buffers;

/* ---- library ---- */
int gymrs_abi_version(void);
const char *gymrs_last_error(void);
/* number of CUDA devices visible (0 when there is no driver/GPU) */
int gymrs_device_count(void);

/* Fills *params (a gymrs_*_params matching `kind`) with the reference defaults
 * (cartpole.rs:94-103, mountain_car.rs:344-351). */
int gymrs_default_params(int kind, void *params);

/* ---- lifetime: CartPoleEnv::new / MountainCarEnv::new (cartpole.rs:91-144,
 *      mountain_car.rs:341-389), Clone (core.rs:25), Drop ---- */
/* global_env_offset: id of this handle's env 0 in the whole (possibly multi-GPU) batch;
 * reset sampling is keyed by the global id, so results do not depend on sharding.
 * params: NULL = defaults.  Like ::new, the initial state is an entropy-seeded reset. */
int gymrs_create(int kind, uint64_t num_envs, int device, uint64_t global_env_offset,
                 const void *params, uint32_t flags, gymrs_env **out);
int gymrs_destroy(gymrs_env *env);
int gymrs_clone(const gymrs_env *env, gymrs_env **out);

/* The reference's physics constants are `pub` fields a caller may mutate between steps. */
int gymrs_set_params(gymrs_env *env, const void *params);
int gymrs_get_params(const gymrs_env *env, void *params);

/* Run the handle's work on a caller-owned cudaStream_t (NULL restores the handle's own). */
int gymrs_set_stream(gymrs_env *env, void *cuda_stream);
int gymrs_get_stream(const gymrs_env *env, void **cuda_stream);

/* ---- Env::reset (core.rs:45-50; cartpole.rs:485-516, mountain_car.rs:464-501) ----
 * seed NULL = draw 64 bits of OS entropy (seeding.rs:22).  *seed_used (optional) receives
 * the seed, mirroring rand_random's returned tuple (seeding.rs:21-26).
 * low/high: host pointers to state_dim floats, or NULL for the reference defaults
 * (`options: Option<BoxR<Obs>>`).  mask: device pointer to num_envs bytes, NULL = all envs.
 * Env i takes its values from Philox4x32-10(key = seed, counter = (global id, epoch 0)). */
int gymrs_reset(gymrs_env *env, const uint64_t *seed, const float *low, const float *high,
                const uint8_t *mask, uint64_t *seed_used);

/* ---- Env::step (core.rs:42; cartpole.rs:398-483, mountain_car.rs:398-435) ----
 * actions: device pointer, int32[num_envs] for CartPole / MountainCar (Action = usize,
 * cartpole.rs:390, mountain_car.rs:393), float[num_envs] for Pendulum.
 * Results land in the handle's buffers (gymrs_buffers).  An action outside the action
 * space leaves that env untouched and raises a sticky GYMRS_ERR_INVALID_ACTION that the
 * next gymrs_sync reports (the reference panics). */
int gymrs_step(gymrs_env *env, const void *actions, uint32_t step_flags);

/* One gymrs_step for each of `count` (handle, action batch) pairs, in order, with a single FFI crossing:
 * for callers that keep a batch in several handles (one per GPU, or several env groups per GPU whose
 * launches then overlap on the handles' streams).  A handle may appear more than once (consecutive
 * steps of pre-generated actions).  Stops at the first error and returns it; *done (optional) receives
 * the number of steps that were enqueued. */
int gymrs_step_many(gymrs_env *const *envs, const void *const *actions, uint32_t count, uint32_t step_flags,
                    uint32_t *done);

/* gymrs_step_many bracketed by two caller-owned CUDA events (cudaEvent_t, either may be NULL): `begin_event` is
 * recorded on the first handle's stream before anything is launched and every other stream of the pass waits for
 * it; `end_event` is recorded on the first handle's stream after every other stream of the pass has been joined
 * into it -- ONE event that completes when the whole pass has (what a consumer on another stream waits for, or,
 * with timing enabled, what a pass took on the device with nothing but its launches between the two records).
 * The handles must share a device; not inside a stream capture.  The reference's counterpart is the caller's
 * loop over its env objects (examples/cartpole.rs:15-30), which needs no such thing on one CPU thread. */
int gymrs_step_pass(gymrs_env *const *envs, const void *const *actions, uint32_t count, uint32_t step_flags,
                    void *begin_event, void *end_event, uint32_t *done);

/* Same step with HOST buffers: copies actions in, steps, copies observation / reward / done
 * out (any output pointer may be NULL to skip it) and synchronises.  obs: [obs_dim][num_envs]. */
int gymrs_step_host(gymrs_env *env, const void *actions, uint32_t step_flags,
                    float *obs, float *reward, uint8_t *done, uint8_t *truncated);

/* Asynchronous form of gymrs_step_host for pipelined host loops (pinned host delivery overlapped
 * with the next step): enqueues copy-in, step and copy-out and returns at once; *ticket
 * identifies the step.  The host buffers belong to the library until gymrs_host_wait(env, ticket)
 * returns.  Up to two host steps may be in flight (use two sets of host buffers); every other
 * entry point on the handle first waits for them. */
int gymrs_step_host_async(gymrs_env *env, const void *actions, uint32_t step_flags,
                          float *obs, float *reward, uint8_t *done, uint8_t *truncated, uint64_t *ticket);
int gymrs_host_wait(gymrs_env *env, uint64_t ticket);

/* Host-buffer rollout: n_steps pipelined host steps driven from inside the library -- the loop
 * `for t { reset-on-done by hand; env.step(action) }` of examples/cartpole.rs:15-30 over a whole
 * batch, with the observation / reward / done of EVERY step delivered to host memory.  Equivalent
 * to calling gymrs_step_host_async(t) / gymrs_host_wait(t - 1) in a loop, without one FFI crossing
 * per step (one process per GPU then spends its host core on the consumer, not on the binding).
 *   step t reads   actions  slot (t % action_slots)
 *   step t writes  results  slot (t % result_slots), result_slots >= 2
 * on_step (optional) runs on the calling thread, in step order, as soon as step t's results are
 * complete in their slot; the slot is reused by step t + result_slots, so consume it before
 * returning.  Synchronises before it returns.
 * transport: compact wire formats for the PCIe crossing (lossless; the device unpacks / packs):
 *   GYMRS_HOST_U8_ACTIONS   discrete actions travel as uint8[num_envs] instead of int32
 *   GYMRS_HOST_PACKED_DONE  done / truncated travel as bits: byte i/8, bit i%8 (LSB first),
 *                           ceil(num_envs / 8) bytes per slot */
#define GYMRS_HOST_U8_ACTIONS 0x1u
#define GYMRS_HOST_PACKED_DONE 0x2u
typedef void (*gymrs_host_step_fn)(void *user, uint32_t step, uint32_t slot);
typedef struct gymrs_host_rollout_desc {
    const void *actions;   /* host [action_slots][num_envs] */
    float *obs;            /* host [result_slots][obs_dim][num_envs], or NULL */
    float *reward;         /* host [result_slots][num_envs], or NULL */
    uint8_t *done;         /* host [result_slots][num_envs] (or [..][ceil(num_envs/8)] packed), or NULL */
    uint8_t *truncated;    /* like done, or NULL */
    uint32_t action_slots;
    uint32_t result_slots;
    uint32_t transport;
    uint32_t _pad;
    gymrs_host_step_fn on_step;
    void *user;
} gymrs_host_rollout_desc;
int gymrs_rollout_host(gymrs_env *env, uint32_t n_steps, uint32_t step_flags,
                       const gymrs_host_rollout_desc *desc);

/* Fused rollout: n_steps consecutive steps in ONE launch, state held in registers.
 * actions: device [n_steps][num_envs].  Per-step results are streamed to the caller's
 * device arrays (any may be NULL): obs_out [n_steps][obs_dim][num_envs],
 * reward_out / done_out [n_steps][num_envs].  The handle's own buffers hold the last step. */
int gymrs_rollout(gymrs_env *env, const void *actions, uint32_t n_steps, uint32_t step_flags,
                  float *obs_out, float *reward_out, uint8_t *done_out);

/* ---- state access: `pub state` (cartpole.rs:60, mountain_car.rs:74), Serialize/Clone ----
 * host [state_dim][num_envs] floats; both synchronise.  sbt (CartPole, may be NULL):
 * int32[num_envs], -1 = None. */
int gymrs_get_state(gymrs_env *env, float *state, int32_t *sbt);
int gymrs_set_state(gymrs_env *env, const float *state, const int32_t *sbt);
int gymrs_get_buffers(gymrs_env *env, gymrs_buffers *out);

/* ---- checkpoint / resume: `Env: Clone + Serialize` (core.rs:25; derive(Serialize)
 *      cartpole.rs:51, mountain_car.rs:46-47) ----
 * A checkpoint is one self-describing host blob: a 256-byte header (kind, num_envs, flags,
 * global offset, the `pub` physics fields, reset bounds) followed by every per-env array of the
 * handle (state, observation, reward, done, truncated, steps_beyond_terminated, elapsed steps).
 * The reference skips its RNG when serialising (serde(skip_serializing), cartpole.rs:85-86,
 * mountain_car.rs:79-81) because a PCG64 state cannot be rebuilt from plain data; here reset
 * sampling is counter-based, so the blob also carries the Philox key and the step counter and a
 * restored handle continues BIT-IDENTICALLY to the one that was saved, auto-resets included.
 * save / load / create synchronise.  load needs a handle of the same kind, num_envs and
 * TIME_LIMIT flag; create builds a new handle on `device` from the blob alone. */
typedef struct gymrs_checkpoint_info {
    int32_t kind;
    uint32_t flags;
    uint64_t num_envs;
    uint64_t global_env_offset;
    uint64_t seed;       /* Philox key of the reset stream */
    uint64_t step_count; /* steps since the last full reset */
    uint64_t bytes;      /* total size of the blob */
} gymrs_checkpoint_info;
int gymrs_checkpoint_size(const gymrs_env *env, size_t *bytes);
int gymrs_checkpoint_save(gymrs_env *env, void *buf, size_t bytes);
int gymrs_checkpoint_load(gymrs_env *env, const void *buf, size_t bytes);
int gymrs_checkpoint_create(const void *buf, size_t bytes, int device, gymrs_env **out);
/* Validates magic, version, size and checksum of a blob without touching a device. */
int gymrs_checkpoint_info_of(const void *buf, size_t bytes, gymrs_checkpoint_info *info);

/* ---- EnvProperties (core.rs:60-90) ---- */
/* Discrete(n): *n = 2 / 3, low/high untouched.  Box (Pendulum): *n = 0, low/high = -+max_torque. */
int gymrs_action_space(const gymrs_env *env, uint64_t *n, float *low, float *high);
/* low/high: obs_dim doubles each (cartpole.rs:105-113, mountain_car.rs:353-364) */
int gymrs_observation_space(const gymrs_env *env, double *low, double *high);
/* (-inf, +inf), core.rs:16-19,81-83 */
int gymrs_reward_range(const gymrs_env *env, double *low, double *high);
int gymrs_num_envs(const gymrs_env *env, uint64_t *n);
int gymrs_kind_of(const gymrs_env *env, int *kind);

/* Wait for the handle's stream; returns the sticky error (invalid action / CUDA) and clears it.
 * *bad_env (optional) receives the global id of one offending env. */
int gymrs_sync(gymrs_env *env, uint64_t *bad_env);

/* Launch tuning (not part of the reference surface).
 * vec: env instances per thread, 0 = widest the buffer alignment allows (4), or 1 / 2 / 4;
 *      8 selects the persistent TMA-staged variant (cp.async.bulk into a shared-memory ring,
 *      4 envs per thread), which falls back to the plain kernel when alignment does not allow it.
 * block: threads per CTA, 0 = the library's choice (step: 128 for CartPole / MountainCar, 256 for
 *      Pendulum; rollout: 128 for CartPole, else 256), else a multiple of 32 up to 256.
 * pdl: 0 = plain stream order.
 *      1 (default) = programmatic dependent launch: the next step's CTAs are scheduled while the
 *        previous launch drains, but touch memory only after it has completed.
 *      2 = pipelined rollouts: back-to-back gymrs_step calls overlap.  A step reads its action
 *        batch immediately and waits, per CTA, only for the CTA of the handle's previous step that
 *        wrote the same env instances (per-handle progress flags), not for the whole previous
 *        launch.  Only valid when the action batch was complete before the PREVIOUS launch on the
 *        stream started (pre-generated rollouts): a kernel that writes the actions right before
 *        the step is not ordered before it.  Every other gymrs_* call on the handle, and every
 *        non-kernel operation on the stream, still acts as a full barrier. */
int gymrs_set_launch_config(gymrs_env *env, int vec, int block, int pdl);

/* Occupancy of the step kernel (launch tuning, not part of the reference surface).
 * wide = 0 (default): 48 registers per thread; a 1M-env launch is 1.38 waves of CTAs.  Fastest when
 *   launches overlap anyway: several handles on several streams, or a chained (pdl = 2) loop over
 *   ONE handle, where consecutive steps then pipeline CTA by CTA.
 * wide = 1: the same kernel compiled to a tighter register budget (MountainCar / Pendulum 32
 *   registers: every CTA of a 1M-env launch is resident at once; CartPole 40).  Faster where the
 *   launches of one stream are independent of their immediate predecessor (several handles stepped
 *   round-robin on one stream, pdl = 1 or 2: MountainCar -10 % / -16 % per step), slower for a
 *   chained loop over one handle (the CTAs of a step then run in lockstep behind their
 *   predecessors) and 2-5 % slower on two streams; same results bit for bit.  Applies to the
 *   128-bit host-counted step; other launches (scalar widths, CUDA-graph-captured steps, rollouts,
 *   the persistent kernel) ignore it. */
int gymrs_set_launch_occupancy(gymrs_env *env, int wide);

/* Pinned host memory for the *_host entry points. */
int gymrs_host_alloc(size_t bytes, void **out);
int gymrs_host_free(void *p);

/* Shared helpers restating small reference functions so bindings need no second copy:
 * clip = utils/custom/util_fns.rs:2-10, contains = spaces/discrete.rs:14-20,
 * rand_random = utils/seeding.rs:21-26 (returns the seed used). */
double gymrs_clip(double value, double left_bound, double right_bound);
int gymrs_discrete_contains(uint64_t n, uint64_t value);
uint64_t gymrs_rand_random(const uint64_t *seed);

#ifdef __cplusplus
}
#endif
#endif /* GYMRS_B200_H */
