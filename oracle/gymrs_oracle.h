/*
 * gymrs_oracle.h -- CPU oracle for the gym-rs classic-control step path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, f64, scalar restatement of the
 * reference's algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libgymrs_b200.so) never links, loads or calls anything in this directory.
 *
 * PARITY STATUS
 *   - clip, Discrete::contains, seed echo: pinned by the reference's own unit
 *     tests (src/utils/custom/util_fns.rs:16-32, src/spaces/discrete.rs:27-41,
 *     src/utils/seeding.rs:33-39), restated in tests/test_oracle.py.
 *   - CartPole / MountainCar step arithmetic: PARITY UNPINNED by the reference
 *     (it has no test that calls step/reset, and no Rust toolchain exists here
 *     to run it).  The oracle is pinned instead against SURVEY.md Appendix B
 *     known-answer vectors and an independent 50-digit mpmath evaluation of the
 *     cited formulas (tests/golden/make_golden.py).
 *   - Pendulum: NOT IN THE REFERENCE.  Follows upstream OpenAI Gym Pendulum-v1
 *     (SURVEY.md Appendix D).  Parity unpinned.
 *   - reset(): the reference draws from rand_pcg::Pcg64 (un-pinned crate
 *     versions, Cargo.lock git-ignored).  Bit parity with that stream is not
 *     attempted; both the oracle and the device use Philox4x32-10 keyed by
 *     (seed, global env id, epoch), so oracle-vs-device reset IS bit-checkable.
 */
#ifndef GYMRS_ORACLE_H
#define GYMRS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shared small pieces ------------------------------------------------ */

/* src/utils/custom/util_fns.rs:2-10 (same branch order). */
double orc_clip(double value, double left_bound, double right_bound);
long long orc_clip_i64(long long value, long long left_bound, long long right_bound);

/* src/spaces/discrete.rs:14-20: value < n on usize. */
int orc_discrete_contains(size_t n, size_t value);

/* src/utils/seeding.rs:21-26: returns the seed that was used (the given one,
 * or 64 bits of OS entropy when has_seed == 0). */
uint64_t orc_rand_random(int has_seed, uint64_t seed);

/* Philox4x32-10 (Salmon et al., SC'11; same function as curand's
 * curand_Philox4x32_10).  ctr is replaced by the 4 output words. */
void orc_philox4x32_10(uint32_t ctr[4], const uint32_t key[2]);

/* Uniform in [low, high) from one Philox word:  low + (w>>8)*2^-24*(high-low). */
double orc_uniform_from_word(uint32_t w, double low, double high);

/* ---- CartPole: src/envs/classical_control/cartpole.rs ------------------- */

typedef struct {
    double gravity;                 /* :94  9.8  */
    double masscart;                /* :95  1.0  */
    double masspole;                /* :96  0.1  */
    double length;                  /* :97  0.5  */
    double force_mag;               /* :98  10.0 */
    double tau;                     /* :99  0.02 */
    double theta_threshold_radians; /* :102 12*2*pi/360 */
    double x_threshold;             /* :103 2.4  */
    int kinematics_integrator;      /* :100 0 = Euler, 1 = Other (semi-implicit) */
} orc_cartpole_params;

/* One reference CartPoleEnv object (the fields the step path touches,
 * cartpole.rs:52-87). */
typedef struct {
    orc_cartpole_params p;
    double state[4];                  /* x, x_dot, theta, theta_dot (:329-334) */
    long long steps_beyond_terminated; /* -1 = None (:81) */
} orc_cartpole_env;

void orc_cartpole_default_params(orc_cartpole_params *p);
void orc_cartpole_new(orc_cartpole_env *env);
/* cartpole.rs:398-483.  Returns 0, or 1 if the action is not in Discrete(2)
 * (the reference panics there, :402-406); state is then left untouched. */
int orc_cartpole_step(orc_cartpole_env *env, size_t action,
                      double *reward, int *done, int *truncated);
/* cartpole.rs:485-516 with the Philox stream described above.
 * low/high may be NULL (defaults -0.05/+0.05 for all four, :353-361). */
void orc_cartpole_reset(orc_cartpole_env *env, uint64_t seed, uint64_t global_env_id,
                        uint64_t epoch, const double *low, const double *high);
/* observation_space: +-(4.8, inf, 2*theta_thr, inf) (:105-113) */
void orc_cartpole_observation_space(const orc_cartpole_params *p, double low[4], double high[4]);

/* ---- MountainCar: src/envs/classical_control/mountain_car.rs ------------ */

typedef struct {
    double min_position;  /* :344 -1.2  */
    double max_position;  /* :345  0.6  */
    double max_speed;     /* :346  0.07 */
    double goal_position; /* :347  0.5  */
    double goal_velocity; /* :348  0.0  */
    double force;         /* :350  0.001  */
    double gravity;       /* :351  0.0025 */
} orc_mountain_car_params;

typedef struct {
    orc_mountain_car_params p;
    double state[2]; /* position, velocity (:123-128) */
} orc_mountain_car_env;

void orc_mountain_car_default_params(orc_mountain_car_params *p);
void orc_mountain_car_new(orc_mountain_car_env *env);
/* mountain_car.rs:398-435.  Returns 1 on an action outside Discrete(3). */
int orc_mountain_car_step(orc_mountain_car_env *env, size_t action,
                          double *reward, int *done, int *truncated);
/* mountain_car.rs:464-501: position ~ U[low0, high0) (default -0.6/-0.4),
 * velocity = 0 regardless of bounds (:162-167). */
void orc_mountain_car_reset(orc_mountain_car_env *env, uint64_t seed, uint64_t global_env_id,
                            uint64_t epoch, const double *low, const double *high);
void orc_mountain_car_observation_space(const orc_mountain_car_params *p, double low[2], double high[2]);

/* ---- Pendulum-v1 (upstream Gym; not in the reference) ------------------- */

typedef struct {
    double max_speed;  /* 8.0  */
    double max_torque; /* 2.0  */
    double dt;         /* 0.05 */
    double g;          /* 10.0 */
    double m;          /* 1.0  */
    double l;          /* 1.0  */
} orc_pendulum_params;

typedef struct {
    orc_pendulum_params p;
    double state[2]; /* theta, theta_dot */
} orc_pendulum_env;

void orc_pendulum_default_params(orc_pendulum_params *p);
void orc_pendulum_new(orc_pendulum_env *env);
double orc_angle_normalize(double x);
/* obs = (cos th', sin th', thdot'); reward = -cost; done = 0 always. */
int orc_pendulum_step(orc_pendulum_env *env, double action, double obs[3],
                      double *reward, int *done, int *truncated);
/* theta ~ U[-pi, pi), theta_dot ~ U[-1, 1) by default. */
void orc_pendulum_reset(orc_pendulum_env *env, uint64_t seed, uint64_t global_env_id,
                        uint64_t epoch, const double *low, const double *high);
void orc_pendulum_observation_space(const orc_pendulum_params *p, double low[3], double high[3]);

/* ---- batched drivers over arrays of scalar env objects ------------------- */
/* kind: 0 cartpole, 1 mountain car, 2 pendulum.  State arrays are SoA f64
 * [state_dim][n] so tests can feed the device's f32 state straight in.
 * actions: int32 for kinds 0/1, f64 for kind 2.  sbt: int64[n] or NULL
 * (treated as None everywhere; cartpole only).  obs: [obs_dim][n] or NULL
 * (for kinds 0/1 the observation is the state).  Returns the number of
 * invalid actions encountered (those envs are not stepped). */
long long orc_step_batch(int kind, const void *params, size_t n, double *state,
                         long long *sbt, const void *actions, double *obs,
                         double *reward, uint8_t *done);
void orc_reset_batch(int kind, size_t n, double *state, uint64_t seed,
                     uint64_t global_env_offset, uint64_t epoch,
                     const double *low, const double *high, const uint8_t *mask);

/* ---- CPU baseline: the reference's scalar loop, timed -------------------- */
/* One "step" = every one of n_envs env OBJECTS (array of structs, as a user
 * of the reference would hold them) stepped once with a random action from a
 * cheap xorshift PRNG, reset on done as in examples/cartpole.rs:23-28, results
 * written to per-env output arrays.  Envs are statically sharded over
 * n_threads pthreads.  Returns wall seconds for n_steps steps (after
 * n_warmup untimed ones); *checksum receives a value that depends on every
 * output so the loop cannot be optimised away. */
double orc_bench_rollout(int kind, size_t n_envs, int n_steps, int n_warmup,
                         int n_threads, uint64_t seed, double *checksum);

/* The same loop over ONE set of env objects and one set of threads, timed as regions:
 * n_burnin untimed steps first (the synchronised initial reset makes most envs end around the
 * same step; the burn-in lets episode phases decorrelate, like the GPU arm's), then n_regions
 * times { n_warmup untimed steps, n_steps timed steps }.  region_s[n_regions] receives the wall
 * seconds of each region's timed steps.  Returns first-timed-step to last-timed-step seconds. */
double orc_bench_regions(int kind, size_t n_envs, int n_steps, int n_warmup, int n_burnin, int n_regions,
                         int n_threads, uint64_t seed, double *region_s, double *checksum);

#ifdef __cplusplus
}
#endif
#endif /* GYMRS_ORACLE_H */
