/*
 * gymrs_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see gymrs_oracle.h).
 *
 * f64 scalar restatement of the reference step path, same operation order as
 * the Rust source.  Compile WITHOUT -ffast-math and with -ffp-contract=off so
 * every line below is one IEEE-754 double operation, as in the reference
 * (ordered-float's OrderedFloat<f64> arithmetic is plain f64 arithmetic; its
 * sin/cos go to the platform libm, like the calls here).
 *
 * Citations are into /root/reference/src/envs/classical_control/{cartpole,
 * mountain_car}.rs unless another file is named.
 */
#define _GNU_SOURCE
#include "gymrs_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* shared                                                                    */
/* ------------------------------------------------------------------------ */

/* Comparisons of O64 = ordered_float::OrderedFloat<f64> (src/utils/custom/types.rs:4).  The
 * reference never compares raw f64: OrderedFloat's Ord/PartialOrd is a TOTAL order in which NaN
 * equals NaN and is greater than every other value (ordered-float >= 3.9.1, Cargo.toml:33).  For
 * finite and infinite values these are the IEEE comparisons; they differ only on NaN. */
static int of_cmp(double a, double b)
{
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    /* at least one NaN */
    if (a != a) return (b != b) ? 0 : 1;
    return -1;
}
static int of_lt(double a, double b) { return of_cmp(a, b) < 0; }
static int of_le(double a, double b) { return of_cmp(a, b) <= 0; }
static int of_gt(double a, double b) { return of_cmp(a, b) > 0; }
static int of_ge(double a, double b) { return of_cmp(a, b) >= 0; }
static int of_eq(double a, double b) { return of_cmp(a, b) == 0; }

/* src/utils/custom/util_fns.rs:2-10 instantiated on O64: a NaN value is "greater than" the right
 * bound and is therefore clipped to it. */
double orc_clip(double value, double left_bound, double right_bound)
{
    if (of_le(left_bound, value) && of_le(value, right_bound)) {
        return value;
    } else if (of_gt(value, right_bound)) {
        return right_bound;
    } else {
        return left_bound;
    }
}

/* Same generic function instantiated on integers, which is what the
 * reference's own unit tests exercise (util_fns.rs:16-32). */
long long orc_clip_i64(long long value, long long left_bound, long long right_bound)
{
    if (left_bound <= value && value <= right_bound) {
        return value;
    } else if (value > right_bound) {
        return right_bound;
    } else {
        return left_bound;
    }
}

/* src/spaces/discrete.rs:14-20 */
int orc_discrete_contains(size_t n, size_t value)
{
    return value < n;
}

/* src/utils/seeding.rs:21-26 */
uint64_t orc_rand_random(int has_seed, uint64_t seed)
{
    if (has_seed) {
        return seed;
    }
    uint64_t s = 0;
    FILE *f = fopen("/dev/urandom", "rb");
    if (f) {
        if (fread(&s, sizeof s, 1, f) != 1) {
            s = 0;
        }
        fclose(f);
    }
    if (s == 0) {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        s = (uint64_t)ts.tv_nsec * 0x9E3779B97F4A7C15ull ^ (uint64_t)ts.tv_sec;
    }
    return s;
}

void orc_philox4x32_10(uint32_t c[4], const uint32_t key[2])
{
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

double orc_uniform_from_word(uint32_t w, double low, double high)
{
    double r = (double)(w >> 8) * (1.0 / 16777216.0); /* [0, 1) on a 2^-24 grid */
    return low + r * (high - low);
}

static void reset_words(uint64_t seed, uint64_t global_env_id, uint64_t epoch, uint32_t w[4])
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    w[0] = (uint32_t)global_env_id;
    w[1] = (uint32_t)(global_env_id >> 32);
    w[2] = (uint32_t)epoch;
    w[3] = (uint32_t)(epoch >> 32);
    orc_philox4x32_10(w, key);
}

/* ------------------------------------------------------------------------ */
/* CartPole                                                                  */
/* ------------------------------------------------------------------------ */

void orc_cartpole_default_params(orc_cartpole_params *p)
{
    p->gravity = 9.8;                                    /* :94 */
    p->masscart = 1.0;                                   /* :95 */
    p->masspole = 0.1;                                   /* :96 */
    p->length = 0.5;                                     /* :97 */
    p->force_mag = 10.0;                                 /* :98 */
    p->tau = 0.02;                                       /* :99 */
    p->kinematics_integrator = 0;                        /* :100 Euler */
    p->theta_threshold_radians = 12. * 2. * M_PI / 360.; /* :102 */
    p->x_threshold = 2.4;                                /* :103 */
}

void orc_cartpole_new(orc_cartpole_env *env)
{
    orc_cartpole_default_params(&env->p);
    memset(env->state, 0, sizeof env->state);
    env->steps_beyond_terminated = -1; /* None, :122 */
}

/* :146-148 */
static double cartpole_total_mass(const orc_cartpole_params *p)
{
    return p->masspole + p->masscart;
}

/* :150-152 -- an ADDITION in the reference (0.6), not the product OpenAI Gym
 * uses (0.05).  Parity is with the reference. */
static double cartpole_polemass_length(const orc_cartpole_params *p)
{
    return p->masspole + p->length;
}

int orc_cartpole_step(orc_cartpole_env *env, size_t action,
                      double *reward, int *done, int *truncated)
{
    const orc_cartpole_params *p = &env->p;

    /* :402-406 assert!(self.action_space.contains(action)) with Discrete(2) */
    if (!orc_discrete_contains(2, action)) {
        return 1;
    }

    /* :408-413 */
    double x = env->state[0];
    double x_dot = env->state[1];
    double theta = env->state[2];
    double theta_dot = env->state[3];

    /* :414-418 */
    double force = (action == 1) ? p->force_mag : -p->force_mag;

    /* :420-421 */
    double costheta = cos(theta);
    double sintheta = sin(theta);

    /* :423-424  (force + PML * theta_dot.powf(2.) * sintheta) / M
     * powf(2.) is llvm.pow.f64(x, 2.0), which LLVM lowers to x*x. */
    double temp = (force + cartpole_polemass_length(p) * (theta_dot * theta_dot) * sintheta)
                  / cartpole_total_mass(p);
    /* :425-428 */
    double thetaacc = (p->gravity * sintheta - costheta * temp)
                      / (p->length
                         * (4.0 / 3.0 - p->masspole * (costheta * costheta) / cartpole_total_mass(p)));
    /* :429 */
    double xacc = temp - cartpole_polemass_length(p) * thetaacc * costheta / cartpole_total_mass(p);

    if (p->kinematics_integrator == 0) {
        /* :431-436 Euler: positions advance with the OLD velocities */
        x += p->tau * x_dot;
        x_dot += p->tau * xacc;
        theta += p->tau * theta_dot;
        theta_dot += p->tau * thetaacc;
    } else {
        /* :437-441 semi-implicit */
        x_dot += p->tau * xacc;
        x += p->tau * x_dot;
        theta_dot += p->tau * thetaacc;
        theta += p->tau * theta_dot;
    }

    /* :443-448 */
    env->state[0] = x;
    env->state[1] = x_dot;
    env->state[2] = theta;
    env->state[3] = theta_dot;

    /* :450-453 strict comparisons on the UPDATED x, theta */
    int d = of_lt(x, -p->x_threshold)
         || of_gt(x, p->x_threshold)
         || of_lt(theta, -p->theta_threshold_radians)
         || of_gt(theta, p->theta_threshold_radians);

    /* :455-464 */
    double r;
    if (!d) {
        r = 1.0;
    } else if (env->steps_beyond_terminated < 0) {
        env->steps_beyond_terminated = 0;
        r = 1.0;
    } else {
        env->steps_beyond_terminated += 1;
        r = 0.;
    }

    /* :476-482 */
    *reward = r;
    *done = d;
    *truncated = 0;
    return 0;
}

void orc_cartpole_reset(orc_cartpole_env *env, uint64_t seed, uint64_t global_env_id,
                        uint64_t epoch, const double *low, const double *high)
{
    /* :353-361 default bounds, :317-324 draw order x, x_dot, theta, theta_dot */
    static const double dlow[4] = { -0.05, -0.05, -0.05, -0.05 };
    static const double dhigh[4] = { 0.05, 0.05, 0.05, 0.05 };
    if (!low) low = dlow;
    if (!high) high = dhigh;
    uint32_t w[4];
    reset_words(seed, global_env_id, epoch, w);
    for (int i = 0; i < 4; ++i) {
        env->state[i] = orc_uniform_from_word(w[i], low[i], high[i]);
    }
    env->steps_beyond_terminated = -1; /* :504 */
}

void orc_cartpole_observation_space(const orc_cartpole_params *p, double low[4], double high[4])
{
    /* :105-113 */
    high[0] = p->x_threshold * 2.;
    high[1] = INFINITY;
    high[2] = p->theta_threshold_radians * 2.;
    high[3] = INFINITY;
    for (int i = 0; i < 4; ++i) low[i] = -high[i];
}

/* ------------------------------------------------------------------------ */
/* MountainCar                                                               */
/* ------------------------------------------------------------------------ */

void orc_mountain_car_default_params(orc_mountain_car_params *p)
{
    p->min_position = -1.2; /* :344 */
    p->max_position = 0.6;  /* :345 */
    p->max_speed = 0.07;    /* :346 */
    p->goal_position = 0.5; /* :347 */
    p->goal_velocity = 0.;  /* :348 */
    p->force = 0.001;       /* :350 */
    p->gravity = 0.0025;    /* :351 */
}

void orc_mountain_car_new(orc_mountain_car_env *env)
{
    orc_mountain_car_default_params(&env->p);
    env->state[0] = -0.5;
    env->state[1] = 0.;
}

int orc_mountain_car_step(orc_mountain_car_env *env, size_t action,
                          double *reward, int *done, int *truncated)
{
    const orc_mountain_car_params *p = &env->p;

    /* :402-406 with Discrete(3) */
    if (!orc_discrete_contains(3, action)) {
        return 1;
    }

    /* :408-409 */
    double position = env->state[0];
    double velocity = env->state[1];

    /* :411-412  velocity += (a - 1) * force + cos(3 * position) * (-gravity)
     * (the right-hand side is evaluated first, then added) */
    double rhs = ((double)action - 1.) * p->force + cos(3. * position) * (-p->gravity);
    velocity = velocity + rhs;
    /* :413 */
    velocity = orc_clip(velocity, -p->max_speed, p->max_speed);

    /* :415-416 */
    position = position + velocity;
    position = orc_clip(position, p->min_position, p->max_position);

    /* :418-420 exact float equality after the clip */
    if (of_eq(position, p->min_position) && of_lt(velocity, 0.)) {
        velocity = 0.;
    }

    /* :422-423 */
    int d = of_ge(position, p->goal_position) && of_ge(velocity, p->goal_velocity);
    double r = -1.0;

    /* :425 */
    env->state[0] = position;
    env->state[1] = velocity;

    /* :428-434 */
    *reward = r;
    *done = d;
    *truncated = 0;
    return 0;
}

void orc_mountain_car_reset(orc_mountain_car_env *env, uint64_t seed, uint64_t global_env_id,
                            uint64_t epoch, const double *low, const double *high)
{
    /* :174-189 default bounds; :162-167 only the position is drawn */
    double lo = low ? low[0] : -0.6;
    double hi = high ? high[0] : -0.4;
    uint32_t w[4];
    reset_words(seed, global_env_id, epoch, w);
    env->state[0] = orc_uniform_from_word(w[0], lo, hi);
    env->state[1] = 0.;
}

void orc_mountain_car_observation_space(const orc_mountain_car_params *p, double low[2], double high[2])
{
    /* :353-354 */
    low[0] = p->min_position;
    low[1] = -p->max_speed;
    high[0] = p->max_position;
    high[1] = p->max_speed;
}

/* ------------------------------------------------------------------------ */
/* Pendulum-v1 (upstream OpenAI Gym pendulum.py)       */
/* ------------------------------------------------------------------------ */

void orc_pendulum_default_params(orc_pendulum_params *p)
{
    p->max_speed = 8.0;
    p->max_torque = 2.0;
    p->dt = 0.05;
    p->g = 10.0;
    p->m = 1.0;
    p->l = 1.0;
}

void orc_pendulum_new(orc_pendulum_env *env)
{
    orc_pendulum_default_params(&env->p);
    env->state[0] = 0.;
    env->state[1] = 0.;
}

/* numpy.clip, which upstream Gym's pendulum.py uses (NaN propagates) */
static double np_clip(double v, double lo, double hi)
{
    if (v != v) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

/* ((x + pi) mod 2 pi) - pi with a floor-mod (result in [-pi, pi)) */
double orc_angle_normalize(double x)
{
    double two_pi = 2. * M_PI;
    double y = x + M_PI;
    double m = y - two_pi * floor(y / two_pi);
    return m - M_PI;
}

int orc_pendulum_step(orc_pendulum_env *env, double action, double obs[3],
                      double *reward, int *done, int *truncated)
{
    const orc_pendulum_params *p = &env->p;
    double th = env->state[0];
    double thdot = env->state[1];

    double u = np_clip(action, -p->max_torque, p->max_torque);
    double an = orc_angle_normalize(th);
    double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u);

    double newthdot = thdot + (3. * p->g / (2. * p->l) * sin(th) + 3.0 / (p->m * (p->l * p->l)) * u) * p->dt;
    newthdot = np_clip(newthdot, -p->max_speed, p->max_speed);
    double newth = th + newthdot * p->dt;

    env->state[0] = newth;
    env->state[1] = newthdot;

    obs[0] = cos(newth);
    obs[1] = sin(newth);
    obs[2] = newthdot;
    *reward = -costs;
    *done = 0;
    *truncated = 0;
    return 0;
}

void orc_pendulum_reset(orc_pendulum_env *env, uint64_t seed, uint64_t global_env_id,
                        uint64_t epoch, const double *low, const double *high)
{
    static const double dlow[2] = { -M_PI, -1.0 };
    static const double dhigh[2] = { M_PI, 1.0 };
    if (!low) low = dlow;
    if (!high) high = dhigh;
    uint32_t w[4];
    reset_words(seed, global_env_id, epoch, w);
    env->state[0] = orc_uniform_from_word(w[0], low[0], high[0]);
    env->state[1] = orc_uniform_from_word(w[1], low[1], high[1]);
}

void orc_pendulum_observation_space(const orc_pendulum_params *p, double low[3], double high[3])
{
    high[0] = 1.0;
    high[1] = 1.0;
    high[2] = p->max_speed;
    for (int i = 0; i < 3; ++i) low[i] = -high[i];
}

/* ------------------------------------------------------------------------ */
/* batched drivers                                                           */
/* ------------------------------------------------------------------------ */

long long orc_step_batch(int kind, const void *params, size_t n, double *state,
                         long long *sbt, const void *actions, double *obs,
                         double *reward, uint8_t *done)
{
    long long invalid = 0;
    int d = 0, t = 0;
    double r = 0.;
    if (kind == 0) {
        orc_cartpole_env e;
        orc_cartpole_new(&e);
        if (params) e.p = *(const orc_cartpole_params *)params;
        const int32_t *a = (const int32_t *)actions;
        for (size_t i = 0; i < n; ++i) {
            for (int k = 0; k < 4; ++k) e.state[k] = state[k * n + i];
            e.steps_beyond_terminated = sbt ? sbt[i] : -1;
            /* a negative i32 is not representable as usize: invalid */
            if (a[i] < 0 || orc_cartpole_step(&e, (size_t)a[i], &r, &d, &t)) {
                ++invalid;
                continue;
            }
            for (int k = 0; k < 4; ++k) state[k * n + i] = e.state[k];
            if (obs) for (int k = 0; k < 4; ++k) obs[k * n + i] = e.state[k];
            if (sbt) sbt[i] = e.steps_beyond_terminated;
            reward[i] = r;
            done[i] = (uint8_t)d;
        }
    } else if (kind == 1) {
        orc_mountain_car_env e;
        orc_mountain_car_new(&e);
        if (params) e.p = *(const orc_mountain_car_params *)params;
        const int32_t *a = (const int32_t *)actions;
        for (size_t i = 0; i < n; ++i) {
            e.state[0] = state[i];
            e.state[1] = state[n + i];
            if (a[i] < 0 || orc_mountain_car_step(&e, (size_t)a[i], &r, &d, &t)) {
                ++invalid;
                continue;
            }
            state[i] = e.state[0];
            state[n + i] = e.state[1];
            if (obs) { obs[i] = e.state[0]; obs[n + i] = e.state[1]; }
            reward[i] = r;
            done[i] = (uint8_t)d;
        }
    } else if (kind == 2) {
        orc_pendulum_env e;
        orc_pendulum_new(&e);
        if (params) e.p = *(const orc_pendulum_params *)params;
        const double *a = (const double *)actions;
        double o[3];
        for (size_t i = 0; i < n; ++i) {
            e.state[0] = state[i];
            e.state[1] = state[n + i];
            orc_pendulum_step(&e, a[i], o, &r, &d, &t);
            state[i] = e.state[0];
            state[n + i] = e.state[1];
            if (obs) { obs[i] = o[0]; obs[n + i] = o[1]; obs[2 * n + i] = o[2]; }
            reward[i] = r;
            done[i] = (uint8_t)d;
        }
    } else {
        return -1;
    }
    return invalid;
}

void orc_reset_batch(int kind, size_t n, double *state, uint64_t seed,
                     uint64_t global_env_offset, uint64_t epoch,
                     const double *low, const double *high, const uint8_t *mask)
{
    for (size_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        uint64_t gid = global_env_offset + i;
        if (kind == 0) {
            orc_cartpole_env e;
            orc_cartpole_new(&e);
            orc_cartpole_reset(&e, seed, gid, epoch, low, high);
            for (int k = 0; k < 4; ++k) state[k * n + i] = e.state[k];
        } else if (kind == 1) {
            orc_mountain_car_env e;
            orc_mountain_car_new(&e);
            orc_mountain_car_reset(&e, seed, gid, epoch, low, high);
            state[i] = e.state[0];
            state[n + i] = e.state[1];
        } else if (kind == 2) {
            orc_pendulum_env e;
            orc_pendulum_new(&e);
            orc_pendulum_reset(&e, seed, gid, epoch, low, high);
            state[i] = e.state[0];
            state[n + i] = e.state[1];
        }
    }
}

/* ------------------------------------------------------------------------ */
/* CPU baseline: the reference's scalar loop over an array of env objects    */
/* ------------------------------------------------------------------------ */

typedef struct {
    int kind;
    size_t begin, end;
    int n_steps, n_warmup;
    uint64_t seed;
    void *envs;          /* array of orc_*_env */
    double *obs;         /* [n][obs_dim] per-env records, like Vec<ActionReward> */
    double *reward;
    uint8_t *done;
    pthread_barrier_t *bar;
    double checksum;
    double t0, t1;
    int n_burnin, n_regions;
    double *region_s;
} bench_arg;

static inline uint64_t xorshift64(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return *s = x;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* one batch step of this thread's shard: every env object steps once with a random action and is
 * reset when it ends (examples/cartpole.rs:18-28) */
static void bench_step_shard(bench_arg *a, uint64_t *rng, uint64_t epoch)
{
    double r = 0.;
    int d = 0, t = 0;
    if (a->kind == 0) {
        orc_cartpole_env *envs = (orc_cartpole_env *)a->envs;
        for (size_t i = a->begin; i < a->end; ++i) {
            size_t act = (size_t)(xorshift64(rng) >> 63);
            orc_cartpole_step(&envs[i], act, &r, &d, &t);
            memcpy(&a->obs[4 * i], envs[i].state, 4 * sizeof(double));
            a->reward[i] = r;
            a->done[i] = (uint8_t)d;
            if (d) orc_cartpole_reset(&envs[i], a->seed, i, epoch, NULL, NULL);
        }
    } else if (a->kind == 1) {
        orc_mountain_car_env *envs = (orc_mountain_car_env *)a->envs;
        for (size_t i = a->begin; i < a->end; ++i) {
            size_t act = (size_t)((xorshift64(rng) >> 33) % 3u);
            orc_mountain_car_step(&envs[i], act, &r, &d, &t);
            memcpy(&a->obs[2 * i], envs[i].state, 2 * sizeof(double));
            a->reward[i] = r;
            a->done[i] = (uint8_t)d;
            if (d) orc_mountain_car_reset(&envs[i], a->seed, i, epoch, NULL, NULL);
        }
    } else {
        orc_pendulum_env *envs = (orc_pendulum_env *)a->envs;
        for (size_t i = a->begin; i < a->end; ++i) {
            double act = ((double)(xorshift64(rng) >> 11) * (1.0 / 9007199254740992.0)) * 4. - 2.;
            orc_pendulum_step(&envs[i], act, &a->obs[3 * i], &r, &d, &t);
            a->reward[i] = r;
            a->done[i] = (uint8_t)d;
        }
    }
}

/* Regions: n_burnin untimed steps once (episode phases decorrelate, like the burn-in of the GPU arm's
 * ring), then n_regions times { n_warmup untimed steps, n_steps timed steps }.  region_s[r] is the
 * wall time of region r from the barrier in front of its first timed step to the barrier behind
 * its last one (taken by the first shard's thread; every thread passes both barriers). */
static void *bench_worker(void *vp)
{
    bench_arg *a = (bench_arg *)vp;
    uint64_t rng = a->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull * (a->begin + 1);
    if (rng == 0) rng = 1;
    uint64_t epoch = 0;
    for (int s = 0; s < a->n_burnin; ++s) {
        bench_step_shard(a, &rng, ++epoch);
        pthread_barrier_wait(a->bar);
    }
    for (int r = 0; r < a->n_regions; ++r) {
        for (int s = 0; s < a->n_warmup; ++s) {
            bench_step_shard(a, &rng, ++epoch);
            pthread_barrier_wait(a->bar);
        }
        pthread_barrier_wait(a->bar);
        const double t0 = now_s();
        for (int s = 0; s < a->n_steps; ++s) {
            bench_step_shard(a, &rng, ++epoch);
            /* a batch step is complete only when every shard is: same barrier a
             * host loop over many reference env objects would need */
            pthread_barrier_wait(a->bar);
        }
        const double t1 = now_s();
        if (r == 0) a->t0 = t0;
        a->t1 = t1;
        if (a->region_s && a->begin == 0) a->region_s[r] = t1 - t0;
    }
    double cs = 0.;
    for (size_t i = a->begin; i < a->end; ++i) cs += a->reward[i] + a->done[i];
    a->checksum = cs;
    return NULL;
}

double orc_bench_regions(int kind, size_t n_envs, int n_steps, int n_warmup, int n_burnin, int n_regions,
                         int n_threads, uint64_t seed, double *region_s, double *checksum)
{
    if (n_regions < 1) n_regions = 1;
    if (n_burnin < 0) n_burnin = 0;
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_envs) n_threads = (int)n_envs;
    int obs_dim = kind == 0 ? 4 : (kind == 1 ? 2 : 3);
    size_t esz = kind == 0 ? sizeof(orc_cartpole_env)
               : kind == 1 ? sizeof(orc_mountain_car_env) : sizeof(orc_pendulum_env);
    void *envs = malloc(esz * n_envs);
    double *obs = (double *)malloc(sizeof(double) * obs_dim * n_envs);
    double *reward = (double *)malloc(sizeof(double) * n_envs);
    uint8_t *done = (uint8_t *)malloc(n_envs);
    if (!envs || !obs || !reward || !done) {
        free(envs); free(obs); free(reward); free(done);
        return -1.;
    }
    for (size_t i = 0; i < n_envs; ++i) {
        if (kind == 0) {
            orc_cartpole_env *e = (orc_cartpole_env *)envs + i;
            orc_cartpole_new(e);
            orc_cartpole_reset(e, seed, i, 0, NULL, NULL);
        } else if (kind == 1) {
            orc_mountain_car_env *e = (orc_mountain_car_env *)envs + i;
            orc_mountain_car_new(e);
            orc_mountain_car_reset(e, seed, i, 0, NULL, NULL);
        } else {
            orc_pendulum_env *e = (orc_pendulum_env *)envs + i;
            orc_pendulum_new(e);
            orc_pendulum_reset(e, seed, i, 0, NULL, NULL);
        }
        reward[i] = 0.;
        done[i] = 0;
    }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    bench_arg *args = (bench_arg *)calloc(n_threads, sizeof(bench_arg));
    size_t per = (n_envs + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        bench_arg *a = &args[t];
        a->kind = kind;
        a->begin = per * t < n_envs ? per * t : n_envs;
        a->end = per * (t + 1) < n_envs ? per * (t + 1) : n_envs;
        a->n_steps = n_steps;
        a->n_warmup = n_warmup;
        a->n_burnin = n_burnin;
        a->n_regions = n_regions;
        a->region_s = region_s;
        a->seed = seed;
        a->envs = envs;
        a->obs = obs;
        a->reward = reward;
        a->done = done;
        a->bar = &bar;
        pthread_create(&th[t], NULL, bench_worker, a);
    }
    double t0 = 1e300, t1 = 0., cs = 0.;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        if (args[t].t0 < t0) t0 = args[t].t0;
        if (args[t].t1 > t1) t1 = args[t].t1;
        cs += args[t].checksum;
    }
    for (size_t i = 0; i < (size_t)obs_dim * n_envs; i += 4097) cs += obs[i];
    if (checksum) *checksum = cs;
    pthread_barrier_destroy(&bar);
    free(th); free(args); free(envs); free(obs); free(reward); free(done);
    return t1 - t0;
}

/* one region, no burn-in: n_warmup untimed + n_steps timed steps from a fresh reset */
double orc_bench_rollout(int kind, size_t n_envs, int n_steps, int n_warmup,
                         int n_threads, uint64_t seed, double *checksum)
{
    return orc_bench_regions(kind, n_envs, n_steps, n_warmup, 0, 1, n_threads, seed, NULL, checksum);
}
