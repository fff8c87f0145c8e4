// envs.cuh -- per-env f32 dynamics for the three classic-control envs (device side).
//
// Each struct is a policy class consumed by the kernels in kernels.cu:
//   SD / OD          rows of state / observation
//   OBS_IS_STATE     the observation IS the post-step state (CartPole, MountainCar)
//   HAS_SBT          keeps the reference's steps_beyond_terminated (CartPole)
//   Action, P        action element type, by-value parameter block
//   valid()          Space::contains on the action      (spaces/discrete.rs:14-20)
//   step()           one env transition on registers
//   reset()          fresh state from 4 Philox words
//
// The arithmetic is NOT a transliteration of the Rust lines: constants that the
// reference recomputes per step in f64 (total_mass(), polemass_length(), the
// divisions by total_mass) are folded on the host in f64 (capi.cu, fold_*) and the
// device evaluates a short FMA chain with a single division.  Results are compared
// with the f64 oracle to 1e-6 * max(1, |ref|) (tests/test_parity_gpu.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"

namespace gymrs {

struct ResetBox {
    float low[4];   // lower bound per state row
    float scale24[4]; // (high - low) * 2^-24: maps the 24 random bits straight onto [low, high)
    float cap[4];   // largest float below high
};

// num / den as MUFU.RCP plus one Newton correction on the quotient (4 instructions, result
// within 1 ulp for normal operands) instead of the ~10-instruction IEEE division sequence with
// its slow-path call.  A zero / non-finite den yields inf or NaN (never a finite wrong value).
__device__ __forceinline__ float div_newton(float num, float den)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    const float q = num * r;
    return fmaf(fmaf(-den, q, num), r, q);
}

// sin and cos of a pole angle.  A live CartPole has |theta| <= 0.21 (+ one step), so the common
// case needs neither range reduction nor quadrant selection: two short minimax polynomials on
// [-pi/4, pi/4] (Cephes sinf/cosf coefficients, <= 1 ulp there).  Anything larger (a pole that
// keeps falling when the caller never resets) takes the full-range sincosf.
__device__ __forceinline__ void sincos_small(float x, float &s, float &c)
{
    if (fabsf(x) <= 0.78539816f) {
        const float z = x * x;
        const float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
        const float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
        s = fmaf(x * z, ps, x);
        c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    } else {
        sincosf(x, &s, &c);
    }
}

// ---------------------------------------------------------------------------
// CartPole -- reference: src/envs/classical_control/cartpole.rs:398-483
// ---------------------------------------------------------------------------
struct CartPoleP {
    float tau;          // :99
    float force_over_m; // force_mag / (masspole + masscart)             (:414-418, :423-424, :146-148)
    float pml_over_m;   // (masspole + length) / total_mass  -- PML is a SUM in the reference (:150-152)
    float gravity;      // :94
    float den_a;        // length * 4/3                                  (:426-428)
    float den_b;        // length * masspole / total_mass
    float x_thr;        // largest float <= x_threshold: (v > x_thr) == (v > 2.4 in f64) for every float v
    float th_thr;       // same for theta_threshold_radians
    int semi_implicit;  // KinematicsIntegrator::Other (:437-441)
    ResetBox rb;        // :353-361 (defaults +-0.05) or caller's BoxR
};

struct CartPole {
    static constexpr int SD = 4, OD = 4;
    static constexpr bool OBS_IS_STATE = true;
    static constexpr bool HAS_SBT = true;
    using Action = int32_t;
    using P = CartPoleP;

    __device__ __forceinline__ static bool valid(Action a) { return (uint32_t)a < 2u; }

    // s = (x, x_dot, theta, theta_dot).  reward is decided by the caller (it depends on
    // steps_beyond_terminated, :455-464).
    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&)[OD], float &reward, bool &done)
    {
        const float x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        const float f = (a == 1) ? p.force_over_m : -p.force_over_m; // :414-418, already / M
        float sn, cs;
        sincos_small(theta, sn, cs); // :420-421 (polynomial / sincosf, never the MUFU __sinf)
        // temp = (force + PML * theta_dot^2 * sin) / M                               :423-424
        const float temp = fmaf(p.pml_over_m * (theta_dot * theta_dot), sn, f);
        // thetaacc = (g sin - cos temp) / (l (4/3 - mp cos^2 / M))                   :425-428
        const float num = fmaf(p.gravity, sn, -(cs * temp));
        const float den = fmaf(-p.den_b, cs * cs, p.den_a);
        const float thetaacc = div_newton(num, den);
        // xacc = temp - PML thetaacc cos / M                                         :429
        const float xacc = fmaf(-(p.pml_over_m * thetaacc), cs, temp);
        float nx, nxd, nth, nthd;
        if (!p.semi_implicit) { // Euler: positions use the OLD velocities          :431-436
            nx = fmaf(p.tau, x_dot, x);
            nxd = fmaf(p.tau, xacc, x_dot);
            nth = fmaf(p.tau, theta_dot, theta);
            nthd = fmaf(p.tau, thetaacc, theta_dot);
        } else { //                                                                  :437-441
            nxd = fmaf(p.tau, xacc, x_dot);
            nx = fmaf(p.tau, nxd, x);
            nthd = fmaf(p.tau, thetaacc, theta_dot);
            nth = fmaf(p.tau, nthd, theta);
        }
        s[0] = nx; s[1] = nxd; s[2] = nth; s[3] = nthd;
        // strict comparisons on the updated x, theta                                 :450-453
        // (x < -T || x > T) == (|x| > T).  The reference compares OrderedFloat values, a total
        // order in which NaN is greater than everything, so a NaN state IS done: !(|x| <= T).
        done = !(fabsf(nx) <= p.x_thr) | !(fabsf(nth) <= p.th_thr);
        reward = 1.0f;
    }

    // four iid uniforms in the order x, x_dot, theta, theta_dot                      :317-324
    __device__ __forceinline__ static void reset(const P &p, float (&s)[SD], float (&)[OD], uint4 w)
    {
        s[0] = uniform_from_word(w.x, p.rb.low[0], p.rb.scale24[0], p.rb.cap[0]);
        s[1] = uniform_from_word(w.y, p.rb.low[1], p.rb.scale24[1], p.rb.cap[1]);
        s[2] = uniform_from_word(w.z, p.rb.low[2], p.rb.scale24[2], p.rb.cap[2]);
        s[3] = uniform_from_word(w.w, p.rb.low[3], p.rb.scale24[3], p.rb.cap[3]);
    }
};

// ---------------------------------------------------------------------------
// MountainCar -- reference: src/envs/classical_control/mountain_car.rs:398-435
// ---------------------------------------------------------------------------
struct MountainCarP {
    float force;       // :350
    float neg_gravity; // -gravity (:351, :412)
    float max_speed;   // :346
    float min_position, max_position; // :344-345
    float goal_position; // smallest float >= goal_position: (v >= it) == (v >= 0.5 in f64)
    float goal_velocity; // same rounding
    ResetBox rb;       // :176-187 position in [-0.6, -0.4); velocity is always 0 (:162-167)
};

// clip (util_fns.rs:2-10) on OrderedFloat values: in range -> value, greater than the right bound
// -> right bound, else left bound.  OrderedFloat's total order puts NaN above everything, so a
// NaN value is clipped to the RIGHT bound (fminf / fmaxf would return the left one).
__device__ __forceinline__ float clip_of(float v, float lo, float hi)
{
    v = !(v <= hi) ? hi : v;
    return v < lo ? lo : v;
}

struct MountainCar {
    static constexpr int SD = 2, OD = 2;
    static constexpr bool OBS_IS_STATE = true;
    static constexpr bool HAS_SBT = false;
    using Action = int32_t;
    using P = MountainCarP;

    __device__ __forceinline__ static bool valid(Action a) { return (uint32_t)a < 3u; }

    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&)[OD], float &reward, bool &done)
    {
        float position = s[0], velocity = s[1];
        // velocity += (a - 1) * force + cos(3 * position) * (-gravity)               :411-412
        const float rhs = fmaf(cosf(3.0f * position), p.neg_gravity, (float)(a - 1) * p.force);
        velocity += rhs;
        velocity = clip_of(velocity, -p.max_speed, p.max_speed); //                  :413
        position += velocity; //                                                      :415
        position = clip_of(position, p.min_position, p.max_position); //              :416
        // exact equality with the clipped wall value                                 :418-420
        if (position == p.min_position && velocity < 0.0f) velocity = 0.0f;
        done = (position >= p.goal_position) & (velocity >= p.goal_velocity); //      :422
        reward = -1.0f; //                                                            :423
        s[0] = position; s[1] = velocity;
    }

    __device__ __forceinline__ static void reset(const P &p, float (&s)[SD], float (&)[OD], uint4 w)
    {
        s[0] = uniform_from_word(w.x, p.rb.low[0], p.rb.scale24[0], p.rb.cap[0]);
        s[1] = 0.0f;
    }
};

// ---------------------------------------------------------------------------
// Pendulum-v1 -- NOT in the reference; upstream OpenAI Gym pendulum.py
// (SURVEY.md Appendix D).  State (theta, theta_dot), obs (cos, sin, theta_dot).
// ---------------------------------------------------------------------------
struct PendulumP {
    float max_speed, max_torque, dt;
    float c_sin; // 3 g / (2 l)
    float c_u;   // 3 / (m l^2)
    ResetBox rb; // theta in [-pi, pi), theta_dot in [-1, 1)
};

// x - 2 pi * floor((x + pi) / (2 pi)) with a two-constant 2 pi (result in [-pi, pi) up to rounding)
__device__ __forceinline__ float angle_normalize(float x)
{
    const float k = floorf(fmaf(x, 0.15915494309189535f, 0.5f));
    float r = fmaf(k, -6.2831854820251465f, x);   // 2 pi, high part (exact float)
    return fmaf(k, 1.7484555e-07f, r);            // minus the low part (2pi_hi - 2pi)
}

struct Pendulum {
    static constexpr int SD = 2, OD = 3;
    static constexpr bool OBS_IS_STATE = false;
    static constexpr bool HAS_SBT = false;
    using Action = float;
    using P = PendulumP;

    __device__ __forceinline__ static bool valid(Action) { return true; } // Box action: clipped, never rejected

    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&o)[OD], float &reward, bool &done)
    {
        const float th = s[0], thdot = s[1];
        const float u = fminf(fmaxf(a, -p.max_torque), p.max_torque);
        const float an = angle_normalize(th);
        // cost uses the PRE-update theta, theta_dot and the clipped torque
        const float costs = fmaf(an, an, fmaf(0.1f * thdot, thdot, 0.001f * (u * u)));
        float newthdot = fmaf(fmaf(p.c_sin, sinf(th), p.c_u * u), p.dt, thdot);
        newthdot = fminf(fmaxf(newthdot, -p.max_speed), p.max_speed);
        float newth = fmaf(newthdot, p.dt, th); // uses the clipped NEW velocity
        // Stored theta is kept wrapped to [-pi, pi): observation and cost are invariant under
        // the wrap, and f32 cos/sin stay accurate on arbitrarily long spins.
        if (fabsf(newth) > 3.14159274f) newth = angle_normalize(newth);
        float sn, cs;
        sincosf(newth, &sn, &cs);
        s[0] = newth; s[1] = newthdot;
        o[0] = cs; o[1] = sn; o[2] = newthdot;
        reward = -costs;
        done = false; // never terminates
    }

    __device__ __forceinline__ static void reset(const P &p, float (&s)[SD], float (&o)[OD], uint4 w)
    {
        s[0] = uniform_from_word(w.x, p.rb.low[0], p.rb.scale24[0], p.rb.cap[0]);
        s[1] = uniform_from_word(w.y, p.rb.low[1], p.rb.scale24[1], p.rb.cap[1]);
        float sn, cs;
        sincosf(s[0], &sn, &cs);
        o[0] = cs; o[1] = sn; o[2] = s[1];
    }
};

} // namespace gymrs
