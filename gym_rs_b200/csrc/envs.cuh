// envs.cuh -- per-env f32 dynamics for the three classic-control envs (device side).
//
// Each struct is a policy class consumed by the kernels in kernels.cu:
//   SD / OD          rows of state / observation
//   OBS_IS_STATE     the observation IS the post-step state (CartPole, MountainCar)
//   HAS_SBT          keeps the reference's steps_beyond_terminated (CartPole)
//   WIDE_MIN_CTAS    register budget of the opt-in high-occupancy step kernel, as the minimum number of
//                    256-thread CTAs per SM in __launch_bounds__: 8 = 32 registers per thread (every CTA
//                    of a 1M-env launch is resident at once), 6 = 40 registers; the tightest budget
//                    ptxas meets without spilling on the default (auto-reset) path
//   Action, P        action element type, by-value parameter block
//   valid()          Space::contains on the action      (spaces/discrete.rs:14-20)
//   pre()            the per-env action term the dynamics consume (force, push, clamped torque)
//   fast_ok()        this env may take the straight-line path (small angle / bounded argument)
//   advance<Lane>()  the dynamics, written once against an arithmetic lane (lanes.cuh): Lane2 steps
//                    TWO envs per instruction with packed fp32 (FFMA2), Lane1 is the scalar form;
//                    both round identically
//   advance_slow()   scalar, any argument (full-range libm trig)
//   terminal()       done flag of the post-step state
//   step()           pre + advance (fast or slow) + terminal for one env on registers
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

#include "lanes.cuh"
#include "philox.cuh"

namespace gymrs {

struct ResetBox {
    float low[4];   // lower bound per state row
    float scale24[4]; // (high - low) * 2^-24: maps the 24 random bits straight onto [low, high)
    float cap[4];   // largest float below high
};

// num / den on a lane: MUFU.RCP plus one Newton correction on the quotient (result within 1 ulp for
// normal operands) instead of the ~10-instruction IEEE division sequence with its slow-path call.
// A zero / non-finite den yields inf or NaN (never a finite wrong value).
template <class L>
__device__ __forceinline__ typename L::T div_newton(typename L::T num, typename L::T den)
{
    const typename L::T r = L::rcp(den);
    const typename L::T q = L::mul(num, r);
    return L::fma(L::fma(L::neg(den), q, num), r, q);
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
    float neg_den_b;    // -(length * masspole / total_mass)
    float x_thr;        // largest float <= x_threshold: (v > x_thr) == (v > 2.4 in f64) for every float v
    float th_thr;       // same for theta_threshold_radians
    int semi_implicit;  // KinematicsIntegrator::Other (:437-441)
    ResetBox rb;        // :353-361 (defaults +-0.05) or caller's BoxR
};

struct CartPole {
    static constexpr int SD = 4, OD = 4;
    static constexpr bool OBS_IS_STATE = true;
    static constexpr bool HAS_SBT = true;
    static constexpr int WIDE_MIN_CTAS = 6;
    using Action = int32_t;
    using P = CartPoleP;

    __device__ __forceinline__ static bool valid(Action a) { return (uint32_t)a < 2u; }

    // force / M with the action's sign                                              :414-418
    __device__ __forceinline__ static float pre(const P &p, Action a) { return (a == 1) ? p.force_over_m : -p.force_over_m; }

    // A live CartPole has |theta| <= 0.21 (+ one step), so the common case needs neither range
    // reduction nor quadrant selection; anything larger (a pole that keeps falling when the caller
    // never resets) takes the full-range sincosf.
    __device__ __forceinline__ static bool fast_ok(const P &, const float (&s)[SD]) { return fabsf(s[2]) <= 0.78539816f; }

    // the equations of motion given sin / cos of the pole angle; s = (x, x_dot, theta, theta_dot)
    template <class L>
    __device__ __forceinline__ static void dynamics(const P &p, typename L::T (&s)[SD], typename L::T f,
                                                    typename L::T sn, typename L::T cs)
    {
        using T = typename L::T;
        const T x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        // temp = (force + PML * theta_dot^2 * sin) / M                               :423-424
        const T temp = L::fma(L::mul(L::mul(theta_dot, theta_dot), L::bc(p.pml_over_m)), sn, f);
        // thetaacc = (g sin - cos temp) / (l (4/3 - mp cos^2 / M))                   :425-428
        const T num = L::fma(L::bc(p.gravity), sn, L::neg(L::mul(cs, temp)));
        const T den = L::fma(L::bc(p.neg_den_b), L::mul(cs, cs), L::bc(p.den_a));
        const T thetaacc = div_newton<L>(num, den);
        // xacc = temp - PML thetaacc cos / M                                         :429
        const T xacc = L::fma(L::neg(L::mul(L::bc(p.pml_over_m), thetaacc)), cs, temp);
        const T tau = L::bc(p.tau);
        // Euler: positions advance with the OLD velocities (:431-436); semi-implicit ("Other"): with the
        // NEW ones (:437-441).  The velocities are the same FMA either way, so the integrator only selects
        // which velocity the position FMA reads: a uniform select instead of two copies of the update
        // (a branch made ptxas shuffle eight registers per pair to reconverge).
        const T x_dot_new = L::fma(tau, xacc, x_dot);
        const T theta_dot_new = L::fma(tau, thetaacc, theta_dot);
        const bool semi = p.semi_implicit != 0;
        s[0] = L::fma(tau, semi ? x_dot_new : x_dot, x);
        s[2] = L::fma(tau, semi ? theta_dot_new : theta_dot, theta);
        s[1] = x_dot_new;
        s[3] = theta_dot_new;
    }

    // fast path: |theta| <= pi/4 for every env of the lane                          :420-421
    template <class L>
    __device__ __forceinline__ static void advance(const P &p, typename L::T (&s)[SD], typename L::T f,
                                                   typename L::T (&)[OD], typename L::T &reward)
    {
        typename L::T sn, cs;
        sincos_poly<L>(s[2], sn, cs); // polynomial, never the MUFU __sinf
        dynamics<L>(p, s, f, sn, cs);
        reward = L::bc(1.0f); // decided by the caller beyond the first terminal step (:455-464)
    }

    __device__ __forceinline__ static void advance_slow(const P &p, float (&s)[SD], float f, float (&)[OD], float &reward)
    {
        float sn, cs;
        sincosf(s[2], &sn, &cs);
        dynamics<Lane1>(p, s, f, sn, cs);
        reward = 1.0f;
    }

    // strict comparisons on the updated x, theta                                     :450-453
    // (x < -T || x > T) == (|x| > T).  The reference compares OrderedFloat values, a total
    // order in which NaN is greater than everything, so a NaN state IS done: !(|x| <= T).
    __device__ __forceinline__ static bool terminal(const P &p, const float (&s)[SD])
    {
        return !(fabsf(s[0]) <= p.x_thr) | !(fabsf(s[2]) <= p.th_thr);
    }

    // one env, scalar: the same bits as a Lane2 pair
    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&o)[OD], float &reward, bool &done)
    {
        const float f = pre(p, a);
        if (fast_ok(p, s)) advance<Lane1>(p, s, f, o, reward);
        else advance_slow(p, s, f, o, reward);
        done = terminal(p, s);
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
// fminf returns its non-NaN operand, so min-then-max is exactly that for every input (two FMNMX).
__device__ __forceinline__ float clip_of(float v, float lo, float hi)
{
    return fmaxf(fminf(v, hi), lo);
}

struct MountainCar {
    static constexpr int SD = 2, OD = 2;
    static constexpr bool OBS_IS_STATE = true;
    static constexpr bool HAS_SBT = false;
    static constexpr int WIDE_MIN_CTAS = 8;
    using Action = int32_t;
    using P = MountainCarP;

    __device__ __forceinline__ static bool valid(Action a) { return (uint32_t)a < 3u; }

    // (a - 1) * force                                                               :411
    __device__ __forceinline__ static float pre(const P &p, Action a) { return __fmul_rn((float)(a - 1), p.force); }

    // the position is clipped into [min_position, max_position] by every step, so 3 p stays tiny;
    // a caller-injected far-out (or NaN) position takes the libm cosf
    __device__ __forceinline__ static bool fast_ok(const P &, const float (&s)[SD]) { return fabsf(s[0]) <= TRIG_FAST_MAX / 4.0f; }

    // everything after cos(3 p); s = (position, velocity)
    template <class L>
    __device__ __forceinline__ static void dynamics(const P &p, typename L::T (&s)[SD], typename L::T push, typename L::T c3p)
    {
        using T = typename L::T;
        // velocity += (a - 1) * force + cos(3 * position) * (-gravity): RHS first      :411-412
        const T rhs = L::fma(c3p, L::bc(p.neg_gravity), push);
        T velocity = L::fma(rhs, L::bc(1.0f), s[1]);
        velocity = L::map(velocity, [&](float v, int) { return clip_of(v, -p.max_speed, p.max_speed); }); // :413
        T position = L::fma(velocity, L::bc(1.0f), s[0]); //                              :415
        position = L::map(position, [&](float v, int) { return clip_of(v, p.min_position, p.max_position); }); // :416
        // exact equality with the clipped wall value                                 :418-420
        velocity = L::map2(velocity, position, [&](float v, float x, int) { return (x == p.min_position && v < 0.0f) ? 0.0f : v; });
        s[0] = position;
        s[1] = velocity;
    }

    template <class L>
    __device__ __forceinline__ static void advance(const P &p, typename L::T (&s)[SD], typename L::T push,
                                                   typename L::T (&)[OD], typename L::T &reward)
    {
        typename L::T sn, cs;
        sincos_reduced<L>(L::mul(s[0], L::bc(3.0f)), sn, cs);
        dynamics<L>(p, s, push, cs);
        reward = L::bc(-1.0f); //                                                     :423
    }

    __device__ __forceinline__ static void advance_slow(const P &p, float (&s)[SD], float push, float (&)[OD], float &reward)
    {
        dynamics<Lane1>(p, s, push, cosf(__fmul_rn(s[0], 3.0f)));
        reward = -1.0f;
    }

    __device__ __forceinline__ static bool terminal(const P &p, const float (&s)[SD])
    {
        return (s[0] >= p.goal_position) & (s[1] >= p.goal_velocity); //              :422
    }

    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&o)[OD], float &reward, bool &done)
    {
        const float push = pre(p, a);
        if (fast_ok(p, s)) advance<Lane1>(p, s, push, o, reward);
        else advance_slow(p, s, push, o, reward);
        done = terminal(p, s);
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
template <class L>
__device__ __forceinline__ typename L::T angle_normalize(typename L::T x)
{
    const typename L::T k = L::floor(L::fma(x, L::bc(0.15915494309189535f), L::bc(0.5f)));
    const typename L::T r = L::fma(k, L::bc(-6.2831854820251465f), x); // 2 pi, high part (exact float)
    return L::fma(k, L::bc(1.7484555e-07f), r);                         // minus the low part (2pi_hi - 2pi)
}

struct Pendulum {
    static constexpr int SD = 2, OD = 3;
    static constexpr bool OBS_IS_STATE = false;
    static constexpr bool HAS_SBT = false;
    static constexpr int WIDE_MIN_CTAS = 8;
    using Action = float;
    using P = PendulumP;

    __device__ __forceinline__ static bool valid(Action) { return true; } // Box action: clipped, never rejected

    // the clipped torque
    __device__ __forceinline__ static float pre(const P &p, Action a) { return fminf(fmaxf(a, -p.max_torque), p.max_torque); }

    // the stored angle is kept wrapped to [-pi, pi); a caller-injected far-out angle takes libm
    __device__ __forceinline__ static bool fast_ok(const P &, const float (&s)[SD]) { return fabsf(s[0]) <= TRIG_FAST_MAX / 2.0f; }

    // s = (theta, theta_dot), u = clipped torque, sin_th = sin(theta); o = (cos, sin, theta_dot) of the
    // new state is filled by the caller from the returned new angle
    template <class L>
    __device__ __forceinline__ static void dynamics(const P &p, typename L::T (&s)[SD], typename L::T u,
                                                    typename L::T sin_th, typename L::T &reward)
    {
        using T = typename L::T;
        const T th = s[0], thdot = s[1];
        const T an = angle_normalize<L>(th);
        // cost uses the PRE-update theta, theta_dot and the clipped torque
        const T costs = L::fma(an, an, L::fma(L::mul(thdot, L::bc(0.1f)), thdot, L::mul(L::mul(u, u), L::bc(0.001f))));
        T newthdot = L::fma(L::fma(L::bc(p.c_sin), sin_th, L::mul(u, L::bc(p.c_u))), L::bc(p.dt), thdot);
        newthdot = L::map(newthdot, [&](float v, int) { return fminf(fmaxf(v, -p.max_speed), p.max_speed); });
        // uses the clipped NEW velocity.  Stored theta is kept wrapped to [-pi, pi): observation and
        // cost are invariant under the wrap, and f32 cos/sin stay accurate on arbitrarily long spins
        // (inside the range the wrap is the identity: k = 0).
        s[0] = angle_normalize<L>(L::fma(newthdot, L::bc(p.dt), th));
        s[1] = newthdot;
        reward = L::neg(costs);
    }

    template <class L>
    __device__ __forceinline__ static void advance(const P &p, typename L::T (&s)[SD], typename L::T u,
                                                   typename L::T (&o)[OD], typename L::T &reward)
    {
        typename L::T sn, cs;
        sincos_reduced<L>(s[0], sn, cs);
        dynamics<L>(p, s, u, sn, reward);
        sincos_reduced<L>(s[0], sn, cs); // |new theta| <= pi
        o[0] = cs; o[1] = sn; o[2] = s[1];
    }

    __device__ __forceinline__ static void advance_slow(const P &p, float (&s)[SD], float u, float (&o)[OD], float &reward)
    {
        dynamics<Lane1>(p, s, u, sinf(s[0]), reward);
        float sn, cs;
        sincos_reduced<Lane1>(s[0], sn, cs);
        o[0] = cs; o[1] = sn; o[2] = s[1];
    }

    __device__ __forceinline__ static bool terminal(const P &, const float (&)[SD]) { return false; } // never terminates

    __device__ __forceinline__ static void step(const P &p, float (&s)[SD], Action a,
                                                float (&o)[OD], float &reward, bool &done)
    {
        const float u = pre(p, a);
        if (fast_ok(p, s)) advance<Lane1>(p, s, u, o, reward);
        else advance_slow(p, s, u, o, reward);
        done = false;
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
