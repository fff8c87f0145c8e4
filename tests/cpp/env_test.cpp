// env_test.cpp -- the reference's own unit tests restated against the C++ host mirror
// (include/gym_rs.hpp) plus a BASELINE-config-1 style plumbing run (one CartPole env, random
// actions, RenderMode::None) cross-checked step by step with the f64 oracle.
//
// Reference tests restated: src/utils/custom/util_fns.rs:16-32 (clip), src/spaces/discrete.rs:27-41
// (Discrete::contains), src/utils/seeding.rs:33-39 (seed echo).  Needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "gym_rs.hpp"
#include "gymrs_oracle.h"

using namespace gym_rs;
using gym_rs::envs::classical_control::cartpole::CartPoleEnv;
using gym_rs::envs::classical_control::cartpole::CartPoleObservation;
using gym_rs::envs::classical_control::mountain_car::MountainCarEnv;
using gym_rs::envs::classical_control::mountain_car::MountainCarObservation;

static int failures = 0;
#define EXPECT(cond)                                                             \
    do {                                                                         \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static bool close6(double gpu, double ref) { return std::fabs(gpu - ref) <= 1e-6 * std::fmax(1.0, std::fabs(ref)); }

static void reference_unit_tests()
{
    // util_fns.rs:16-32
    EXPECT(utils::custom::clip(2, 0, 1) == 1);
    EXPECT(utils::custom::clip(-1, 0, 1) == 0);
    EXPECT(utils::custom::clip(1, -1, 2) == 1);
    // discrete.rs:27-41
    spaces::Discrete obj{3};
    EXPECT(!obj.contains(3) && !obj.contains(4));
    EXPECT(obj.contains(1) && obj.contains(2));
    // seeding.rs:33-39 and the doctest :11-20
    EXPECT(utils::seeding::rand_random(42).second == 42);
    EXPECT(utils::seeding::rand_random(64).second == 64);
}

static void cartpole_plumbing()
{
    CartPoleEnv env(utils::RenderMode::None);
    EXPECT(env.action_space == spaces::Discrete{2});
    EXPECT(env.observation_space.high.x == 4.8 && std::isinf(env.observation_space.high.x_dot));
    EXPECT(env.observation_space.high.theta == 0.41887902047863906);
    EXPECT(std::isinf(env.reward_range().upper_bound));
    std::mt19937 rng(0);
    int total_steps = 0;
    for (int ep = 0; ep < 15; ++ep) {
        auto [state, info] = env.reset((uint64_t)ep, true, std::nullopt);
        EXPECT(info.has_value());
        EXPECT(std::fabs(state.x) <= 0.05 && std::fabs(state.theta_dot) <= 0.05);
        orc_cartpole_env o;
        orc_cartpole_new(&o);
        double current_reward = 0.;
        for (int t = 0; t < 475; ++t) {
            auto v = env.state.to_vec();
            for (int k = 0; k < 4; ++k) o.state[k] = v[k]; // per-step resync with the device's f32 state
            size_t action = rng() & 1u;
            auto sr = env.step(action);
            double r; int d, tr;
            orc_cartpole_step(&o, action, &r, &d, &tr);
            auto g = sr.observation.to_vec();
            for (int k = 0; k < 4; ++k) EXPECT(close6(g[k], o.state[k]));
            EXPECT(sr.reward == r && sr.done == (d != 0) && !sr.truncated && sr.info.has_value());
            current_reward += sr.reward;
            ++total_steps;
            if (sr.done) { EXPECT(env.steps_beyond_terminated() == std::optional<size_t>(0)); break; }
        }
        EXPECT(current_reward >= 5. && current_reward <= 475.);
    }
    EXPECT(total_steps >= 150);
    // stepping after termination: reward 1.0 on the first done step, 0.0 afterwards (cartpole.rs:455-464)
    env.reset(3, false, std::nullopt);
    double last = 1.0; bool seen_done = false;
    static int warnings = 0; // ... and the reference logs a warning for each of those steps (log::warn!, :461)
    auto keep = CartPoleEnv::warn_sink();
    CartPoleEnv::warn_sink() = [](const char *m) { if (std::string(m).find("after termination") != std::string::npos) ++warnings; };
    int late_steps = 0;
    for (int t = 0; t < 80; ++t) {
        auto sr = env.step(1);
        if (seen_done) { last = sr.reward; ++late_steps; }
        seen_done |= sr.done;
    }
    CartPoleEnv::warn_sink() = keep;
    EXPECT(seen_done && last == 0.0);
    EXPECT(late_steps > 0 && warnings == late_steps);
    // invalid action: the reference's assert! message (cartpole.rs:402-406)
    bool threw = false;
    try { env.step(2); } catch (const Panic &e) { threw = std::string(e.what()) == "2 usize invalid"; }
    EXPECT(threw);
    // Clone is a deep copy (core.rs:25)
    CartPoleEnv twin(env);
    auto a = env.step(0), b = twin.step(0);
    EXPECT(a.observation.to_vec() == b.observation.to_vec());
    // Serialize: checkpoint, diverge, restore, and the restored env retraces the same steps
    auto blob = env.checkpoint();
    std::vector<double> first = env.step(1).observation.to_vec();
    env.step(0);
    env.restore(blob);
    EXPECT(env.step(1).observation.to_vec() == first);
    gymrs_checkpoint_info info;
    EXPECT(gymrs_checkpoint_info_of(blob.data(), blob.size(), &info) == GYMRS_OK && info.kind == GYMRS_CARTPOLE && info.num_envs == 1);
    blob[blob.size() / 2] ^= 1;
    EXPECT(gymrs_checkpoint_info_of(blob.data(), blob.size(), &info) == GYMRS_ERR_BAD_ARG);
    // options: BoxR bounds for this reset only (cartpole.rs:351-365)
    spaces::BoxR<CartPoleObservation> box{{1, 2, 3, 4}, {2, 3, 4, 5}};
    auto st = env.reset(1, false, box).first;
    EXPECT(st.x >= 1 && st.x < 2 && st.theta_dot >= 4 && st.theta_dot < 5);
    env.close();
}

static void mountain_car_plumbing()
{
    MountainCarEnv mc(utils::RenderMode::None);
    EXPECT(mc.action_space == spaces::Discrete{3});
    EXPECT(mc.observation_space.low.position == -1.2 && mc.observation_space.high.velocity == 0.07);
    auto st = mc.reset(0, false, std::nullopt).first;
    EXPECT(st.position >= -0.6 && st.position < -0.4 && st.velocity == 0.0);
    orc_mountain_car_env o;
    orc_mountain_car_new(&o);
    std::mt19937 rng(1);
    for (int t = 0; t < 200; ++t) {
        o.state[0] = mc.state.position; o.state[1] = mc.state.velocity;
        size_t action = rng() % 3;
        auto sr = mc.step(action);
        double r; int d, tr;
        orc_mountain_car_step(&o, action, &r, &d, &tr);
        EXPECT(close6(sr.observation.position, o.state[0]) && close6(sr.observation.velocity, o.state[1]));
        EXPECT(sr.reward == -1.0 && !sr.info.has_value() && !sr.truncated);
    }
    bool threw = false;
    try { mc.step(3); } catch (const Panic &e) { threw = std::string(e.what()) == "3 (usize) invalid"; }
    EXPECT(threw);
}

static void batched_plumbing()
{
    using gym_rs::batched::BatchedEnv;
    using gym_rs::batched::Kind;
    const size_t n = 5000;
    BatchedEnv env(Kind::CartPole, n, 0, 100);
    EXPECT(env.reset(7) == 7 && env.obs_dim() == 4);
    std::vector<size_t> act(n);
    std::mt19937 rng(3);
    // sampled envs against their own scalar oracle object, resynchronised each step
    std::vector<float> before = env.step(std::vector<size_t>(n, 1), false).observation; // a copy: the views are reused
    for (int t = 0; t < 20; ++t) {
        for (auto &a : act) a = rng() % 2;
        auto st = env.step(act, false);
        for (size_t i = 0; i < n; i += 97) {
            orc_cartpole_env o;
            orc_cartpole_new(&o);
            for (int k = 0; k < 4; ++k) o.state[k] = before[k * n + i];
            double r; int d, tr;
            orc_cartpole_step(&o, act[i], &r, &d, &tr);
            for (int k = 0; k < 4; ++k) EXPECT(close6(st.observation[k * n + i], o.state[k]));
            const bool in_band = std::fabs(std::fabs(o.state[0]) - 2.4) < 1e-6 ||
                                 std::fabs(std::fabs(o.state[2]) - 0.20943951023931953) < 1e-6;
            EXPECT(in_band || (st.done[i] != 0) == (d != 0));
        }
        before = st.observation;
    }
    // auto-reset keeps every env inside the live region; checkpoint / restore retraces the same step
    for (int t = 0; t < 50; ++t) {
        for (auto &a : act) a = rng() % 2;
        env.step(act, true);
    }
    auto blob = env.checkpoint();
    std::vector<float> first = env.step(act, true).observation;
    env.step(act, true);
    env.restore(blob);
    EXPECT(env.step(act, true).observation == first);
    for (size_t i = 0; i < n; ++i) EXPECT(std::fabs(first[i]) <= 2.4f + 1e-3f);
    // an invalid action anywhere in the batch panics with the reference's message
    act[1234] = 2;
    bool threw = false;
    try { env.step(act, true); } catch (const Panic &e) { threw = std::string(e.what()) == "2 usize invalid"; }
    EXPECT(threw);
    // Pendulum takes float torques
    BatchedEnv pd(Kind::Pendulum, 1000);
    pd.reset(1);
    auto ps = pd.step(std::vector<float>(1000, 0.5f), true);
    EXPECT(pd.obs_dim() == 3 && ps.reward[0] <= 0.0f && ps.done[0] == 0);
}

int main()
{
    reference_unit_tests();
    if (gymrs_device_count() <= 0) {
        // no CPU fallback: construction must fail loudly
        bool threw = false;
        try { CartPoleEnv env; } catch (const Panic &) { threw = true; }
        EXPECT(threw);
        std::printf("no CUDA device: unit tests only, failures=%d\n", failures);
        return failures ? 1 : 0;
    }
    cartpole_plumbing();
    mountain_car_plumbing();
    batched_plumbing();
    std::printf("env_test: failures=%d\n", failures);
    return failures ? 1 : 0;
}
