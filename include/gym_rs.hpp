// gym_rs.hpp -- header-only C++ host mirror of gym-rs's Env surface over the C ABI
// (include/gymrs_b200.h).  The reference is compiled code (Rust) and this image has no Rust
// toolchain, so this is the compiled-language host side: same names, argument meaning and error
// behaviour as the crate, so that tests read like the reference's own.
//
//   gym_rs::core::ActionReward<T, E>            src/core.rs:94-106
//   gym_rs::core::RewardRange                   src/core.rs:109-122
//   gym_rs::spaces::Discrete / BoxR<T>          src/spaces/discrete.rs:12-20, box_r.rs:5-13
//   gym_rs::utils::custom::clip                 src/utils/custom/util_fns.rs:2-10
//   gym_rs::utils::seeding::rand_random         src/utils/seeding.rs:21-26
//   gym_rs::envs::classical_control::cartpole::CartPoleEnv       cartpole.rs:51-87, :389-516
//   gym_rs::envs::classical_control::mountain_car::MountainCarEnv mountain_car.rs:46-84, :391-501
//   gym_rs::batched::BatchedEnv                 N instances per handle (no counterpart: F9 in SURVEY.md)
//
// The reference panics on an invalid action (assert!, cartpole.rs:402-406); here that is a
// gym_rs::Panic exception carrying the reference's message text.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gymrs_b200.h"

namespace gym_rs {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int rc)
{
    if (rc != GYMRS_OK) throw Panic(std::string("gymrs error ") + std::to_string(rc) + ": " + gymrs_last_error());
}

namespace utils {
enum class RenderMode { Human, SingleRgbArray, RgbArray, DepthArray, SingleDepthArray, None }; // renderer.rs:83-114
enum class Renders { None };                                                                  // renderer.rs:118-130
namespace custom {
template <class T> T clip(T value, T left_bound, T right_bound) // util_fns.rs:2-10, same branch order
{
    if (left_bound <= value && value <= right_bound) return value;
    else if (value > right_bound) return right_bound;
    else return left_bound;
}
inline double clip(double v, double lo, double hi) { return gymrs_clip(v, lo, hi); }
} // namespace custom
namespace seeding {
// returns (generator key, seed used); the generator is counter-based Philox keyed by the seed
inline std::pair<uint64_t, uint64_t> rand_random(std::optional<uint64_t> seed)
{
    uint64_t s = seed ? *seed : 0;
    uint64_t used = gymrs_rand_random(seed ? &s : nullptr);
    return {used, used};
}
} // namespace seeding
} // namespace utils

namespace spaces {
struct Discrete { // discrete.rs:12
    uint64_t n;
    bool contains(uint64_t value) const { return gymrs_discrete_contains(n, value) != 0; } // :14-20
    bool operator==(const Discrete &o) const { return n == o.n; }
};
template <class T> struct BoxR { // box_r.rs:5-13 (no contains in the reference either)
    T low, high;
};
} // namespace spaces

namespace core {
template <class T, class E> struct ActionReward { // core.rs:94-106
    T observation;
    double reward;
    bool done;
    bool truncated;
    std::optional<E> info;
};
struct RewardRange { // core.rs:109-122, default (-inf, inf) :16-19
    double lower_bound = -std::numeric_limits<double>::infinity();
    double upper_bound = std::numeric_limits<double>::infinity();
};
struct Unit {
    bool operator==(const Unit &) const { return true; }
};

// Shared implementation of one env object (a batch-of-1 handle) behind the Env surface.
template <class Derived, class Obs, int KIND, int DIM> class EnvBase {
  public:
    using Action = size_t;
    using Observation = Obs;
    using Info = Unit;
    using ResetInfo = Unit;

    explicit EnvBase(utils::RenderMode render_mode = utils::RenderMode::None) : render_mode(render_mode)
    {
        if (render_mode != utils::RenderMode::None) throw Panic("only RenderMode::None is supported (rendering is out of scope)");
        check(gymrs_create(KIND, 1, 0, 0, nullptr, 0, &handle_));
        pull_state();
    }
    EnvBase(const EnvBase &o) : render_mode(o.render_mode), state(o.state) { check(gymrs_clone(o.handle_, &handle_)); }
    EnvBase &operator=(const EnvBase &) = delete;
    ~EnvBase() { close(); }

    // where the log::warn! of the reference goes (default: stderr); set to nullptr to silence
    using WarnSink = void (*)(const char *);
    static WarnSink &warn_sink()
    {
        static WarnSink sink = [](const char *m) { std::fprintf(stderr, "WARN gym_rs: %s\n", m); };
        return sink;
    }

    // core.rs:42
    ActionReward<Obs, Unit> step(Action action)
    {
        if (!static_cast<Derived *>(this)->action_space.contains(action))
            throw Panic(Derived::invalid_action_message(action)); // assert!, cartpole.rs:402-406
        int32_t a = (int32_t)action;
        float obs[DIM], reward = 0;
        uint8_t done = 0, truncated = 0;
        check(gymrs_step_host(handle_, &a, 0, obs, &reward, &done, &truncated));
        check(gymrs_sync(handle_, nullptr));
        pull_state();
        // reward 0 on a terminal step means the env had already terminated before this call: the
        // reference logs a warning there (log::warn!, cartpole.rs:461)
        if (Derived::WARNS_AFTER_TERMINATION && done && reward == 0.0f && warn_sink())
            warn_sink()("Calling step after termination may result in undefined behaviour. Consider reseting.");
        return {Derived::make_obs(obs), (double)reward, done != 0, truncated != 0, Derived::step_info()};
    }

    // core.rs:45-50
    std::pair<Obs, std::optional<Unit>> reset(std::optional<uint64_t> seed, bool return_info,
                                              std::optional<spaces::BoxR<Obs>> options)
    {
        uint64_t s = seed ? *seed : 0;
        float lo[DIM], hi[DIM];
        if (options) {
            auto l = options->low.to_vec(), h = options->high.to_vec();
            for (int i = 0; i < DIM; ++i) { lo[i] = (float)l[i]; hi[i] = (float)h[i]; }
        }
        check(gymrs_reset(handle_, seed ? &s : nullptr, options ? lo : nullptr, options ? hi : nullptr, nullptr, &seed_used_));
        pull_state();
        return {state, return_info ? std::optional<Unit>(Unit{}) : std::nullopt};
    }

    utils::Renders render(utils::RenderMode) { return utils::Renders::None; } // renderer.rs:52-61 under None
    void close()
    {
        if (handle_) gymrs_destroy(handle_);
        handle_ = nullptr;
    }

    // `Env: Serialize` (core.rs:25): the handle as one blob, reset stream included, so restore()
    // resumes bit-identically (the reference's serde derive skips its RNG, cartpole.rs:85-86)
    std::vector<unsigned char> checkpoint()
    {
        size_t bytes = 0;
        check(gymrs_checkpoint_size(handle_, &bytes));
        std::vector<unsigned char> blob(bytes);
        check(gymrs_checkpoint_save(handle_, blob.data(), bytes));
        return blob;
    }
    void restore(const std::vector<unsigned char> &blob)
    {
        check(gymrs_checkpoint_load(handle_, blob.data(), blob.size()));
        pull_state();
    }

    // EnvProperties, core.rs:60-90
    RewardRange reward_range() const { return {}; }
    uint64_t rand_random() const { return seed_used_; }

    utils::RenderMode render_mode;
    Obs state; // `pub state`, cartpole.rs:60 / mountain_car.rs:74
    gymrs_env *handle() { return handle_; }

  protected:
    void pull_state()
    {
        float s[DIM];
        check(gymrs_get_state(handle_, s, nullptr));
        state = Derived::make_obs(s);
    }
    gymrs_env *handle_ = nullptr;
    uint64_t seed_used_ = 0;
};
} // namespace core

namespace envs::classical_control {
namespace cartpole {
struct CartPoleObservation { // cartpole.rs:327-334
    double x = 0, x_dot = 0, theta = 0, theta_dot = 0;
    std::vector<double> to_vec() const { return {x, x_dot, theta, theta_dot}; } // Vec<f64>::from, :336-349
    CartPoleObservation operator-() const { return {-x, -x_dot, -theta, -theta_dot}; } // :367-378
};
class CartPoleEnv : public core::EnvBase<CartPoleEnv, CartPoleObservation, GYMRS_CARTPOLE, 4> {
  public:
    using Base = core::EnvBase<CartPoleEnv, CartPoleObservation, GYMRS_CARTPOLE, 4>;
    explicit CartPoleEnv(utils::RenderMode m = utils::RenderMode::None) : Base(m)
    {
        double lo[4], hi[4];
        check(gymrs_observation_space(handle_, lo, hi));
        observation_space = {{lo[0], lo[1], lo[2], lo[3]}, {hi[0], hi[1], hi[2], hi[3]}}; // :105-113
    }
    spaces::Discrete action_space{2}; // :112
    spaces::BoxR<CartPoleObservation> observation_space;
    // steps_beyond_terminated: None / Some(k), cartpole.rs:81
    std::optional<size_t> steps_beyond_terminated()
    {
        float s[4];
        int32_t sbt = -1;
        check(gymrs_get_state(handle_, s, &sbt));
        return sbt < 0 ? std::nullopt : std::optional<size_t>((size_t)sbt);
    }
    static CartPoleObservation make_obs(const float *v) { return {v[0], v[1], v[2], v[3]}; }
    static constexpr bool WARNS_AFTER_TERMINATION = true; // :455-464
    static std::optional<core::Unit> step_info() { return core::Unit{}; } // info: Some(()), :481
    static std::string invalid_action_message(size_t a) { return std::to_string(a) + " usize invalid"; } // :404
};
} // namespace cartpole

namespace mountain_car {
struct MountainCarObservation { // mountain_car.rs:121-128
    double position = 0, velocity = 0;
    std::vector<double> to_vec() const { return {position, velocity}; } // :193-197
};
class MountainCarEnv : public core::EnvBase<MountainCarEnv, MountainCarObservation, GYMRS_MOUNTAIN_CAR, 2> {
  public:
    using Base = core::EnvBase<MountainCarEnv, MountainCarObservation, GYMRS_MOUNTAIN_CAR, 2>;
    explicit MountainCarEnv(utils::RenderMode m = utils::RenderMode::None) : Base(m)
    {
        double lo[2], hi[2];
        check(gymrs_observation_space(handle_, lo, hi));
        observation_space = {{lo[0], lo[1]}, {hi[0], hi[1]}}; // :353-354
    }
    spaces::Discrete action_space{3}; // :363
    spaces::BoxR<MountainCarObservation> observation_space;
    static MountainCarObservation make_obs(const float *v) { return {v[0], v[1]}; }
    static constexpr bool WARNS_AFTER_TERMINATION = false; // reward is -1 whatever happened before, mountain_car.rs:423
    static std::optional<core::Unit> step_info() { return std::nullopt; } // info: None, :433
    static std::string invalid_action_message(size_t a) { return std::to_string(a) + " (usize) invalid"; } // :404
};
} // namespace mountain_car
} // namespace envs::classical_control

// N env instances per handle: the shape the GPU path is built for.  ActionReward carries a scalar
// reward / done (core.rs:94-106), so the batched surface hands out struct-of-arrays host vectors
// instead (observation[k * num_envs + i] is field k of env i).  Same role as rust/src/batched.rs.
namespace batched {
enum class Kind { CartPole = GYMRS_CARTPOLE, MountainCar = GYMRS_MOUNTAIN_CAR, Pendulum = GYMRS_PENDULUM };

struct BatchStep {
    const std::vector<float> &observation;
    const std::vector<float> &reward;
    const std::vector<uint8_t> &done;
    const std::vector<uint8_t> &truncated;
};

class BatchedEnv {
  public:
    // global_env_offset keys the reset RNG, so shards of one logical batch on several GPUs give the
    // same per-env results as a single handle
    BatchedEnv(Kind kind, size_t num_envs, int device = 0, uint64_t global_env_offset = 0, bool time_limit = false)
        : kind_(kind), n_(num_envs), global_off_(global_env_offset)
    {
        check(gymrs_create((int)kind, num_envs, device, global_env_offset, nullptr,
                           time_limit ? GYMRS_FLAG_TIME_LIMIT : 0u, &handle_));
        gymrs_buffers b;
        check(gymrs_get_buffers(handle_, &b));
        obs_dim_ = b.obs_dim;
        obs_.resize(obs_dim_ * n_);
        reward_.resize(n_);
        done_.resize(n_);
        truncated_.resize(n_);
    }
    BatchedEnv(const BatchedEnv &) = delete;
    BatchedEnv &operator=(const BatchedEnv &) = delete;
    ~BatchedEnv()
    {
        if (handle_) gymrs_destroy(handle_);
    }

    uint64_t reset(std::optional<uint64_t> seed)
    {
        uint64_t s = seed ? *seed : 0, used = 0;
        check(gymrs_reset(handle_, seed ? &s : nullptr, nullptr, nullptr, nullptr, &used));
        return used;
    }

    // Discrete envs.  Panics like the reference on an action outside the action space.
    BatchStep step(const std::vector<size_t> &actions, bool autoreset)
    {
        if (actions.size() != n_) throw Panic("one action per env instance");
        actions_.resize(n_);
        for (size_t i = 0; i < n_; ++i) actions_[i] = actions[i] > 0x7fffffffu ? 0x7fffffff : (int32_t)actions[i];
        return run(actions_.data(), autoreset, [&](uint64_t bad) {
            const size_t a = actions[(bad - global_off_) % n_]; // gymrs_sync reports the GLOBAL env id
            return std::to_string(a) + (kind_ == Kind::MountainCar ? " (usize) invalid" : " usize invalid");
        });
    }
    // Pendulum: continuous torque per env
    BatchStep step(const std::vector<float> &actions, bool autoreset)
    {
        if (actions.size() != n_) throw Panic("one action per env instance");
        return run(actions.data(), autoreset, [](uint64_t) { return std::string("invalid action"); });
    }

    // the whole handle as one blob; restore() resumes bit-identically (checkpoint / resume)
    std::vector<unsigned char> checkpoint()
    {
        size_t bytes = 0;
        check(gymrs_checkpoint_size(handle_, &bytes));
        std::vector<unsigned char> blob(bytes);
        check(gymrs_checkpoint_save(handle_, blob.data(), bytes));
        return blob;
    }
    void restore(const std::vector<unsigned char> &blob) { check(gymrs_checkpoint_load(handle_, blob.data(), blob.size())); }

    size_t num_envs() const { return n_; }
    size_t obs_dim() const { return obs_dim_; }
    gymrs_env *handle() { return handle_; } // for callers that keep actions / results on the device

  private:
    template <class Msg> BatchStep run(const void *actions, bool autoreset, Msg message)
    {
        check(gymrs_step_host(handle_, actions, autoreset ? GYMRS_STEP_AUTORESET : 0u, obs_.data(), reward_.data(),
                              done_.data(), truncated_.data()));
        uint64_t bad = 0;
        const int rc = gymrs_sync(handle_, &bad);
        if (rc == GYMRS_ERR_INVALID_ACTION) throw Panic(message(bad)); // assert!, cartpole.rs:402-406
        check(rc);
        return {obs_, reward_, done_, truncated_};
    }
    Kind kind_;
    size_t n_, obs_dim_ = 0;
    uint64_t global_off_ = 0;
    gymrs_env *handle_ = nullptr;
    std::vector<int32_t> actions_;
    std::vector<float> obs_, reward_;
    std::vector<uint8_t> done_, truncated_;
};
} // namespace batched

} // namespace gym_rs
