// kernels.hpp -- host-visible launch interface of kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "envs.cuh"

namespace gymrs {

// Everything one step / rollout launch needs besides the env parameter block.
// All pointers are device pointers.  Rows of state / obs are `ld` elements apart.
struct BatchArgs {
    float *state;        // [SD][ld]
    float *obs;          // [OD][ld] (Pendulum) or nullptr when OBS_IS_STATE
    uint64_t ld;
    const void *actions; // [n] (step) or [n_steps][act_ld] (rollout)
    float *reward;       // [n]
    uint8_t *done;       // [n]
    uint8_t *truncated;  // [n], written only with TIME_LIMIT
    int32_t *sbt;        // [n] CartPole steps_beyond_terminated (-1 = None) or nullptr
    uint32_t *elapsed;   // [n] TIME_LIMIT step counters or nullptr
    uint32_t max_steps;  // TIME_LIMIT horizon
    uint64_t n;          // envs in this launch
    uint64_t global_off; // global id of env 0
    uint64_t seed;       // Philox key
    uint64_t epoch;      // Philox counter high half for auto-resets in this step (rollout: first step)
    uint32_t *err;       // device-visible words: [0] sticky flag, [1..2] one offending global id
    int early_actions;   // step: read the action row before griddepcontrol.wait (LaunchOpts::pdl == 2)
    // rollout only
    uint32_t n_steps;
    uint64_t act_ld;     // row stride of actions
    float *obs_out;      // [n_steps][OD][out_ld] or nullptr
    float *reward_out;   // [n_steps][out_ld] or nullptr
    uint8_t *done_out;   // [n_steps][out_ld] or nullptr
    uint64_t out_ld;
};

struct LaunchOpts {
    bool autoreset;
    bool use_sbt;     // CartPole: read/update steps_beyond_terminated
    bool time_limit;
    int pdl;          // 0 off; 1 programmatic dependent launch; 2 = 1 + actions read before the dependency wait
    int vec;          // envs per thread: 1, 2 or 4 (0 = pick)
    int block;        // threads per CTA (0 = default)
};

template <class E>
cudaError_t launch_step(const typename E::P &p, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s);

template <class E>
cudaError_t launch_rollout(const typename E::P &p, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s);

// reset of the envs selected by mask (nullptr = all) at epoch 0
template <class E>
cudaError_t launch_reset(const typename E::P &p, const BatchArgs &a, const uint8_t *mask, cudaStream_t s);

// recompute the Pendulum observation rows from the state rows (after set_state)
cudaError_t launch_pendulum_obs(const BatchArgs &a, cudaStream_t s);

} // namespace gymrs
