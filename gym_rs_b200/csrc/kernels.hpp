// kernels.hpp -- host-visible launch interface of kernels.cu.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
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
    PhiloxKeys rk;       // Philox round keys of the handle's seed (philox_round_keys)
    uint64_t epoch;      // Philox counter high half for auto-resets in this step (rollout: first step)
    // Device copies of the handle's step counter: epoch_dev[2] = seed of the last full reset,
    // epoch_dev[4 + b] = steps since that reset as seen by CTA b (epoch_slots copies, one per CTA
    // of the finest launch geometry).  epoch_from_dev == 0: the kernel uses `epoch` (counted on the
    // host) and never touches them; == 1: each CTA reads and advances its own copy after the
    // dependency wait -- the form a CUDA graph needs, where kernel arguments are frozen at capture
    // but every replay is a new step.
    uint64_t *epoch_dev;
    // epoch_dev[3] = the handle's parameter generation: bumped whenever something a captured launch
    // froze on the host changes afterwards (gymrs_set_params, the first non-auto-reset step that
    // makes steps_beyond_terminated matter).  A device-counted launch compares it with the value it
    // was recorded under and raises err[6] on a mismatch instead of silently using stale constants.
    uint64_t generation;
    uint64_t epoch_slots;
    int epoch_from_dev;
    uint32_t *err;       // device-visible words: [0] invalid-action flag, [1..2] one offending global id,
                         // [3] chained-dependency timeout flag, [4] the offending action's bits,
                         // [5] a graph captured under another seed was replayed
                         // [6] a graph captured under older parameters / step options was replayed
    int early_actions;   // step: read the action row before any dependency is resolved (LaunchOpts::pdl == 2)
    // chained launches: per-CTA progress flags of this handle (see kernels_impl.cuh)
    uint32_t *chain_flags; // [number of CTAs]; CTA b stores chain_seq here when its stores are done
    uint32_t chain_seq;    // sequence number of this step (previous step of the handle = chain_seq - 1)
    int chain;             // 1: wait on chain_flags[blockIdx.x] instead of the whole previous grid
    int publish;           // 1: store chain_seq to chain_flags[blockIdx.x] at the end (pdl == 2 only)
    int l2_prefetch;       // step_kernel: bulk-prefetch the CTA's input rows into L2 before the dependency wait
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
    int vec;          // envs per thread: 1, 2 or 4 (0 = pick); 8 = persistent TMA-staged kernel
    int block;        // threads per CTA (0 = default)
    bool wide;        // step: the high-occupancy build (E::WIDE_MIN_CTAS), gymrs_set_launch_occupancy
};

// ---- launch geometry (shared by the launchers and the host library, which needs to know
// whether two consecutive steps of a handle use the same CTA -> env mapping) ----
inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// widest V the pointers of this launch allow
inline int pick_vec(const BatchArgs &a, int want, bool rollout)
{
    int v = (want == 1 || want == 2 || want == 4) ? want : 4;
    while (v > 1) {
        const size_t fa = 4 * v;
        bool ok = aligned(a.state, fa) && (a.ld % v) == 0 && aligned(a.actions, fa) &&
                  aligned(a.reward, fa) && aligned(a.done, v) &&
                  (!a.obs || aligned(a.obs, fa)) && (!a.sbt || aligned(a.sbt, fa)) &&
                  (!a.elapsed || (aligned(a.elapsed, fa) && aligned(a.truncated, v)));
        if (rollout) {
            ok = ok && (a.act_ld % v) == 0 && (a.out_ld % v) == 0 &&
                 (!a.obs_out || aligned(a.obs_out, fa)) && (!a.reward_out || aligned(a.reward_out, fa)) &&
                 (!a.done_out || aligned(a.done_out, v));
        }
        if (ok) break;
        v >>= 1;
    }
    return v;
}

inline int pick_block(const LaunchOpts &o)
{
    return (o.block >= 32 && o.block <= 256 && o.block % 32 == 0) ? o.block : 256;
}

// The persistent TMA-staged step kernel (step_stream_kernel) is opt-in: vec = 8 in the launch
// config.  It needs every row to allow 128-bit access and a whole number of 4-env groups;
// otherwise the launch silently uses step_kernel.
inline bool use_tma_step(const BatchArgs &a, const LaunchOpts &o)
{
    return o.vec == 8 && (o.block == 0 || o.block == 256) && (a.n % 4) == 0 && pick_vec(a, 0, false) == 4;
}

// L2 prefetch ahead of the dependency wait: on for pdl = 1 (the default), where a launch must wait
// for the whole previous grid and the prefetch is what overlaps its HBM reads with that grid's
// tail (measured, one stream, cold batches: CartPole 9.4 -> 8.7 us/step, MountainCar 6.7 -> 5.9).
// Off for pdl = 2: chained launches already overlap, and on an L2-resident batch the extra L2
// lookups cost ~10 %.  GYMRS_L2_PREFETCH=0 turns it off everywhere (A/B measurements).
inline bool use_l2_prefetch(const LaunchOpts &o)
{
    static const bool disabled = [] {
        const char *e = std::getenv("GYMRS_L2_PREFETCH");
        return e && std::string(e) == "0";
    }();
    return o.pdl == 1 && !disabled;
}

// CTAs of a device-counted launch (they use the 128-bit path or the scalar one, dispatch_vec)
inline uint64_t device_counted_ctas(const BatchArgs &a, const LaunchOpts &o, bool rollout)
{
    const uint64_t v = pick_vec(a, o.vec, rollout) == 4 ? 4 : 1;
    const uint64_t per_cta = v * (uint64_t)pick_block(o);
    return (a.n + per_cta - 1) / per_cta;
}

// env instances covered by one chained-launch progress flag for this launch
inline uint32_t flag_envs(const BatchArgs &a, const LaunchOpts &o)
{
    if (use_tma_step(a, o)) return 1024u;
    return (uint32_t)pick_vec(a, o.vec, false) * (uint32_t)pick_block(o);
}

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
