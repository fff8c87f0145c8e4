// capi.cu -- the extern "C" boundary (include/gymrs_b200.h) over kernels.cu.
//
// One gymrs_env owns: the f32 SoA state of num_envs independent env instances in HBM,
// the per-step result arrays (reward / done / truncated), a CUDA stream, a pinned error
// word the kernels raise on an invalid action, and the f64 parameter block the
// reference keeps as `pub` fields.  There is deliberately no CPU path in this file:
// without a device every entry point that needs one fails with GYMRS_ERR_NO_DEVICE.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gymrs_b200.h"
#include "kernels.hpp"

using namespace gymrs;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(GYMRS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(expr)                                                        \
    do {                                                                \
        cudaError_t e_ = (expr);                                        \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr);             \
    } while (0)

// Makes `device` current for the rest of the enclosing scope and restores the caller's device on
// exit: a library must not change the calling thread's current device behind its back.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1; // already current: nothing to restore
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(dev)                                                  \
    DeviceGuard device_guard_(dev);                                     \
    if (device_guard_.err != cudaSuccess) return cuda_fail(device_guard_.err, "cudaSetDevice")

// largest float <= v  /  smallest float >= v: lets a float compare against a double
// threshold be decided exactly ((x > T) == (x > f32_floor(T)) for every float x).
float f32_floor(double v)
{
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
float f32_ceil(double v)
{
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

void make_reset_box(ResetBox &rb, int dim, const float *low, const float *high)
{
    for (int i = 0; i < 4; ++i) {
        rb.low[i] = 0.f; rb.scale24[i] = 0.f; rb.cap[i] = 0.f;
    }
    for (int i = 0; i < dim; ++i) {
        rb.low[i] = low[i];
        rb.scale24[i] = (high[i] - low[i]) * 5.9604644775390625e-08f; // * 2^-24, exact
        rb.cap[i] = high[i] > low[i] ? std::nextafterf(high[i], low[i]) : low[i];
    }
}

} // namespace

// event slots of the host-buffer pipeline
struct HostEv {
    static constexpr int CHUNKS = 8;
    static constexpr int in_ready(int par, int c) { return par * CHUNKS + c; }
    static constexpr int out_ready(int par, int c) { return 2 * CHUNKS + par * CHUNKS + c; }
    static constexpr int copied(int c) { return 4 * CHUNKS + c; }
    static constexpr int host_done(int par) { return 5 * CHUNKS + par; }
    static constexpr int FENCE = 5 * CHUNKS + 2;
    static constexpr int COUNT = 5 * CHUNKS + 3;
};

struct gymrs_env {
    int kind = 0;
    uint64_t n = 0, ld = 0, global_off = 0;
    int device = 0;
    uint32_t flags = 0;
    uint32_t state_dim = 0, obs_dim = 0;

    gymrs_cartpole_params cp{};
    gymrs_mountain_car_params mc{};
    gymrs_pendulum_params pd{};
    CartPoleP dcp{};
    MountainCarP dmc{};
    PendulumP dpd{};
    float reset_low[4] = {0, 0, 0, 0}, reset_high[4] = {0, 0, 0, 0};

    float *state = nullptr, *obs = nullptr, *reward = nullptr;
    uint8_t *done = nullptr, *truncated = nullptr;
    int32_t *sbt = nullptr;
    uint32_t *elapsed = nullptr;
    void *d_actions = nullptr; // staging for *_host entry points
    uint8_t *d_actions8 = nullptr;  // GYMRS_HOST_U8_ACTIONS: the uint8 form as it arrives, two buffers
    uint8_t *d_done_bits = nullptr, *d_trunc_bits = nullptr; // GYMRS_HOST_PACKED_DONE: bit rows, ld / 8 bytes each
    uint32_t *err_host = nullptr, *err_dev = nullptr;
    // chained launches (kernels_impl.cuh): word [0] = "protocol broken", words [1..] = per-CTA flags
    uint32_t *chain_mem = nullptr;
    uint32_t chain_seq = 0;         // sequence number of the last step launched on this handle
    bool chain_ok = false;          // flags describe the handle's current state for (chain_vec, chain_block)
    int chain_flag_envs = 0;        // env instances per flag of the launch that wrote the flags
    cudaStream_t chain_stream = nullptr;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_streams[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> hev; // events of the host-buffer pipeline (HostEv), created on first use
    uint64_t host_seq = 0;        // tickets handed out by gymrs_step_host_async
    bool host_inflight = false;   // a host step may still be copying results out

    uint64_t seed = 0;       // Philox key of the auto-reset stream
    uint64_t step_count = 0; // steps since the last full reset; epoch of an auto-reset = step_count + 1
    // Device copies of step_count, see BatchArgs::epoch_dev.  Once a step of this handle has been
    // captured into a CUDA graph the host can no longer count steps (replays happen behind its
    // back): from then on the device copies are the authority (device_counted), every CTA of a
    // launch reads and advances its own, and step_count is refreshed from copy 0 when needed.
    // A launch of G CTAs advances copies [0, G) only.  Invariant between launches: the copies of
    // the handle's canonical step geometry (canonical_ctas) are current; a launch with any other
    // grid refreshes all copies from copy 0 before and after itself (spread_step_count), so the
    // invariant holds whatever order graphs are replayed in.
    uint64_t *epoch_mem = nullptr;
    uint64_t epoch_slots = 0;
    bool device_counted = false;
    uint64_t generation = 0; // see BatchArgs::generation
    cudaEvent_t switch_ev = nullptr; // orders a new stream after the old one (gymrs_set_stream)
    cudaEvent_t join_ev = nullptr;   // joins this handle's stream into the first handle's (gymrs_step_pass), created on first use
    bool sbt_dirty = false;  // some env may hold steps_beyond_terminated = Some(_)
    int vec = 0, block = 0, pdl = 1;
    bool wide = false;              // gymrs_set_launch_occupancy
};

namespace {

void default_reset_bounds(int kind, float *low, float *high)
{
    if (kind == GYMRS_CARTPOLE) { // cartpole.rs:353-361
        for (int i = 0; i < 4; ++i) { low[i] = -0.05f; high[i] = 0.05f; }
    } else if (kind == GYMRS_MOUNTAIN_CAR) { // mountain_car.rs:176-187
        low[0] = -0.6f; high[0] = -0.4f; low[1] = 0.f; high[1] = 0.f;
    } else {
        low[0] = -3.14159274f; high[0] = 3.14159274f; low[1] = -1.f; high[1] = 1.f;
    }
}

// Fold the f64 `pub` fields into the f32 block the kernels take by value.
void fold_params(gymrs_env *e)
{
    if (e->kind == GYMRS_CARTPOLE) {
        const gymrs_cartpole_params &p = e->cp;
        const double total_mass = p.masspole + p.masscart;    // cartpole.rs:146-148
        const double polemass_length = p.masspole + p.length; // cartpole.rs:150-152 (a SUM, sic)
        CartPoleP &d = e->dcp;
        d.tau = (float)p.tau;
        d.force_over_m = (float)(p.force_mag / total_mass);
        d.pml_over_m = (float)(polemass_length / total_mass);
        d.gravity = (float)p.gravity;
        d.den_a = (float)(p.length * (4.0 / 3.0));
        d.neg_den_b = (float)(-(p.length * p.masspole / total_mass));
        d.x_thr = f32_floor(p.x_threshold);
        d.th_thr = f32_floor(p.theta_threshold_radians);
        d.semi_implicit = p.kinematics_integrator != 0;
        make_reset_box(d.rb, 4, e->reset_low, e->reset_high);
    } else if (e->kind == GYMRS_MOUNTAIN_CAR) {
        const gymrs_mountain_car_params &p = e->mc;
        MountainCarP &d = e->dmc;
        d.force = (float)p.force;
        d.neg_gravity = (float)(-p.gravity);
        d.max_speed = (float)p.max_speed;
        d.min_position = (float)p.min_position;
        d.max_position = (float)p.max_position;
        d.goal_position = f32_ceil(p.goal_position);
        d.goal_velocity = f32_ceil(p.goal_velocity);
        make_reset_box(d.rb, 1, e->reset_low, e->reset_high);
    } else {
        const gymrs_pendulum_params &p = e->pd;
        PendulumP &d = e->dpd;
        d.max_speed = (float)p.max_speed;
        d.max_torque = (float)p.max_torque;
        d.dt = (float)p.dt;
        d.c_sin = (float)(3.0 * p.g / (2.0 * p.l));
        d.c_u = (float)(3.0 / (p.m * p.l * p.l));
        make_reset_box(d.rb, 2, e->reset_low, e->reset_high);
    }
}

uint32_t max_episode_steps(const gymrs_env *e)
{
    int32_t m = e->kind == GYMRS_CARTPOLE ? e->cp.max_episode_steps
              : e->kind == GYMRS_MOUNTAIN_CAR ? e->mc.max_episode_steps : e->pd.max_episode_steps;
    return m > 0 ? (uint32_t)m : 0xFFFFFFFFu;
}

BatchArgs base_args(const gymrs_env *e)
{
    BatchArgs a = {};
    a.state = e->state;
    a.obs = (e->obs == e->state) ? nullptr : e->obs;
    a.ld = e->ld;
    a.reward = e->reward;
    a.done = e->done;
    a.truncated = e->truncated;
    a.sbt = e->sbt;
    a.elapsed = e->elapsed;
    a.max_steps = max_episode_steps(e);
    a.n = e->n;
    a.global_off = e->global_off;
    a.rk = philox_round_keys(e->seed);
    a.epoch = e->step_count + 1;
    a.epoch_dev = e->epoch_mem;
    a.generation = e->generation;
    a.epoch_slots = e->epoch_slots;
    a.epoch_from_dev = e->device_counted ? 1 : 0;
    a.err = e->err_dev;
    a.chain_flags = e->chain_mem + 1;
    a.chain_seq = e->chain_seq;
    a.chain = 0;
    return a;
}

// a view of envs [begin, begin + count) of the handle
BatchArgs slice_args(const gymrs_env *e, uint64_t begin, uint64_t count)
{
    BatchArgs a = base_args(e);
    a.state += begin;
    if (a.obs) a.obs += begin;
    a.reward += begin;
    a.done += begin;
    a.truncated += begin;
    if (a.sbt) a.sbt += begin;
    if (a.elapsed) a.elapsed += begin;
    a.n = count;
    a.global_off += begin;
    return a;
}

int default_step_block(int kind) { return kind == GYMRS_PENDULUM ? 256 : 128; }

LaunchOpts make_opts(const gymrs_env *e, uint32_t step_flags)
{
    LaunchOpts o = {};
    o.autoreset = (step_flags & GYMRS_STEP_AUTORESET) != 0;
    o.time_limit = (e->flags & GYMRS_FLAG_TIME_LIMIT) != 0;
    // steps_beyond_terminated only matters once a terminated env has been stepped without a
    // reset; with auto-reset from a clean state every env is None at every step.
    o.use_sbt = e->kind == GYMRS_CARTPOLE && (!o.autoreset || e->sbt_dirty);
    o.pdl = e->pdl;
    o.vec = e->vec;
    o.wide = e->wide;
    // CTA size of a step launch.  128 threads (10 CTAs per SM instead of 5, twice as many progress
    // flags) measured 1 % faster than 256 on the two-stream ring and 3-8 % faster for chained and
    // L2-resident launches for CartPole and MountainCar; Pendulum is 1 % faster at 256
    // (profiles/r01_sweeps.md).  The persistent kernel (vec = 8) has its own fixed geometry.
    o.block = e->block;
    if (o.block == 0 && o.vec != 8) o.block = default_step_block(e->kind);
    return o;
}

cudaError_t do_step(gymrs_env *e, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s, bool rollout)
{
    switch (e->kind) {
    case GYMRS_CARTPOLE:
        return rollout ? launch_rollout<CartPole>(e->dcp, a, o, s) : launch_step<CartPole>(e->dcp, a, o, s);
    case GYMRS_MOUNTAIN_CAR:
        return rollout ? launch_rollout<MountainCar>(e->dmc, a, o, s) : launch_step<MountainCar>(e->dmc, a, o, s);
    default:
        return rollout ? launch_rollout<Pendulum>(e->dpd, a, o, s) : launch_step<Pendulum>(e->dpd, a, o, s);
    }
}

cudaError_t do_reset(gymrs_env *e, const BatchArgs &a, const uint8_t *mask, cudaStream_t s)
{
    switch (e->kind) {
    case GYMRS_CARTPOLE: return launch_reset<CartPole>(e->dcp, a, mask, s);
    case GYMRS_MOUNTAIN_CAR: return launch_reset<MountainCar>(e->dmc, a, mask, s);
    default: return launch_reset<Pendulum>(e->dpd, a, mask, s);
    }
}

uint64_t entropy64()
{
    std::random_device rd; // seeding.rs:22 thread_rng().gen()
    return ((uint64_t)rd() << 32) ^ (uint64_t)rd();
}

// A host step submitted with gymrs_step_host_async may still be copying results out of the
// handle's arrays; every other entry point that touches them waits for that first.
int drain_host(gymrs_env *e)
{
    if (!e->host_inflight) return GYMRS_OK;
    CU(cudaStreamSynchronize(e->copy_streams[0]));
    CU(cudaStreamSynchronize(e->copy_streams[1]));
    e->host_inflight = false;
    return GYMRS_OK;
}

// every per-CTA copy of the step count := count (hand-over at first capture, host steps,
// checkpoint load) ...
__global__ void fill_step_count_kernel(uint64_t *epoch_mem, uint64_t slots, uint64_t count)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slots) epoch_mem[4 + i] = count;
}
// ... or := copy 0, which every launch advances (CTA 0 always exists)
__global__ void spread_step_count_kernel(uint64_t *epoch_mem, uint64_t slots)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t count = epoch_mem[4];
    if (i >= 1 && i < slots) epoch_mem[4 + i] = count;
}

cudaError_t fill_step_count(gymrs_env *e, uint64_t count, cudaStream_t s)
{
    fill_step_count_kernel<<<(unsigned)((e->epoch_slots + 255) / 256), 256, 0, s>>>(e->epoch_mem, e->epoch_slots, count);
    return cudaGetLastError();
}

cudaError_t spread_step_count(gymrs_env *e, cudaStream_t s)
{
    spread_step_count_kernel<<<(unsigned)((e->epoch_slots + 255) / 256), 256, 0, s>>>(e->epoch_mem, e->epoch_slots);
    return cudaGetLastError();
}

// CTAs of a step launch in the library's default geometry (4 envs per thread): the copies a
// device-counted launch may rely on without refreshing them
uint64_t canonical_ctas(const gymrs_env *e)
{
    const uint64_t per_cta = 4ull * (uint64_t)default_step_block(e->kind);
    return (e->n + per_cta - 1) / per_cta;
}

// Is the handle's stream being captured into a CUDA graph?  (The legacy default stream cannot be.)
bool capturing(const gymrs_env *e)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(e->stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st == cudaStreamCaptureStatusActive;
}

// A step is about to be recorded into a graph: from now on the device counts this handle's steps.
// Entry points that cannot be captured (they synchronise or use several streams) refuse instead.
int refuse_in_capture(const gymrs_env *e, const char *what)
{
    if (!capturing(e)) return GYMRS_OK;
    return fail(GYMRS_ERR_UNSUPPORTED, std::string(what) + " cannot be captured into a CUDA graph "
                "(only gymrs_step, gymrs_rollout and a seeded full gymrs_reset can)");
}

// The first time a step of this handle is recorded into a graph, hand the step count over to the
// device.  The handle's stream is capturing, so the hand-over runs on a side stream: nothing
// queued earlier on the handle's stream touches the device counter (host-counted launches ignore
// it), and every replay is launched after this function has returned.  Synchronising calls are
// off limits while any stream captures in the default (global) capture mode, hence the relaxed
// mode for the duration -- the same thing torch's allocator does around its event queries.
int begin_device_counting(gymrs_env *e)
{
    if (e->device_counted) return GYMRS_OK;
    if (e->host_inflight) return fail(GYMRS_ERR_UNSUPPORTED, "a host step is in flight: gymrs_host_wait before capturing");
    uint64_t *stage = reinterpret_cast<uint64_t *>(e->err_host + 8);
    stage[0] = e->seed;
    stage[1] = e->generation;
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    CU(cudaThreadExchangeStreamCaptureMode(&mode));
    cudaError_t ce = cudaMemcpyAsync(e->epoch_mem + 2, stage, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, e->copy_streams[0]);
    if (ce == cudaSuccess) ce = fill_step_count(e, e->step_count, e->copy_streams[0]);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->copy_streams[0]);
    cudaThreadExchangeStreamCaptureMode(&mode);
    if (ce != cudaSuccess) return cuda_fail(ce, "handing the step counter over to the device");
    e->device_counted = true;
    return GYMRS_OK;
}

// Bring the host's step_count up to date with the device copy (synchronises the stream).
int refresh_step_count(gymrs_env *e)
{
    if (!e->device_counted) return GYMRS_OK;
    CU(cudaMemcpyAsync(&e->step_count, e->epoch_mem + 4, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return GYMRS_OK;
}

// Something a captured launch froze on the host has changed (see BatchArgs::generation): graphs
// recorded earlier now raise a sticky error when replayed.  The new value is raised on the device
// by a one-thread kernel on the handle's stream, so it is ordered like the call that caused it and,
// inside a capture, becomes part of the recording (max: replaying an old recording never lowers it).
__global__ void raise_generation_kernel(uint64_t *epoch_mem, uint64_t generation)
{
    if (epoch_mem[3] < generation) epoch_mem[3] = generation;
}

cudaError_t bump_generation(gymrs_env *e)
{
    e->generation += 1;
    if (!e->device_counted) return cudaSuccess; // nothing recorded yet: the hand-over writes the value
    raise_generation_kernel<<<1, 1, 0, e->stream>>>(e->epoch_mem, e->generation);
    return cudaGetLastError();
}

// Before the launch arguments of a step are built: the first CartPole step without auto-reset
// makes steps_beyond_terminated matter, so auto-reset steps must clear it from now on (the SBT
// kernel variant); a graph that recorded the cheaper variant is stale.
cudaError_t before_step(gymrs_env *e, uint32_t step_flags)
{
    if (e->kind == GYMRS_CARTPOLE && !(step_flags & GYMRS_STEP_AUTORESET) && !e->sbt_dirty) {
        e->sbt_dirty = true;
        return bump_generation(e);
    }
    return cudaSuccess;
}

void after_step(gymrs_env *e, uint32_t n_steps) { e->step_count += n_steps; }

int free_env(gymrs_env *e)
{
    if (!e) return GYMRS_OK;
    DeviceGuard device_guard_(e->device);
    for (auto &s : e->copy_streams) if (s) cudaStreamSynchronize(s);
    if (e->own_stream) cudaStreamSynchronize(e->own_stream);
    cudaFree(e->state);
    if (e->obs && e->obs != e->state) cudaFree(e->obs);
    cudaFree(e->reward);
    cudaFree(e->done);
    cudaFree(e->truncated);
    cudaFree(e->sbt);
    cudaFree(e->elapsed);
    cudaFree(e->d_actions);
    cudaFree(e->d_actions8);
    cudaFree(e->d_done_bits);
    cudaFree(e->d_trunc_bits);
    cudaFree(e->chain_mem);
    cudaFree(e->epoch_mem);
    if (e->err_host) cudaFreeHost(e->err_host);
    for (auto &s : e->copy_streams) if (s) cudaStreamDestroy(s);
    for (auto &v : e->hev) if (v) cudaEventDestroy(v);
    if (e->switch_ev) cudaEventDestroy(e->switch_ev);
    if (e->join_ev) cudaEventDestroy(e->join_ev);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
    return GYMRS_OK;
}

int alloc_env(gymrs_env *e)
{
    const uint64_t n = e->n;
    e->ld = (n + 127) / 128 * 128; // rows start 512-byte aligned
    if (e->ld == 0) e->ld = 128;
    ON_DEVICE(e->device);
    CU(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    for (auto &s : e->copy_streams) CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&e->switch_ev, cudaEventDisableTiming));
    CU(cudaMalloc(&e->state, sizeof(float) * e->state_dim * e->ld));
    CU(cudaMemsetAsync(e->state, 0, sizeof(float) * e->state_dim * e->ld, e->stream));
    if (e->kind == GYMRS_PENDULUM) {
        CU(cudaMalloc(&e->obs, sizeof(float) * e->obs_dim * e->ld));
        CU(cudaMemsetAsync(e->obs, 0, sizeof(float) * e->obs_dim * e->ld, e->stream));
    } else {
        e->obs = e->state;
    }
    CU(cudaMalloc(&e->reward, sizeof(float) * e->ld));
    CU(cudaMalloc(&e->done, e->ld));
    CU(cudaMalloc(&e->truncated, e->ld));
    CU(cudaMemsetAsync(e->reward, 0, sizeof(float) * e->ld, e->stream));
    CU(cudaMemsetAsync(e->done, 0, e->ld, e->stream));
    CU(cudaMemsetAsync(e->truncated, 0, e->ld, e->stream));
    if (e->kind == GYMRS_CARTPOLE) {
        CU(cudaMalloc(&e->sbt, sizeof(int32_t) * e->ld));
        CU(cudaMemsetAsync(e->sbt, 0xFF, sizeof(int32_t) * e->ld, e->stream)); // -1 = None
    }
    if (e->flags & GYMRS_FLAG_TIME_LIMIT) {
        CU(cudaMalloc(&e->elapsed, sizeof(uint32_t) * e->ld));
        CU(cudaMemsetAsync(e->elapsed, 0, sizeof(uint32_t) * e->ld, e->stream));
    }
    const size_t chain_words = (size_t)((n + 31) / 32) + 2; // V = 1, 32-thread CTAs is the finest geometry
    CU(cudaMalloc(&e->chain_mem, chain_words * sizeof(uint32_t)));
    CU(cudaMemsetAsync(e->chain_mem, 0, chain_words * sizeof(uint32_t), e->stream));
    e->epoch_slots = (n + 31) / 32 + 1; // one copy per CTA of the finest geometry (V = 1, 32 threads)
    CU(cudaMalloc(&e->epoch_mem, (4 + e->epoch_slots) * sizeof(uint64_t)));
    CU(cudaMemsetAsync(e->epoch_mem, 0, (4 + e->epoch_slots) * sizeof(uint64_t), e->stream));
    CU(cudaHostAlloc(&e->err_host, 16 * sizeof(uint32_t), cudaHostAllocMapped)); // [8..13]: staging, begin_device_counting
    std::memset(e->err_host, 0, 16 * sizeof(uint32_t));
    CU(cudaHostGetDevicePointer(&e->err_dev, e->err_host, 0));
    return GYMRS_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------
extern "C" {

int gymrs_abi_version(void) { return GYMRS_ABI_VERSION; }

const char *gymrs_last_error(void) { return g_last_error.c_str(); }

int gymrs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gymrs_default_params(int kind, void *params)
{
    if (!params) return fail(GYMRS_ERR_BAD_ARG, "params is NULL");
    if (kind == GYMRS_CARTPOLE) {
        gymrs_cartpole_params p = {};
        p.gravity = 9.8;                                       // cartpole.rs:94
        p.masscart = 1.0;                                      // :95
        p.masspole = 0.1;                                      // :96
        p.length = 0.5;                                        // :97
        p.force_mag = 10.0;                                    // :98
        p.tau = 0.02;                                          // :99
        p.kinematics_integrator = 0;                           // :100
        p.theta_threshold_radians = 12. * 2. * 3.14159265358979323846 / 360.; // :102
        p.x_threshold = 2.4;                                   // :103
        p.max_episode_steps = 500;                             // doc :50
        *(gymrs_cartpole_params *)params = p;
    } else if (kind == GYMRS_MOUNTAIN_CAR) {
        gymrs_mountain_car_params p = {};
        p.min_position = -1.2;  // mountain_car.rs:344
        p.max_position = 0.6;   // :345
        p.max_speed = 0.07;     // :346
        p.goal_position = 0.5;  // :347
        p.goal_velocity = 0.;   // :348
        p.force = 0.001;        // :350
        p.gravity = 0.0025;     // :351
        p.max_episode_steps = 200; // doc :45
        *(gymrs_mountain_car_params *)params = p;
    } else if (kind == GYMRS_PENDULUM) {
        gymrs_pendulum_params p = {};
        p.max_speed = 8.0; p.max_torque = 2.0; p.dt = 0.05; p.g = 10.0; p.m = 1.0; p.l = 1.0;
        p.max_episode_steps = 200;
        *(gymrs_pendulum_params *)params = p;
    } else {
        return fail(GYMRS_ERR_BAD_ARG, "unknown env kind");
    }
    return GYMRS_OK;
}

int gymrs_create(int kind, uint64_t num_envs, int device, uint64_t global_env_offset,
                 const void *params, uint32_t flags, gymrs_env **out)
{
    if (!out) return fail(GYMRS_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (kind < GYMRS_CARTPOLE || kind > GYMRS_PENDULUM) return fail(GYMRS_ERR_BAD_ARG, "unknown env kind");
    if (num_envs == 0) return fail(GYMRS_ERR_BAD_ARG, "num_envs must be > 0");
    // the kernels index the envs of one launch with 32 bits; larger batches are several handles
    // (global_env_offset keeps their reset streams disjoint)
    if (num_envs > (1ull << 31)) return fail(GYMRS_ERR_BAD_ARG, "num_envs too large: at most 2^31 envs per handle");
    int ndev = gymrs_device_count();
    if (ndev <= 0) return fail(GYMRS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(GYMRS_ERR_BAD_ARG, "device ordinal out of range");

    gymrs_env *e = new (std::nothrow) gymrs_env();
    if (!e) return fail(GYMRS_ERR_ALLOC, "host allocation failed");
    e->kind = kind;
    e->n = num_envs;
    e->device = device;
    e->global_off = global_env_offset;
    e->flags = flags;
    e->state_dim = kind == GYMRS_CARTPOLE ? 4 : 2;
    e->obs_dim = kind == GYMRS_CARTPOLE ? 4 : (kind == GYMRS_MOUNTAIN_CAR ? 2 : 3);
    gymrs_default_params(kind, kind == GYMRS_CARTPOLE ? (void *)&e->cp
                               : kind == GYMRS_MOUNTAIN_CAR ? (void *)&e->mc : (void *)&e->pd);
    if (params) {
        if (kind == GYMRS_CARTPOLE) e->cp = *(const gymrs_cartpole_params *)params;
        else if (kind == GYMRS_MOUNTAIN_CAR) e->mc = *(const gymrs_mountain_car_params *)params;
        else e->pd = *(const gymrs_pendulum_params *)params;
    }
    default_reset_bounds(kind, e->reset_low, e->reset_high);
    fold_params(e);
    int rc = alloc_env(e);
    if (rc != GYMRS_OK) {
        std::string msg = g_last_error;
        free_env(e);
        return fail(rc, msg);
    }
    // ::new samples an initial state from an entropy-seeded RNG (cartpole.rs:92,120)
    rc = gymrs_reset(e, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc == GYMRS_OK && cudaStreamSynchronize(e->stream) != cudaSuccess) rc = fail(GYMRS_ERR_CUDA, "initial reset failed");
    if (rc != GYMRS_OK) {
        std::string msg = g_last_error;
        free_env(e);
        return fail(rc, msg);
    }
    *out = e;
    return GYMRS_OK;
}

int gymrs_destroy(gymrs_env *env) { return free_env(env); }

int gymrs_clone(const gymrs_env *src, gymrs_env **out)
{
    if (!src || !out) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    if (int rc_ = refuse_in_capture(src, "gymrs_clone")) return rc_;
    ON_DEVICE(src->device);
    gymrs_env *e = new (std::nothrow) gymrs_env();
    if (!e) return fail(GYMRS_ERR_ALLOC, "host allocation failed");
    e->kind = src->kind; e->n = src->n; e->device = src->device; e->global_off = src->global_off;
    e->flags = src->flags; e->state_dim = src->state_dim; e->obs_dim = src->obs_dim;
    e->cp = src->cp; e->mc = src->mc; e->pd = src->pd;
    e->dcp = src->dcp; e->dmc = src->dmc; e->dpd = src->dpd;
    std::memcpy(e->reset_low, src->reset_low, sizeof e->reset_low);
    std::memcpy(e->reset_high, src->reset_high, sizeof e->reset_high);
    e->seed = src->seed; e->step_count = src->step_count; e->sbt_dirty = src->sbt_dirty;
    e->vec = src->vec; e->block = src->block; e->pdl = src->pdl; e->wide = src->wide;
    e->device_counted = src->device_counted;
    e->generation = src->generation;
    int rc = alloc_env(e);
    if (rc != GYMRS_OK) {
        std::string msg = g_last_error;
        free_env(e);
        return fail(rc, msg);
    }
    // order the copies after everything already queued on the source's streams
    cudaError_t ce = cudaStreamSynchronize(src->stream);
    for (auto &s : src->copy_streams) if (ce == cudaSuccess && s) ce = cudaStreamSynchronize(s);
    auto cp = [&](void *d, const void *s, size_t b) {
        if (ce == cudaSuccess && d && s) ce = cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, e->stream);
    };
    cp(e->state, src->state, sizeof(float) * e->state_dim * e->ld);
    if (e->obs != e->state) cp(e->obs, src->obs, sizeof(float) * e->obs_dim * e->ld);
    cp(e->reward, src->reward, sizeof(float) * e->ld);
    cp(e->done, src->done, e->ld);
    cp(e->truncated, src->truncated, e->ld);
    cp(e->sbt, src->sbt, sizeof(int32_t) * e->ld);
    cp(e->elapsed, src->elapsed, sizeof(uint32_t) * e->ld);
    cp(e->epoch_mem + 2, src->epoch_mem + 2, 2 * sizeof(uint64_t)); // the seed and the parameter generation
    if (ce == cudaSuccess && src->device_counted) {
        // copy 0 of the source is always current: take the count from it and fill every copy
        ce = cudaMemcpyAsync(&e->step_count, src->epoch_mem + 4, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce == cudaSuccess) ce = fill_step_count(e, e->step_count, e->stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) {
        free_env(e);
        return cuda_fail(ce, "gymrs_clone");
    }
    *out = e;
    return GYMRS_OK;
}

int gymrs_set_params(gymrs_env *e, const void *params)
{
    if (!e || !params) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (e->kind == GYMRS_CARTPOLE) e->cp = *(const gymrs_cartpole_params *)params;
    else if (e->kind == GYMRS_MOUNTAIN_CAR) e->mc = *(const gymrs_mountain_car_params *)params;
    else e->pd = *(const gymrs_pendulum_params *)params;
    fold_params(e);
    ON_DEVICE(e->device);
    CU(bump_generation(e)); // graphs recorded under the old constants report GYMRS_ERR_UNSUPPORTED when replayed
    return GYMRS_OK;
}

int gymrs_get_params(const gymrs_env *e, void *params)
{
    if (!e || !params) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (e->kind == GYMRS_CARTPOLE) *(gymrs_cartpole_params *)params = e->cp;
    else if (e->kind == GYMRS_MOUNTAIN_CAR) *(gymrs_mountain_car_params *)params = e->mc;
    else *(gymrs_pendulum_params *)params = e->pd;
    return GYMRS_OK;
}

int gymrs_set_stream(gymrs_env *e, void *cuda_stream)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    if (next == e->stream) return GYMRS_OK;
    if (int rc_ = refuse_in_capture(e, "gymrs_set_stream")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    // Work already queued on the old stream precedes whatever the handle does on the new one: an
    // event dependency, not a host synchronisation.  If the new stream is recording a CUDA graph,
    // earlier work cannot become a dependency of the graph and no synchronising call is allowed;
    // torch.cuda.graph() synchronises the device before it starts recording, any other caller
    // must have done the same.
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(next, &st) != cudaSuccess) {
        cudaGetLastError();
        st = cudaStreamCaptureStatusNone;
    }
    if (st != cudaStreamCaptureStatusActive) {
        CU(cudaEventRecord(e->switch_ev, e->stream));
        CU(cudaStreamWaitEvent(next, e->switch_ev, 0));
    }
    e->stream = next;
    e->chain_ok = false;
    return GYMRS_OK;
}

int gymrs_get_stream(const gymrs_env *e, void **cuda_stream)
{
    if (!e || !cuda_stream) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *cuda_stream = (void *)e->stream;
    return GYMRS_OK;
}

// Tuning knobs (not part of the reference surface); see the header for the pdl contract.
int gymrs_set_launch_config(gymrs_env *e, int vec, int block, int pdl)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    if (!(vec == 0 || vec == 1 || vec == 2 || vec == 4 || vec == 8)) return fail(GYMRS_ERR_BAD_ARG, "vec must be 0, 1, 2, 4 or 8");
    if (block != 0 && (block < 32 || block > 256 || block % 32)) return fail(GYMRS_ERR_BAD_ARG, "block must be a multiple of 32 in [32, 256]");
    if (pdl < 0 || pdl > 2) return fail(GYMRS_ERR_BAD_ARG, "pdl must be 0, 1 or 2");
    e->vec = vec; e->block = block; e->pdl = pdl;
    return GYMRS_OK;
}

int gymrs_set_launch_occupancy(gymrs_env *e, int wide)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    if (wide != 0 && wide != 1) return fail(GYMRS_ERR_BAD_ARG, "wide must be 0 or 1");
    e->wide = wide != 0;
    return GYMRS_OK;
}

int gymrs_reset(gymrs_env *e, const uint64_t *seed, const float *low, const float *high,
                const uint8_t *mask, uint64_t *seed_used)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    if ((low == nullptr) != (high == nullptr)) return fail(GYMRS_ERR_BAD_ARG, "low and high must be given together");
    ON_DEVICE(e->device);
    const bool cap = capturing(e);
    if (cap && (!seed || mask)) return fail(GYMRS_ERR_UNSUPPORTED, "only a seeded full gymrs_reset can be captured into a CUDA graph");
    if (cap)
        if (int rc_ = begin_device_counting(e)) return rc_;
    if (int rc_ = drain_host(e)) return rc_;
    const uint64_t s = seed ? *seed : entropy64(); // seeding.rs:22
    if (seed_used) *seed_used = s;
    // `options: Option<BoxR<Obs>>` applies to this call only (cartpole.rs:351-365)
    float lo[4], hi[4];
    default_reset_bounds(e->kind, lo, hi);
    if (low) {
        for (uint32_t i = 0; i < e->state_dim; ++i) {
            if (!(low[i] <= high[i])) return fail(GYMRS_ERR_BAD_ARG, "reset bounds need low <= high");
            lo[i] = low[i]; hi[i] = high[i];
        }
    }
    float keep_lo[4], keep_hi[4];
    std::memcpy(keep_lo, e->reset_low, sizeof keep_lo);
    std::memcpy(keep_hi, e->reset_high, sizeof keep_hi);
    std::memcpy(e->reset_low, lo, sizeof lo);
    std::memcpy(e->reset_high, hi, sizeof hi);
    fold_params(e);
    BatchArgs a = base_args(e);
    a.rk = philox_round_keys(s);
    e->chain_ok = false;
    cudaError_t ce = do_reset(e, a, mask, e->stream);
    std::memcpy(e->reset_low, keep_lo, sizeof keep_lo);
    std::memcpy(e->reset_high, keep_hi, sizeof keep_hi);
    fold_params(e);
    if (ce != cudaSuccess) return cuda_fail(ce, "reset launch");
    if (!mask) { // a full reset restarts the handle's auto-reset stream
        e->seed = s; // on a device-counted handle the reset kernel also zeroes the device counters and records the seed
        e->step_count = 0;
        e->sbt_dirty = false;
    }
    return GYMRS_OK;
}

int gymrs_step(gymrs_env *e, const void *actions, uint32_t step_flags)
{
    if (!e || !actions) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    ON_DEVICE(e->device);
    if (capturing(e))
        if (int rc_ = begin_device_counting(e)) return rc_;
    if (int rc_ = drain_host(e)) return rc_;
    CU(before_step(e, step_flags));
    BatchArgs a = base_args(e);
    a.actions = actions;
    LaunchOpts o = make_opts(e, step_flags);
    // device-counted steps read the counter after a grid-wide dependency: no chained launches
    if (e->device_counted && o.pdl == 2) o.pdl = 1;
    const bool odd_grid = e->device_counted && device_counted_ctas(a, o, false) != canonical_ctas(e);
    if (odd_grid) CU(spread_step_count(e, e->stream));
    // Chained launch: this step may skip the grid-wide dependency on the previous launch when the
    // handle's per-CTA flags describe its current state for exactly this CTA -> env mapping.
    const int fe = (int)flag_envs(a, o);
    a.chain_seq = ++e->chain_seq;
    a.chain = (o.pdl == 2 && e->chain_ok && e->chain_flag_envs == fe && e->chain_stream == e->stream) ? 1 : 0;
    a.publish = (o.pdl == 2) ? 1 : 0;
    e->chain_ok = false;
    CU(do_step(e, a, o, e->stream, false));
    if (odd_grid) CU(spread_step_count(e, e->stream));
    e->chain_ok = a.publish != 0; // a pdl == 2 step publishes its flags, chained or not
    e->chain_flag_envs = fe;
    e->chain_stream = e->stream;
    after_step(e, 1);
    return GYMRS_OK;
}

// Host-buffer step, asynchronous.  Chunked pipeline over three streams:
//   h2d:  actions chunk c            -> staging[ticket % 2]
//   cs :  step kernel on chunk c     (after its actions arrived AND the previous host step's
//                                     copy-out of chunk c finished: the result arrays are reused)
//   d2h:  observation / reward / done of chunk c -> caller's buffers
// Chunk c + 1's copy-in and chunk c - 1's copy-out overlap chunk c's kernel, and because nothing
// here blocks the host, step t + 1 can be submitted while step t's results are still streaming
// out: the D2H engine (the bottleneck at 21 B per env-step vs 4 B in) never idles.  Chunk
// boundaries are multiples of 1024 envs so every chunk keeps the 128-bit access path.
int gymrs_step_many(gymrs_env *const *envs, const void *const *actions, uint32_t count, uint32_t step_flags,
                    uint32_t *done)
{
    if (done) *done = 0;
    if (count && (!envs || !actions)) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    for (uint32_t i = 0; i < count; ++i) {
        if (int rc = gymrs_step(envs[i], actions[i], step_flags)) return rc;
        if (done) *done = i + 1;
    }
    return GYMRS_OK;
}

// A pass over several handles bracketed by two caller-owned events: `begin` is recorded on the first
// handle's stream before anything is launched and every other stream of the pass waits for it, `end`
// is recorded on the first handle's stream once every other stream of the pass has been joined into
// it.  All in one FFI crossing, so nothing but the launches themselves sits between the two events.
int gymrs_step_pass(gymrs_env *const *envs, const void *const *actions, uint32_t count, uint32_t step_flags,
                    void *begin_event, void *end_event, uint32_t *done)
{
    if (done) *done = 0;
    if (count && (!envs || !actions)) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (!begin_event && !end_event) return gymrs_step_many(envs, actions, count, step_flags, done);
    if (count == 0) return fail(GYMRS_ERR_BAD_ARG, "a pass with events needs at least one step");
    for (uint32_t i = 0; i < count; ++i) {
        if (!envs[i]) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
        if (envs[i]->device != envs[0]->device) return fail(GYMRS_ERR_BAD_ARG, "the handles of a pass with events must share a device");
    }
    gymrs_env *first = envs[0];
    if (int rc_ = refuse_in_capture(first, "gymrs_step_pass with events")) return rc_;
    ON_DEVICE(first->device);
    // the other streams of the pass, each with the first handle that uses it
    std::vector<gymrs_env *> others;
    for (uint32_t i = 1; i < count; ++i) {
        bool seen = envs[i]->stream == first->stream;
        for (gymrs_env *o : others) seen = seen || o->stream == envs[i]->stream;
        if (!seen) others.push_back(envs[i]);
    }
    if (begin_event) CU(cudaEventRecord((cudaEvent_t)begin_event, first->stream));
    // each other stream is forked from the begin event right before its first launch of the pass, so that
    // the first launch follows the begin record with nothing in between
    std::vector<cudaStream_t> forked;
    int rc = GYMRS_OK;
    for (uint32_t i = 0; i < count && rc == GYMRS_OK; ++i) {
        if (begin_event && envs[i]->stream != first->stream) {
            bool is_forked = false;
            for (cudaStream_t s : forked) is_forked = is_forked || s == envs[i]->stream;
            if (!is_forked) {
                CU(cudaStreamWaitEvent(envs[i]->stream, (cudaEvent_t)begin_event, 0));
                forked.push_back(envs[i]->stream);
            }
        }
        rc = gymrs_step(envs[i], actions[i], step_flags);
        if (rc == GYMRS_OK && done) *done = i + 1;
    }
    if (end_event) { // also after a failed launch: what was enqueued is still bracketed
        for (gymrs_env *o : others) {
            if (!o->join_ev) CU(cudaEventCreateWithFlags(&o->join_ev, cudaEventDisableTiming));
            CU(cudaEventRecord(o->join_ev, o->stream));
            CU(cudaStreamWaitEvent(first->stream, o->join_ev, 0));
        }
        CU(cudaEventRecord((cudaEvent_t)end_event, first->stream));
    }
    return rc;
}

} // extern "C"

namespace {

// compact wire formats of the host path (GYMRS_HOST_U8_ACTIONS / GYMRS_HOST_PACKED_DONE): tiny
// conversion kernels next to the copies, so the step kernel itself keeps one action type
__global__ void widen_u8_kernel(const uint8_t *__restrict__ in, int32_t *__restrict__ out, uint64_t n)
{
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n) {
        const uchar4 v = *reinterpret_cast<const uchar4 *>(in + i);
        *reinterpret_cast<int4 *>(out + i) = make_int4(v.x, v.y, v.z, v.w);
    } else {
        for (uint64_t j = i; j < n; ++j) out[j] = in[j];
    }
}
// flag bytes (0 / 1) -> bits, byte i/8 bit i%8
__global__ void pack_bits_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint64_t n)
{
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    uint32_t bits = 0;
    if (i + 8 <= n) {
        const uint2 w = *reinterpret_cast<const uint2 *>(in + i);
        // four 0/1 bytes -> a nibble: the multiply moves byte k's bit to position 24 + k, no carries
        bits = (((w.x & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) | ((((w.y & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << 4);
    } else {
        for (uint64_t j = i; j < n; ++j) bits |= (in[j] ? 1u : 0u) << (j - i);
    }
    out[i >> 3] = (uint8_t)bits;
}

int host_step_submit(gymrs_env *e, const void *actions, uint32_t step_flags, uint32_t transport,
                     float *obs, float *reward, uint8_t *done, uint8_t *truncated, uint64_t *ticket)
{
    if (!e || !actions) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (int rc_ = refuse_in_capture(e, "gymrs_step_host")) return rc_;
    if (transport & ~(GYMRS_HOST_U8_ACTIONS | GYMRS_HOST_PACKED_DONE)) return fail(GYMRS_ERR_BAD_ARG, "unknown transport flag");
    const bool u8 = (transport & GYMRS_HOST_U8_ACTIONS) != 0, packed = (transport & GYMRS_HOST_PACKED_DONE) != 0;
    if (u8 && e->kind == GYMRS_PENDULUM) return fail(GYMRS_ERR_UNSUPPORTED, "GYMRS_HOST_U8_ACTIONS needs a discrete action space");
    ON_DEVICE(e->device);
    if (u8 && !e->d_actions8) CU(cudaMalloc(&e->d_actions8, 2 * e->ld));
    if (packed && !e->d_done_bits) {
        CU(cudaMalloc(&e->d_done_bits, e->ld / 8));
        CU(cudaMalloc(&e->d_trunc_bits, e->ld / 8));
    }
    if (e->device_counted && !e->host_inflight) {
        // host steps are sliced into several launches that must share one epoch: they are
        // host-counted (from the device's current count), and the new count is written back to
        // every device copy after the slices (fill_step_count below)
        if (int rc_ = refresh_step_count(e)) return rc_;
    }
    const uint64_t n = e->n;
    const uint64_t tk = e->host_seq;
    const int par = (int)(tk & 1);
    if (!e->d_actions) CU(cudaMalloc(&e->d_actions, 2 * 4 * e->ld)); // two staging buffers
    if (e->hev.empty()) {
        e->hev.resize(HostEv::COUNT);
        for (auto &v : e->hev) CU(cudaEventCreateWithFlags(&v, cudaEventDisableTiming));
    }
    e->chain_ok = false; // the slice launches below use their own CTA numbering
    // Chunk count: PCIe copies below ~2 MB lose bandwidth (measured on the B200 box: 512 KB
    // copies reach 42 GB/s, 4 MB copies 54 GB/s), so rows are split into at most a few pieces,
    // each at least 256 K envs (1 MB per f32 row).  GYMRS_HOST_CHUNKS overrides for experiments.
    static const int chunk_override = [] {
        const char *s = std::getenv("GYMRS_HOST_CHUNKS");
        const int v = s ? std::atoi(s) : 0;
        return v >= 1 && v <= HostEv::CHUNKS ? v : 0;
    }();
    uint64_t nchunk = chunk_override ? (uint64_t)chunk_override : 2;
    while (nchunk > 1 && n / nchunk < (1u << 18)) --nchunk;
    const uint64_t per = ((n + nchunk - 1) / nchunk + 1023) / 1024 * 1024;
    cudaStream_t h2d = e->copy_streams[0], d2h = e->copy_streams[1], cs = e->stream;
    char *staging = (char *)e->d_actions + (size_t)par * 4 * e->ld;
    CU(before_step(e, step_flags));
    LaunchOpts o = make_opts(e, step_flags);
    o.pdl = 0;
    if (!e->host_inflight) {
        // everything queued earlier on the compute stream (reset, set_state, device steps)
        // precedes this host step's copies
        CU(cudaEventRecord(e->hev[HostEv::FENCE], cs));
        CU(cudaStreamWaitEvent(h2d, e->hev[HostEv::FENCE], 0));
        CU(cudaStreamWaitEvent(d2h, e->hev[HostEv::FENCE], 0));
    }
    int c = 0;
    for (uint64_t b = 0; b < n; b += per, ++c) {
        const uint64_t cnt = (b + per <= n) ? per : n - b;
        cudaEvent_t in_ready = e->hev[HostEv::in_ready(par, c)], out_ready = e->hev[HostEv::out_ready(par, c)],
                    copied = e->hev[HostEv::copied(c)];
        // staging[par] chunk c was last read by the kernel of host step tk - 2
        if (tk >= 2) CU(cudaStreamWaitEvent(h2d, out_ready, 0));
        if (u8)
            CU(cudaMemcpyAsync(e->d_actions8 + (size_t)par * e->ld + b, (const uint8_t *)actions + b, cnt, cudaMemcpyHostToDevice, h2d));
        else
            CU(cudaMemcpyAsync(staging + 4 * b, (const char *)actions + 4 * b, 4 * cnt, cudaMemcpyHostToDevice, h2d));
        CU(cudaEventRecord(in_ready, h2d));
        CU(cudaStreamWaitEvent(cs, in_ready, 0));
        if (u8) {
            widen_u8_kernel<<<(unsigned)((cnt + 1023) / 1024), 256, 0, cs>>>(
                e->d_actions8 + (size_t)par * e->ld + b, reinterpret_cast<int32_t *>(staging + 4 * b), cnt);
            CU(cudaGetLastError());
        }
        // the result rows of chunk c are still being copied out for host step tk - 1
        if (e->host_inflight) CU(cudaStreamWaitEvent(cs, copied, 0));
        BatchArgs a = slice_args(e, b, cnt);
        a.epoch_from_dev = 0;
        a.actions = staging + 4 * b;
        CU(do_step(e, a, o, cs, false));
        if (packed) { // chunk starts are multiples of 1024 envs, so every chunk owns whole bytes
            const unsigned g = (unsigned)((cnt + 2047) / 2048);
            if (done) pack_bits_kernel<<<g, 256, 0, cs>>>(e->done + b, e->d_done_bits + b / 8, cnt);
            if (truncated) pack_bits_kernel<<<g, 256, 0, cs>>>(e->truncated + b, e->d_trunc_bits + b / 8, cnt);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(out_ready, cs));
        CU(cudaStreamWaitEvent(d2h, out_ready, 0));
        if (obs)
            CU(cudaMemcpy2DAsync(obs + b, sizeof(float) * n, e->obs + b, sizeof(float) * e->ld,
                                 sizeof(float) * cnt, e->obs_dim, cudaMemcpyDeviceToHost, d2h));
        if (reward) CU(cudaMemcpyAsync(reward + b, e->reward + b, sizeof(float) * cnt, cudaMemcpyDeviceToHost, d2h));
        if (packed) {
            const size_t nb = (size_t)((cnt + 7) / 8);
            if (done) CU(cudaMemcpyAsync(done + b / 8, e->d_done_bits + b / 8, nb, cudaMemcpyDeviceToHost, d2h));
            if (truncated) CU(cudaMemcpyAsync(truncated + b / 8, e->d_trunc_bits + b / 8, nb, cudaMemcpyDeviceToHost, d2h));
        } else {
            if (done) CU(cudaMemcpyAsync(done + b, e->done + b, cnt, cudaMemcpyDeviceToHost, d2h));
            if (truncated) CU(cudaMemcpyAsync(truncated + b, e->truncated + b, cnt, cudaMemcpyDeviceToHost, d2h));
        }
        CU(cudaEventRecord(copied, d2h));
    }
    CU(cudaEventRecord(e->hev[HostEv::host_done(par)], d2h));
    after_step(e, 1);
    if (e->device_counted) CU(fill_step_count(e, e->step_count, cs)); // the slices were host-counted
    e->host_inflight = true;
    e->host_seq = tk + 1;
    if (ticket) *ticket = tk;
    return GYMRS_OK;
}

} // namespace

extern "C" {

int gymrs_step_host_async(gymrs_env *e, const void *actions, uint32_t step_flags,
                          float *obs, float *reward, uint8_t *done, uint8_t *truncated, uint64_t *ticket)
{
    return host_step_submit(e, actions, step_flags, 0u, obs, reward, done, truncated, ticket);
}

int gymrs_host_wait(gymrs_env *e, uint64_t ticket)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    if (ticket >= e->host_seq) return fail(GYMRS_ERR_BAD_ARG, "unknown ticket");
    if (e->hev.empty()) return GYMRS_OK;
    ON_DEVICE(e->device);
    // The event slot is shared by tickets of the same parity.  If a later step re-recorded it,
    // that record sits behind this ticket's copies on the same in-order stream, so waiting for
    // it can only be longer, never shorter.
    CU(cudaEventSynchronize(e->hev[HostEv::host_done((int)(ticket & 1))]));
    if (ticket + 1 == e->host_seq) {
        // newest host step: nothing of it is in flight any more
        CU(cudaStreamSynchronize(e->copy_streams[1]));
        e->host_inflight = false;
    }
    return GYMRS_OK;
}

int gymrs_step_host(gymrs_env *e, const void *actions, uint32_t step_flags,
                    float *obs, float *reward, uint8_t *done, uint8_t *truncated)
{
    uint64_t ticket = 0;
    int rc = gymrs_step_host_async(e, actions, step_flags, obs, reward, done, truncated, &ticket);
    if (rc != GYMRS_OK) return rc;
    rc = gymrs_host_wait(e, ticket);
    if (rc != GYMRS_OK) return rc;
    ON_DEVICE(e->device);
    CU(cudaStreamSynchronize(e->stream));
    return GYMRS_OK;
}

// The host loop of examples/cartpole.rs:15-30 for a whole batch, inside the library: step t is
// submitted while step t - 1 is still streaming out, the consumer sees the steps in order.
int gymrs_rollout_host(gymrs_env *e, uint32_t n_steps, uint32_t step_flags, const gymrs_host_rollout_desc *d)
{
    if (!e || !d || !d->actions) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (d->action_slots == 0) return fail(GYMRS_ERR_BAD_ARG, "action_slots must be >= 1");
    if (d->result_slots < 2) return fail(GYMRS_ERR_BAD_ARG, "result_slots must be >= 2 (step t + 1 is submitted while step t streams out)");
    const uint64_t n = e->n;
    const bool u8 = (d->transport & GYMRS_HOST_U8_ACTIONS) != 0, packed = (d->transport & GYMRS_HOST_PACKED_DONE) != 0;
    const size_t act_stride = (size_t)n * (u8 ? 1 : 4), flag_stride = packed ? (size_t)((n + 7) / 8) : (size_t)n;
    uint64_t prev = 0;
    for (uint32_t t = 0; t < n_steps; ++t) {
        const uint32_t slot = t % d->result_slots;
        uint64_t tk = 0;
        int rc = host_step_submit(e, (const char *)d->actions + (size_t)(t % d->action_slots) * act_stride, step_flags, d->transport,
                                  d->obs ? d->obs + (size_t)slot * e->obs_dim * n : nullptr,
                                  d->reward ? d->reward + (size_t)slot * n : nullptr,
                                  d->done ? d->done + (size_t)slot * flag_stride : nullptr,
                                  d->truncated ? d->truncated + (size_t)slot * flag_stride : nullptr, &tk);
        if (rc != GYMRS_OK) return rc;
        if (t > 0) {
            if (int rc_ = gymrs_host_wait(e, prev)) return rc_;
            if (d->on_step) d->on_step(d->user, t - 1, (t - 1) % d->result_slots);
        }
        prev = tk;
    }
    if (n_steps > 0) {
        if (int rc_ = gymrs_host_wait(e, prev)) return rc_;
        if (d->on_step) d->on_step(d->user, n_steps - 1, (n_steps - 1) % d->result_slots);
    }
    ON_DEVICE(e->device);
    CU(cudaStreamSynchronize(e->stream));
    return GYMRS_OK;
}

int gymrs_rollout(gymrs_env *e, const void *actions, uint32_t n_steps, uint32_t step_flags,
                  float *obs_out, float *reward_out, uint8_t *done_out)
{
    if (!e || !actions) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (n_steps == 0) return GYMRS_OK;
    ON_DEVICE(e->device);
    if (capturing(e))
        if (int rc_ = begin_device_counting(e)) return rc_;
    if (int rc_ = drain_host(e)) return rc_;
    CU(before_step(e, step_flags));
    BatchArgs a = base_args(e);
    a.actions = actions;
    a.n_steps = n_steps;
    a.act_ld = e->n;
    a.out_ld = e->n;
    a.obs_out = obs_out;
    a.reward_out = reward_out;
    a.done_out = done_out;
    e->chain_ok = false;
    LaunchOpts o = make_opts(e, step_flags);
    // measured (profiles/r01_sweeps.md): CartPole 209 G env-steps/s at 128 threads vs 200 G at 256;
    // MountainCar 287 vs 305, Pendulum 236 vs 246
    if (e->block == 0) o.block = e->kind == GYMRS_CARTPOLE ? 128 : 256;
    const bool odd_grid = e->device_counted && device_counted_ctas(a, o, true) != canonical_ctas(e);
    if (odd_grid) CU(spread_step_count(e, e->stream));
    CU(do_step(e, a, o, e->stream, true));
    if (odd_grid) CU(spread_step_count(e, e->stream));
    after_step(e, n_steps);
    return GYMRS_OK;
}

int gymrs_get_state(gymrs_env *e, float *state, int32_t *sbt)
{
    if (!e || !state) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (int rc_ = refuse_in_capture(e, "gymrs_get_state")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    CU(cudaMemcpy2DAsync(state, sizeof(float) * e->n, e->state, sizeof(float) * e->ld,
                         sizeof(float) * e->n, e->state_dim, cudaMemcpyDeviceToHost, e->stream));
    if (sbt) {
        if (!e->sbt) return fail(GYMRS_ERR_UNSUPPORTED, "steps_beyond_terminated exists for CartPole only");
        CU(cudaMemcpyAsync(sbt, e->sbt, sizeof(int32_t) * e->n, cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    return GYMRS_OK;
}

int gymrs_set_state(gymrs_env *e, const float *state, const int32_t *sbt)
{
    if (!e || !state) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (int rc_ = refuse_in_capture(e, "gymrs_set_state")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    e->chain_ok = false;
    CU(cudaMemcpy2DAsync(e->state, sizeof(float) * e->ld, state, sizeof(float) * e->n,
                         sizeof(float) * e->n, e->state_dim, cudaMemcpyHostToDevice, e->stream));
    if (e->sbt) {
        if (sbt) {
            CU(cudaMemcpyAsync(e->sbt, sbt, sizeof(int32_t) * e->n, cudaMemcpyHostToDevice, e->stream));
            e->sbt_dirty = true;
        } else {
            CU(cudaMemsetAsync(e->sbt, 0xFF, sizeof(int32_t) * e->ld, e->stream));
            e->sbt_dirty = false;
        }
    } else if (sbt) {
        return fail(GYMRS_ERR_UNSUPPORTED, "steps_beyond_terminated exists for CartPole only");
    }
    if (e->elapsed) CU(cudaMemsetAsync(e->elapsed, 0, sizeof(uint32_t) * e->ld, e->stream));
    if (e->kind == GYMRS_PENDULUM) CU(launch_pendulum_obs(base_args(e), e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return GYMRS_OK;
}

// ---- checkpoint / resume ----------------------------------------------------------------
// Blob = CkptHeader (256 bytes) + the handle's per-env arrays packed densely ([rows][num_envs],
// no row padding), each section rounded up to 8 bytes.  The checksum covers header + payload.
} // extern "C"

namespace {

struct CkptHeader {
    char magic[8];
    uint32_t version, kind;
    uint64_t n, global_off;
    uint32_t flags, sbt_dirty;
    uint64_t seed, step_count;
    uint64_t total_bytes;
    uint64_t checksum; // of the whole blob with this field zero
    float reset_low[4], reset_high[4];
    unsigned char params[96]; // gymrs_<kind>_params, zero padded
    unsigned char reserved[56];
};
static_assert(sizeof(CkptHeader) == 256, "checkpoint header is 256 bytes");
static_assert(sizeof(gymrs_cartpole_params) <= 96 && sizeof(gymrs_mountain_car_params) <= 96 &&
              sizeof(gymrs_pendulum_params) <= 96, "params fit the header");
const char CKPT_MAGIC[8] = {'G', 'Y', 'M', 'R', 'S', 'C', 'K', 'P'};
constexpr uint32_t CKPT_VERSION = 1;

size_t pad8(size_t b) { return (b + 7) & ~(size_t)7; }

struct CkptLayout {
    size_t state, obs, reward, done, truncated, sbt, elapsed, total; // byte offsets; 0 = absent
};

CkptLayout ckpt_layout(int kind, uint64_t n, uint32_t flags)
{
    const uint32_t sd = kind == GYMRS_CARTPOLE ? 4 : 2;
    CkptLayout l = {};
    size_t off = sizeof(CkptHeader);
    l.state = off; off += pad8(sizeof(float) * sd * n);
    if (kind == GYMRS_PENDULUM) { l.obs = off; off += pad8(sizeof(float) * 3 * n); }
    l.reward = off; off += pad8(sizeof(float) * n);
    l.done = off; off += pad8(n);
    l.truncated = off; off += pad8(n);
    if (kind == GYMRS_CARTPOLE) { l.sbt = off; off += pad8(sizeof(int32_t) * n); }
    if (flags & GYMRS_FLAG_TIME_LIMIT) { l.elapsed = off; off += pad8(sizeof(uint32_t) * n); }
    l.total = off;
    return l;
}

// 64-bit multiply-xorshift hash over 8-byte words (every section is padded to 8 bytes)
uint64_t ckpt_hash(const unsigned char *p, size_t bytes, size_t skip_off)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)bytes;
    for (size_t i = 0; i + 8 <= bytes; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        if (i == skip_off) w = 0;
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 29;
    }
    return h;
}

int ckpt_parse(const void *buf, size_t bytes, CkptHeader *h)
{
    if (!buf) return fail(GYMRS_ERR_BAD_ARG, "checkpoint buffer is NULL");
    if (bytes < sizeof(CkptHeader)) return fail(GYMRS_ERR_BAD_ARG, "checkpoint blob shorter than its header");
    std::memcpy(h, buf, sizeof *h);
    if (std::memcmp(h->magic, CKPT_MAGIC, 8) != 0) return fail(GYMRS_ERR_BAD_ARG, "not a gymrs checkpoint (bad magic)");
    if (h->version != CKPT_VERSION) return fail(GYMRS_ERR_UNSUPPORTED, "unsupported checkpoint version");
    if (h->kind > (uint32_t)GYMRS_PENDULUM || h->n == 0 || h->n > (1ull << 31))
        return fail(GYMRS_ERR_BAD_ARG, "corrupt checkpoint header");
    const CkptLayout l = ckpt_layout((int)h->kind, h->n, h->flags);
    if (h->total_bytes != l.total || bytes < l.total) return fail(GYMRS_ERR_BAD_ARG, "checkpoint blob is truncated");
    if (ckpt_hash((const unsigned char *)buf, l.total, offsetof(CkptHeader, checksum)) != h->checksum)
        return fail(GYMRS_ERR_BAD_ARG, "checkpoint checksum mismatch");
    return GYMRS_OK;
}

} // namespace

extern "C" {

int gymrs_checkpoint_size(const gymrs_env *e, size_t *bytes)
{
    if (!e || !bytes) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *bytes = ckpt_layout(e->kind, e->n, e->flags).total;
    return GYMRS_OK;
}

int gymrs_checkpoint_info_of(const void *buf, size_t bytes, gymrs_checkpoint_info *info)
{
    if (!info) return fail(GYMRS_ERR_BAD_ARG, "info is NULL");
    CkptHeader h;
    if (int rc = ckpt_parse(buf, bytes, &h)) return rc;
    info->kind = (int32_t)h.kind;
    info->flags = h.flags;
    info->num_envs = h.n;
    info->global_env_offset = h.global_off;
    info->seed = h.seed;
    info->step_count = h.step_count;
    info->bytes = h.total_bytes;
    return GYMRS_OK;
}

int gymrs_checkpoint_save(gymrs_env *e, void *buf, size_t bytes)
{
    if (!e || !buf) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    const CkptLayout l = ckpt_layout(e->kind, e->n, e->flags);
    if (bytes < l.total) return fail(GYMRS_ERR_BAD_ARG, "checkpoint buffer too small (see gymrs_checkpoint_size)");
    if (int rc_ = refuse_in_capture(e, "gymrs_checkpoint_save")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    if (int rc_ = refresh_step_count(e)) return rc_;
    unsigned char *out = (unsigned char *)buf;
    std::memset(out, 0, l.total); // section padding is part of the checksum
    const uint64_t n = e->n;
    cudaStream_t s = e->stream;
    CU(cudaMemcpy2DAsync(out + l.state, sizeof(float) * n, e->state, sizeof(float) * e->ld, sizeof(float) * n,
                         e->state_dim, cudaMemcpyDeviceToHost, s));
    if (l.obs)
        CU(cudaMemcpy2DAsync(out + l.obs, sizeof(float) * n, e->obs, sizeof(float) * e->ld, sizeof(float) * n,
                             e->obs_dim, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out + l.reward, e->reward, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out + l.done, e->done, n, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(out + l.truncated, e->truncated, n, cudaMemcpyDeviceToHost, s));
    if (l.sbt) CU(cudaMemcpyAsync(out + l.sbt, e->sbt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    if (l.elapsed) CU(cudaMemcpyAsync(out + l.elapsed, e->elapsed, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
    CkptHeader h = {};
    std::memcpy(h.magic, CKPT_MAGIC, 8);
    h.version = CKPT_VERSION;
    h.kind = (uint32_t)e->kind;
    h.n = n;
    h.global_off = e->global_off;
    h.flags = e->flags;
    h.sbt_dirty = e->sbt_dirty ? 1u : 0u;
    h.seed = e->seed;
    h.step_count = e->step_count;
    h.total_bytes = l.total;
    std::memcpy(h.reset_low, e->reset_low, sizeof h.reset_low);
    std::memcpy(h.reset_high, e->reset_high, sizeof h.reset_high);
    if (e->kind == GYMRS_CARTPOLE) std::memcpy(h.params, &e->cp, sizeof e->cp);
    else if (e->kind == GYMRS_MOUNTAIN_CAR) std::memcpy(h.params, &e->mc, sizeof e->mc);
    else std::memcpy(h.params, &e->pd, sizeof e->pd);
    CU(cudaStreamSynchronize(s));
    std::memcpy(out, &h, sizeof h);
    h.checksum = ckpt_hash(out, l.total, offsetof(CkptHeader, checksum));
    std::memcpy(out, &h, sizeof h);
    return GYMRS_OK;
}

int gymrs_checkpoint_load(gymrs_env *e, const void *buf, size_t bytes)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    CkptHeader h;
    if (int rc = ckpt_parse(buf, bytes, &h)) return rc;
    if ((int)h.kind != e->kind) return fail(GYMRS_ERR_BAD_ARG, "checkpoint is of a different env kind");
    if (h.n != e->n) return fail(GYMRS_ERR_BAD_ARG, "checkpoint holds a different number of envs");
    if ((h.flags ^ e->flags) & GYMRS_FLAG_TIME_LIMIT)
        return fail(GYMRS_ERR_BAD_ARG, "checkpoint and handle differ in GYMRS_FLAG_TIME_LIMIT");
    if (int rc_ = refuse_in_capture(e, "gymrs_checkpoint_load")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    const CkptLayout l = ckpt_layout(e->kind, e->n, e->flags);
    const unsigned char *in = (const unsigned char *)buf;
    const uint64_t n = e->n;
    cudaStream_t s = e->stream;
    e->chain_ok = false;
    CU(cudaMemcpy2DAsync(e->state, sizeof(float) * e->ld, in + l.state, sizeof(float) * n, sizeof(float) * n,
                         e->state_dim, cudaMemcpyHostToDevice, s));
    if (l.obs)
        CU(cudaMemcpy2DAsync(e->obs, sizeof(float) * e->ld, in + l.obs, sizeof(float) * n, sizeof(float) * n,
                             e->obs_dim, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->reward, in + l.reward, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->done, in + l.done, n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->truncated, in + l.truncated, n, cudaMemcpyHostToDevice, s));
    if (l.sbt) CU(cudaMemcpyAsync(e->sbt, in + l.sbt, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    if (l.elapsed) CU(cudaMemcpyAsync(e->elapsed, in + l.elapsed, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->epoch_mem + 2, &h.seed, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    e->generation += 1; // the parameters are replaced below: graphs recorded on this handle are stale
    CU(cudaMemcpyAsync(e->epoch_mem + 3, &e->generation, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CU(fill_step_count(e, h.step_count, s));
    CU(cudaStreamSynchronize(s));
    e->global_off = h.global_off;
    e->seed = h.seed;
    e->step_count = h.step_count;
    e->sbt_dirty = h.sbt_dirty != 0;
    std::memcpy(e->reset_low, h.reset_low, sizeof h.reset_low);
    std::memcpy(e->reset_high, h.reset_high, sizeof h.reset_high);
    if (e->kind == GYMRS_CARTPOLE) std::memcpy(&e->cp, h.params, sizeof e->cp);
    else if (e->kind == GYMRS_MOUNTAIN_CAR) std::memcpy(&e->mc, h.params, sizeof e->mc);
    else std::memcpy(&e->pd, h.params, sizeof e->pd);
    fold_params(e);
    return GYMRS_OK;
}

int gymrs_checkpoint_create(const void *buf, size_t bytes, int device, gymrs_env **out)
{
    if (!out) return fail(GYMRS_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    CkptHeader h;
    if (int rc = ckpt_parse(buf, bytes, &h)) return rc;
    gymrs_env *e = nullptr;
    int rc = gymrs_create((int)h.kind, h.n, device, h.global_off, h.params, h.flags & GYMRS_FLAG_TIME_LIMIT, &e);
    if (rc != GYMRS_OK) return rc;
    rc = gymrs_checkpoint_load(e, buf, bytes);
    if (rc != GYMRS_OK) {
        std::string msg = g_last_error;
        free_env(e);
        return fail(rc, msg);
    }
    *out = e;
    return GYMRS_OK;
}

int gymrs_get_buffers(gymrs_env *e, gymrs_buffers *out)
{
    if (!e || !out) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    out->num_envs = e->n;
    out->ld = e->ld;
    out->state_dim = e->state_dim;
    out->obs_dim = e->obs_dim;
    out->state = e->state;
    out->obs = e->obs;
    out->reward = e->reward;
    out->done = e->done;
    out->truncated = e->truncated;
    out->steps_beyond_terminated = e->sbt;
    out->elapsed_steps = e->elapsed;
    return GYMRS_OK;
}

int gymrs_action_space(const gymrs_env *e, uint64_t *n, float *low, float *high)
{
    if (!e || !n) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    if (e->kind == GYMRS_CARTPOLE) *n = 2;          // Discrete(2), cartpole.rs:112
    else if (e->kind == GYMRS_MOUNTAIN_CAR) *n = 3; // Discrete(3), mountain_car.rs:363
    else {
        *n = 0;
        if (low) *low = (float)-e->pd.max_torque;
        if (high) *high = (float)e->pd.max_torque;
    }
    return GYMRS_OK;
}

int gymrs_observation_space(const gymrs_env *e, double *low, double *high)
{
    if (!e || !low || !high) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    const double inf = std::numeric_limits<double>::infinity();
    if (e->kind == GYMRS_CARTPOLE) { // cartpole.rs:105-113
        high[0] = e->cp.x_threshold * 2.;
        high[1] = inf;
        high[2] = e->cp.theta_threshold_radians * 2.;
        high[3] = inf;
        for (int i = 0; i < 4; ++i) low[i] = -high[i];
    } else if (e->kind == GYMRS_MOUNTAIN_CAR) { // mountain_car.rs:353-354
        low[0] = e->mc.min_position; low[1] = -e->mc.max_speed;
        high[0] = e->mc.max_position; high[1] = e->mc.max_speed;
    } else {
        high[0] = 1.; high[1] = 1.; high[2] = e->pd.max_speed;
        for (int i = 0; i < 3; ++i) low[i] = -high[i];
    }
    return GYMRS_OK;
}

int gymrs_reward_range(const gymrs_env *e, double *low, double *high)
{
    if (!e || !low || !high) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *low = -std::numeric_limits<double>::infinity(); // core.rs:16-19
    *high = std::numeric_limits<double>::infinity();
    return GYMRS_OK;
}

int gymrs_num_envs(const gymrs_env *e, uint64_t *n)
{
    if (!e || !n) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *n = e->n;
    return GYMRS_OK;
}

int gymrs_kind_of(const gymrs_env *e, int *kind)
{
    if (!e || !kind) return fail(GYMRS_ERR_BAD_ARG, "NULL argument");
    *kind = e->kind;
    return GYMRS_OK;
}

int gymrs_sync(gymrs_env *e, uint64_t *bad_env)
{
    if (!e) return fail(GYMRS_ERR_BAD_ARG, "NULL handle");
    if (int rc_ = refuse_in_capture(e, "gymrs_sync")) return rc_;
    ON_DEVICE(e->device);
    if (int rc_ = drain_host(e)) return rc_;
    CU(cudaStreamSynchronize(e->stream));
    if (e->err_host[3]) {
        e->err_host[3] = 0;
        e->chain_ok = false;
        CU(cudaMemsetAsync(e->chain_mem, 0, sizeof(uint32_t), e->stream));
        return fail(GYMRS_ERR_CUDA, "chained step launch timed out waiting for the previous step (pdl = 2 protocol error)");
    }
    if (e->err_host[5]) {
        e->err_host[5] = 0;
        return fail(GYMRS_ERR_UNSUPPORTED, "a CUDA graph whose steps were captured under a different seed was replayed: "
                    "captured steps bake the handle's Philox key in, so re-capture after re-seeding the handle");
    }
    if (e->err_host[6]) {
        e->err_host[6] = 0;
        return fail(GYMRS_ERR_UNSUPPORTED, "a CUDA graph recorded before gymrs_set_params (or before the handle's first step without "
                    "auto-reset) was replayed: captured steps bake the parameter block and the step variant in, so re-capture");
    }
    if (e->err_host[0]) {
        const uint64_t gid = (uint64_t)e->err_host[1] | ((uint64_t)e->err_host[2] << 32);
        if (bad_env) *bad_env = gid;
        e->err_host[0] = 0;
        char buf[128];
        // the reference's panic text: "{} usize invalid" (cartpole.rs:404) /
        // "{} (usize) invalid" (mountain_car.rs:404), followed by where it happened
        const int act = (int)e->err_host[4];
        std::snprintf(buf, sizeof buf, e->kind == GYMRS_MOUNTAIN_CAR ? "%d (usize) invalid (env %llu)" : "%d usize invalid (env %llu)",
                      act, (unsigned long long)gid);
        return fail(GYMRS_ERR_INVALID_ACTION, buf);
    }
    return GYMRS_OK;
}

int gymrs_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(GYMRS_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (gymrs_device_count() <= 0) return fail(GYMRS_ERR_NO_DEVICE, "no CUDA device visible");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return GYMRS_OK;
}

int gymrs_host_free(void *p)
{
    if (!p) return GYMRS_OK;
    CU(cudaFreeHost(p));
    return GYMRS_OK;
}

// utils/custom/util_fns.rs:2-10 on O64 = OrderedFloat<f64> (types.rs:4): same branch order, with
// OrderedFloat's total order (NaN == NaN, NaN greater than everything else), so a NaN value is
// clipped to the right bound.
double gymrs_clip(double value, double left_bound, double right_bound)
{
    auto le = [](double a, double b) { return a <= b || (b != b); };           // a <= b in the total order
    auto gt = [](double a, double b) { return a > b || (a != a && b == b); };  // a >  b in the total order
    if (le(left_bound, value) && le(value, right_bound)) return value;
    else if (gt(value, right_bound)) return right_bound;
    else return left_bound;
}

// spaces/discrete.rs:14-20
int gymrs_discrete_contains(uint64_t n, uint64_t value) { return value < n; }

// utils/seeding.rs:21-26
uint64_t gymrs_rand_random(const uint64_t *seed) { return seed ? *seed : entropy64(); }

} // extern "C"
