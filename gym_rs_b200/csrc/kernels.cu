// kernels.cu -- sm_100a kernels of the batched classic-control step path.
//
// One thread owns V consecutive env instances (V = 4 by default: every SoA row is
// read and written with one 128-bit access per thread, 512 contiguous bytes per
// warp; V = 1 is the literal "one lane per env" mapping, used for unaligned caller
// buffers and tails).  The path is an elementwise map at ~0.9 flop/B, so there are
// no tensor cores here: the roofline is HBM bandwidth.
//
// Kernels
//   step_kernel     one transition of every env (+ same-launch auto-reset)
//                   reference: Env::step, cartpole.rs:398-483, mountain_car.rs:398-435
//   rollout_kernel  n_steps transitions with the state held in registers; actions are
//                   prefetched one step ahead, per-step results streamed out
//   reset_kernel    Env::reset, cartpole.rs:485-516, mountain_car.rs:464-501
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.hpp"

namespace gymrs {

namespace {

// ---- programmatic dependent launch (PDL) ----------------------------------
// wait: block until the previous grid in the stream has completed and its writes are
// visible.  Everything issued before it (parameter loads, the action prefetch -- the
// action buffer is never written by a step kernel) overlaps the previous kernel's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- V-wide row access ------------------------------------------------------
template <int V> struct Pack;
template <> struct Pack<1> { using F = float;  using I = int32_t; using B = uint8_t; using U = uint32_t; };
template <> struct Pack<2> { using F = float2; using I = int2;    using B = uchar2;  using U = uint2; };
template <> struct Pack<4> { using F = float4; using I = int4;    using B = uchar4;  using U = uint4; };

template <int V, class T, class PK>
__device__ __forceinline__ void unpack(const PK &v, T (&r)[V])
{
    static_assert(sizeof(PK) == sizeof(T) * V, "pack size");
    const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) r[j] = e[j];
}
template <int V, class T, class PK>
__device__ __forceinline__ PK pack(const T (&r)[V])
{
    PK v;
    T *e = reinterpret_cast<T *>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) e[j] = r[j];
    return v;
}

// plain (coherent) load: the row is overwritten in place by this same thread later
template <int V>
__device__ __forceinline__ void ld_row(const float *p, float (&r)[V], bool full, int nvalid)
{
    if (full) {
        unpack<V>(*reinterpret_cast<const typename Pack<V>::F *>(p), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? p[j] : 0.0f;
    }
}
template <int V>
__device__ __forceinline__ void ld_row(const uint32_t *p, uint32_t (&r)[V], bool full, int nvalid)
{
    if (full) {
        unpack<V>(*reinterpret_cast<const typename Pack<V>::U *>(p), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? p[j] : 0u;
    }
}
template <int V>
__device__ __forceinline__ void ld_row(const int32_t *p, int32_t (&r)[V], bool full, int nvalid)
{
    if (full) {
        unpack<V>(*reinterpret_cast<const typename Pack<V>::I *>(p), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? p[j] : -1;
    }
}
// read-only streaming load (actions are consumed once): ld.global.nc
template <int V, class T>
__device__ __forceinline__ void ld_stream(const T *p, T (&r)[V], bool full, int nvalid)
{
    static_assert(sizeof(T) == 4, "4-byte actions");
    if (full) {
        unpack<V>(__ldg(reinterpret_cast<const typename Pack<V>::U *>(p)), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? __ldg(p + j) : T(0);
    }
}
template <int V, bool STREAM = false>
__device__ __forceinline__ void st_row(float *p, const float (&r)[V], bool full, int nvalid)
{
    if (full) {
        using F = typename Pack<V>::F;
        if (STREAM) __stcs(reinterpret_cast<F *>(p), pack<V, float, F>(r));
        else *reinterpret_cast<F *>(p) = pack<V, float, F>(r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (j < nvalid) p[j] = r[j];
    }
}
template <int V>
__device__ __forceinline__ void st_row(uint32_t *p, const uint32_t (&r)[V], bool full, int nvalid)
{
    if (full) {
        using U = typename Pack<V>::U;
        *reinterpret_cast<U *>(p) = pack<V, uint32_t, U>(r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (j < nvalid) p[j] = r[j];
    }
}
template <int V>
__device__ __forceinline__ void st_row(int32_t *p, const int32_t (&r)[V], bool full, int nvalid)
{
    if (full) {
        using I = typename Pack<V>::I;
        *reinterpret_cast<I *>(p) = pack<V, int32_t, I>(r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (j < nvalid) p[j] = r[j];
    }
}
template <int V, bool STREAM = false>
__device__ __forceinline__ void st_row(uint8_t *p, const uint8_t (&r)[V], bool full, int nvalid)
{
    if (full) {
        using B = typename Pack<V>::B;
        if (STREAM) __stcs(reinterpret_cast<B *>(p), pack<V, uint8_t, B>(r));
        else *reinterpret_cast<B *>(p) = pack<V, uint8_t, B>(r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (j < nvalid) p[j] = r[j];
    }
}

__device__ __forceinline__ void report_invalid(uint32_t *err, uint64_t gid)
{
    // sticky; any one offender is reported (the reference panics on the first it meets)
    err[1] = (uint32_t)gid;
    err[2] = (uint32_t)(gid >> 32);
    __threadfence_system();
    err[0] = 1u;
}

// ---- one transition of the V envs a thread owns, all on registers ------------
template <class E, int V, bool AR, bool SBT, bool TL>
__device__ __forceinline__ void transition(const typename E::P &p, const BatchArgs &a,
                                           float (&s)[E::SD][V], float (&o)[E::OD][V],
                                           const typename E::Action (&act)[V],
                                           int32_t (&sbt)[V], uint32_t (&el)[V],
                                           float (&rew)[V], uint8_t (&dn)[V], uint8_t (&tr)[V],
                                           uint64_t gid0, uint64_t epoch, int nvalid)
{
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float sj[E::SD], oj[E::OD];
#pragma unroll
        for (int r = 0; r < E::SD; ++r) sj[r] = s[r][j];
#pragma unroll
        for (int r = 0; r < E::OD; ++r) oj[r] = o[r][j];
        float reward = 0.0f;
        bool done = false, trunc = false;
        if (E::valid(act[j])) {
            E::step(p, sj, act[j], oj, reward, done);
            if (SBT && E::HAS_SBT) {
                // reward 1.0 while alive and on the FIRST terminal step, 0.0 afterwards
                // (cartpole.rs:455-464); the state keeps integrating.
                if (done) {
                    if (sbt[j] < 0) { sbt[j] = 0; }
                    else { sbt[j] += 1; reward = 0.0f; }
                }
            }
            if (TL) {
                el[j] += 1u;
                trunc = el[j] >= a.max_steps;
            }
            if (AR && (done || trunc)) {
                // same-launch auto-reset (examples/cartpole.rs:23-28 does it by hand)
                E::reset(p, sj, oj, reset_words(a.seed, gid0 + j, epoch));
                if (SBT) sbt[j] = -1; //                                   cartpole.rs:504
                if (TL) el[j] = 0u;
            }
        } else if (j < nvalid) {
            report_invalid(a.err, gid0 + j); //                            cartpole.rs:402-406
        }
#pragma unroll
        for (int r = 0; r < E::SD; ++r) s[r][j] = sj[r];
#pragma unroll
        for (int r = 0; r < E::OD; ++r) o[r][j] = oj[r];
        rew[j] = reward;
        dn[j] = done ? 1 : 0;
        tr[j] = trunc ? 1 : 0;
    }
}

// ---- step -------------------------------------------------------------------
template <class E, int V, bool AR, bool SBT, bool TL>
__global__ void __launch_bounds__(256)
step_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a)
{
    using A = typename E::Action;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const bool live = i0 < a.n;
    const bool full = i0 + V <= a.n;
    const int nvalid = live ? (full ? V : (int)(a.n - i0)) : 0;

    A act[V];
    // pdl == 2: the caller guarantees the action batch predates the previous launch, so its
    // load latency can hide under that launch's tail
    if (a.early_actions && live) ld_stream<V>(reinterpret_cast<const A *>(a.actions) + i0, act, full, nvalid);

    // let the next launch in the stream get its CTAs scheduled now ...
    pdl_launch_dependents();
    // ... but touch nothing the previous launch writes (state rows, and possibly the action
    // batch) until it has completed and flushed
    pdl_wait();
    if (!live) return;
    if (!a.early_actions) ld_stream<V>(reinterpret_cast<const A *>(a.actions) + i0, act, full, nvalid);

    float s[E::SD][V], o[E::OD][V];
#pragma unroll
    for (int r = 0; r < E::SD; ++r) ld_row<V>(a.state + r * a.ld + i0, s[r], full, nvalid);
    int32_t sbt[V];
    uint32_t el[V];
    if (SBT) ld_row<V>(a.sbt + i0, sbt, full, nvalid);
    if (TL) ld_row<V>(a.elapsed + i0, el, full, nvalid);

    float rew[V];
    uint8_t dn[V], tr[V];
    transition<E, V, AR, SBT, TL>(p, a, s, o, act, sbt, el, rew, dn, tr, a.global_off + i0, a.epoch, nvalid);

#pragma unroll
    for (int r = 0; r < E::SD; ++r) st_row<V>(a.state + r * a.ld + i0, s[r], full, nvalid);
    if (!E::OBS_IS_STATE) {
#pragma unroll
        for (int r = 0; r < E::OD; ++r) st_row<V>(a.obs + r * a.ld + i0, o[r], full, nvalid);
    }
    st_row<V>(a.reward + i0, rew, full, nvalid);
    st_row<V>(a.done + i0, dn, full, nvalid);
    if (TL) st_row<V>(a.truncated + i0, tr, full, nvalid);
    if (SBT) st_row<V>(a.sbt + i0, sbt, full, nvalid);
    if (TL) st_row<V>(a.elapsed + i0, el, full, nvalid);
}

// ---- fused rollout ------------------------------------------------------------
// n_steps transitions in one launch.  State lives in registers; per step the kernel
// reads one action row and streams out observation / reward / done.  The actions of
// step k+1 are loaded before the math of step k so their latency is covered.
template <class E, int V, bool AR, bool SBT, bool TL>
__global__ void __launch_bounds__(256)
rollout_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a)
{
    using A = typename E::Action;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i0 >= a.n) return;
    const bool full = i0 + V <= a.n;
    const int nvalid = full ? V : (int)(a.n - i0);
    const A *actp = reinterpret_cast<const A *>(a.actions) + i0;

    A act[V], act_next[V];
    ld_stream<V>(actp, act, full, nvalid);

    float s[E::SD][V], o[E::OD][V];
#pragma unroll
    for (int r = 0; r < E::SD; ++r) ld_row<V>(a.state + r * a.ld + i0, s[r], full, nvalid);
    int32_t sbt[V];
    uint32_t el[V];
    if (SBT) ld_row<V>(a.sbt + i0, sbt, full, nvalid);
    if (TL) ld_row<V>(a.elapsed + i0, el, full, nvalid);
    float rew[V];
    uint8_t dn[V], tr[V];

    for (uint32_t k = 0; k < a.n_steps; ++k) {
        if (k + 1 < a.n_steps) ld_stream<V>(actp + (uint64_t)(k + 1) * a.act_ld, act_next, full, nvalid);
        transition<E, V, AR, SBT, TL>(p, a, s, o, act, sbt, el, rew, dn, tr, a.global_off + i0,
                                      a.epoch + k, nvalid);
        if (a.obs_out) {
            float *ob = a.obs_out + (uint64_t)k * E::OD * a.out_ld + i0;
#pragma unroll
            for (int r = 0; r < E::OD; ++r) {
                if constexpr (E::OBS_IS_STATE) st_row<V, true>(ob + r * a.out_ld, s[r], full, nvalid);
                else st_row<V, true>(ob + r * a.out_ld, o[r], full, nvalid);
            }
        }
        if (a.reward_out) st_row<V, true>(a.reward_out + (uint64_t)k * a.out_ld + i0, rew, full, nvalid);
        if (a.done_out) st_row<V, true>(a.done_out + (uint64_t)k * a.out_ld + i0, dn, full, nvalid);
#pragma unroll
        for (int j = 0; j < V; ++j) act[j] = act_next[j];
    }

#pragma unroll
    for (int r = 0; r < E::SD; ++r) st_row<V>(a.state + r * a.ld + i0, s[r], full, nvalid);
    if (!E::OBS_IS_STATE) {
#pragma unroll
        for (int r = 0; r < E::OD; ++r) st_row<V>(a.obs + r * a.ld + i0, o[r], full, nvalid);
    }
    st_row<V>(a.reward + i0, rew, full, nvalid);
    st_row<V>(a.done + i0, dn, full, nvalid);
    if (TL) st_row<V>(a.truncated + i0, tr, full, nvalid);
    if (SBT) st_row<V>(a.sbt + i0, sbt, full, nvalid);
    if (TL) st_row<V>(a.elapsed + i0, el, full, nvalid);
}

// ---- reset ----------------------------------------------------------------------
template <class E>
__global__ void __launch_bounds__(256)
reset_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a,
             const uint8_t *__restrict__ mask)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (mask && !mask[i]) return;
    float s[E::SD], o[E::OD];
    E::reset(p, s, o, reset_words(a.seed, a.global_off + i, 0));
#pragma unroll
    for (int r = 0; r < E::SD; ++r) a.state[r * a.ld + i] = s[r];
    if (!E::OBS_IS_STATE) {
#pragma unroll
        for (int r = 0; r < E::OD; ++r) a.obs[r * a.ld + i] = o[r];
    }
    a.reward[i] = 0.0f;
    a.done[i] = 0;
    if (a.sbt) a.sbt[i] = -1; //              cartpole.rs:504
    if (a.elapsed) { a.elapsed[i] = 0u; a.truncated[i] = 0; }
}

__global__ void __launch_bounds__(256) pendulum_obs_kernel(const __grid_constant__ BatchArgs a)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float sn, cs;
    sincosf(a.state[i], &sn, &cs);
    a.obs[i] = cs;
    a.obs[a.ld + i] = sn;
    a.obs[2 * a.ld + i] = a.state[a.ld + i];
}

// ---- launch helpers ----------------------------------------------------------------
template <class K, class P>
cudaError_t launch_ex(K kernel, uint64_t threads, int block, bool pdl, cudaStream_t s,
                      const P &p, BatchArgs a)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((threads + block - 1) / block));
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (!pdl) a.early_actions = 0;
    return cudaLaunchKernelEx(&cfg, kernel, p, a);
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// widest V the pointers of this launch allow
template <class E>
int pick_vec(const BatchArgs &a, int want, bool rollout)
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

template <class E, int V, bool ROLLOUT>
cudaError_t dispatch_flags(const typename E::P &p, const BatchArgs &a_in, const LaunchOpts &o, cudaStream_t s)
{
    const uint64_t threads = (a_in.n + V - 1) / V;
    const int block = (o.block >= 32 && o.block <= 256 && o.block % 32 == 0) ? o.block : 256;
    const bool sbt = E::HAS_SBT && o.use_sbt;
    BatchArgs a = a_in;
    a.early_actions = (o.pdl == 2);
    const int key = (o.autoreset ? 4 : 0) | (sbt ? 2 : 0) | (o.time_limit ? 1 : 0);
#define GYMRS_CASE(K, AR, SB, TL)                                                                   \
    case K:                                                                                         \
        return ROLLOUT ? launch_ex(rollout_kernel<E, V, AR, SB, TL>, threads, block, false, s, p, a) \
                       : launch_ex(step_kernel<E, V, AR, SB, TL>, threads, block, o.pdl != 0, s, p, a);
    switch (key) {
        GYMRS_CASE(0, false, false, false)
        GYMRS_CASE(1, false, false, true)
        GYMRS_CASE(2, false, true, false)
        GYMRS_CASE(3, false, true, true)
        GYMRS_CASE(4, true, false, false)
        GYMRS_CASE(5, true, false, true)
        GYMRS_CASE(6, true, true, false)
        GYMRS_CASE(7, true, true, true)
    }
#undef GYMRS_CASE
    return cudaErrorInvalidValue;
}

template <class E, bool ROLLOUT>
cudaError_t dispatch_vec(const typename E::P &p, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s)
{
    if (a.n == 0) return cudaSuccess;
    switch (pick_vec<E>(a, o.vec, ROLLOUT)) {
    case 4: return dispatch_flags<E, 4, ROLLOUT>(p, a, o, s);
    case 2: return dispatch_flags<E, 2, ROLLOUT>(p, a, o, s);
    default: return dispatch_flags<E, 1, ROLLOUT>(p, a, o, s);
    }
}

} // namespace

template <class E>
cudaError_t launch_step(const typename E::P &p, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s)
{
    return dispatch_vec<E, false>(p, a, o, s);
}

template <class E>
cudaError_t launch_rollout(const typename E::P &p, const BatchArgs &a, const LaunchOpts &o, cudaStream_t s)
{
    return dispatch_vec<E, true>(p, a, o, s);
}

template <class E>
cudaError_t launch_reset(const typename E::P &p, const BatchArgs &a, const uint8_t *mask, cudaStream_t s)
{
    if (a.n == 0) return cudaSuccess;
    reset_kernel<E><<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(p, a, mask);
    return cudaGetLastError();
}

cudaError_t launch_pendulum_obs(const BatchArgs &a, cudaStream_t s)
{
    if (a.n == 0) return cudaSuccess;
    pendulum_obs_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

#define GYMRS_INSTANTIATE(E)                                                                              \
    template cudaError_t launch_step<E>(const E::P &, const BatchArgs &, const LaunchOpts &, cudaStream_t);   \
    template cudaError_t launch_rollout<E>(const E::P &, const BatchArgs &, const LaunchOpts &, cudaStream_t); \
    template cudaError_t launch_reset<E>(const E::P &, const BatchArgs &, const uint8_t *, cudaStream_t);
GYMRS_INSTANTIATE(CartPole)
GYMRS_INSTANTIATE(MountainCar)
GYMRS_INSTANTIATE(Pendulum)

} // namespace gymrs
