// kernels_impl.cuh -- sm_100a kernels of the batched classic-control step path (templates;
// instantiated per env in kernels_cartpole.cu, kernels_mountain_car.cu, kernels_pendulum.cu).
//
// One thread owns V consecutive env instances (V = 4 by default: every SoA row is read and
// written with one 128-bit access per thread, 512 contiguous bytes per warp; V = 1 is the literal
// "one lane per env" mapping, used for unaligned caller buffers).  The path is an elementwise map
// at ~0.9 flop/B, so there are no tensor cores here: the roofline is HBM bandwidth.
//
// Kernels
//   step_kernel     one transition of every env (+ same-launch auto-reset)
//                   reference: Env::step, cartpole.rs:398-483, mountain_car.rs:398-435
//   rollout_kernel  n_steps transitions with the state held in registers; actions are
//                   prefetched one step ahead, per-step results streamed out
//   reset_kernel    Env::reset, cartpole.rs:485-516, mountain_car.rs:464-501
//
// Launch-to-launch pipelining (step_kernel).  A 1M-env step moves only ~43 MB, i.e. ~6.5 us of
// HBM time, so the ramp-up / drain of a stand-alone launch (~3 us) is a large fraction of it.
// Steps are therefore launched with programmatic dependent launch (PDL) and, in chained mode,
// do NOT wait for the whole previous grid: env i of step t+1 depends only on env i of step t,
// which the CTA with the same index wrote.  Each CTA publishes a per-handle flag (release) when
// its stores are done and the same-index CTA of the next step acquires it before loading.
// Consecutive steps then overlap like one continuous stream of CTAs.  Deadlock-free because a
// dependent grid is only launched once every CTA of its predecessor has started.
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.hpp"

namespace gymrs {

namespace {

// Env index inside one launch.  A handle holds at most 2^31 envs (gymrs_create), so 32 bits do:
// one register instead of two per index, and one IMAD / LEA less per address.  Global ids (the
// Philox counter) stay 64-bit: global_off + index.
using idx_t = uint32_t;

// ---- programmatic dependent launch (PDL) ----------------------------------
// wait: block until the previous grid in the stream has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v)
{
#ifdef GYMRS_DEBUG_RELAXED_PUBLISH // measurement only: NOT a correct publish
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// ---- V-wide row access ------------------------------------------------------
template <int V> struct Pack;
template <> struct Pack<1> { using F = float;  using I = int32_t; using B = uint8_t; using U = uint32_t; };
template <> struct Pack<2> { using F = float2; using I = int2;    using B = uchar2;  using U = uint2; };
template <> struct Pack<4> { using F = float4; using I = int4;    using B = uchar4;  using U = uint4; };
template <int V, class T> struct PackOf;
template <int V> struct PackOf<V, float>    { using type = typename Pack<V>::F; };
template <int V> struct PackOf<V, int32_t>  { using type = typename Pack<V>::I; };
template <int V> struct PackOf<V, uint32_t> { using type = typename Pack<V>::U; };
template <int V> struct PackOf<V, uint8_t>  { using type = typename Pack<V>::B; };

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

// FULL: all V elements are in range -> one vector access; otherwise guarded scalar accesses
// (only the single ragged group at the end of a batch takes that path).
//
// Row loads use ld.global.cg (L2 only): every line is touched once per step, and bypassing the
// non-coherent L1 keeps chained launches from ever seeing a stale line.
template <int V, bool FULL, class T>
__device__ __forceinline__ void ld_row(const T *p, T (&r)[V], int nvalid, T fill)
{
    if (FULL) {
        using PK = typename PackOf<V, T>::type;
        unpack<V>(__ldcg(reinterpret_cast<const PK *>(p)), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? __ldcg(p + j) : fill;
    }
}
// read-only streaming load (actions are consumed once): ld.global.nc
template <int V, bool FULL, class T>
__device__ __forceinline__ void ld_stream(const T *p, T (&r)[V], int nvalid)
{
    static_assert(sizeof(T) == 4, "4-byte actions");
    if (FULL) {
        unpack<V>(__ldg(reinterpret_cast<const typename Pack<V>::U *>(p)), r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) r[j] = (j < nvalid) ? __ldg(p + j) : T(0);
    }
}
// STREAM: st.global.cs (evict-first) for outputs nobody re-reads on the device
template <int V, bool FULL, bool STREAM = false, class T>
__device__ __forceinline__ void st_row(T *p, const T (&r)[V], int nvalid)
{
    if (FULL) {
        using PK = typename PackOf<V, T>::type;
        if (STREAM) __stcs(reinterpret_cast<PK *>(p), pack<V, T, PK>(r));
        else *reinterpret_cast<PK *>(p) = pack<V, T, PK>(r);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) if (j < nvalid) p[j] = r[j];
    }
}

__device__ __forceinline__ void report_invalid(uint32_t *err, uint64_t gid, uint32_t action_bits)
{
    // sticky; any one offender is reported (the reference panics on the first it meets)
    err[1] = (uint32_t)gid;
    err[2] = (uint32_t)(gid >> 32);
    err[4] = action_bits;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(err) = 1u;
}

// ---- step counter on the device (BatchArgs::epoch_dev) ----------------------------
// DEVC = false (every eager launch of a handle that was never captured): the epoch is the kernel
// argument the host counted; the device copies are not touched and the code below compiles away.
// DEVC = true (the handle's steps have been captured into a CUDA graph, where kernel arguments
// are frozen but every replay is a new step): the count lives in HBM, one private copy per CTA
// (epoch_dev[4 + blockIdx.x]).  Every thread loads its CTA's copy right after the dependency wait
// (one request per warp, 2048 different lines per grid) and nothing uses the value before the
// reset loop, so the load rides along with the row loads; thread 0 advances the copy behind one
// barrier at the very end, when every warp of the CTA has its value.  Only CTA b ever touches
// copy b within a grid and the next device-counted launch reads it behind a grid-wide dependency
// (such launches never chain), so there is no atomic and no hot spot.  Measured alternatives
// (profiles/r01_sweeps.md): one shared counter + an atomicAdd per CTA, per-CTA copies handed
// through shared memory at the top or behind the row loads, a lazy read inside the reset loop.
// The host keeps the copies of CTAs a launch geometry does not cover up to date (capi.cu,
// spread_step_count).
template <bool DEVC>
__device__ __forceinline__ uint64_t first_epoch(const BatchArgs &a)
{
    if constexpr (DEVC) return __ldcg(a.epoch_dev + 4 + blockIdx.x) + 1u;
    else return a.epoch;
}
template <bool DEVC>
__device__ __forceinline__ void advance_epoch(const BatchArgs &a, uint64_t first, uint32_t n_steps)
{
    if constexpr (DEVC) {
        __syncthreads(); // every warp of the CTA has read the copy
        if (threadIdx.x == 0) {
            a.epoch_dev[4 + blockIdx.x] = first - 1u + n_steps;
            // epoch_dev[2] is the seed of the handle's last full reset.  The Philox keys of this
            // launch were frozen when it was captured; if the handle has been re-seeded since, the
            // replay drew from the old stream -- raise the sticky error word (gymrs_sync reports it).
            if (blockIdx.x == 0 && __ldcg(a.epoch_dev + 2) != ((uint64_t)a.rk.k[0][0] | ((uint64_t)a.rk.k[0][1] << 32)))
                *reinterpret_cast<volatile uint32_t *>(a.err + 5) = 1u;
            // likewise the by-value parameter block and the step variant (capi.cu, bump_generation)
            if (blockIdx.x == 0 && __ldcg(a.epoch_dev + 3) != a.generation)
                *reinterpret_cast<volatile uint32_t *>(a.err + 6) = 1u;
        }
    }
}

// ---- one transition of the V envs a thread owns, all on registers ------------
#ifndef GYMRS_FAST_LANE
#define GYMRS_FAST_LANE 2 // 2: packed pairs (Lane2); 1: the same straight-line path on scalar lanes (A/B builds)
#endif

// Straight-line condition: every env of this thread has a valid action and a state inside the range
// of the branch-free trig.  Always true for envs that are reset when they end (a live CartPole has
// |theta| <= 0.21, a MountainCar position is clipped, a Pendulum angle is stored wrapped).
template <class E, int V>
__device__ __forceinline__ bool all_fast(const typename E::P &p, const float (&s)[E::SD][V], const typename E::Action (&act)[V])
{
    bool fast = true;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float sj[E::SD];
#pragma unroll
        for (int r = 0; r < E::SD; ++r) sj[r] = s[r][j];
        fast = fast & E::valid(act[j]) & E::fast_ok(p, sj);
    }
    return fast;
}

// The dynamics of V envs on the straight-line path, two at a time on packed fp32 (Lane2: FFMA2 /
// FMUL2, one issue slot per two FMAs).
template <class E, int V>
__device__ __forceinline__ void advance_fast(const typename E::P &p, float (&s)[E::SD][V], float (&o)[E::OD][V],
                                             const typename E::Action (&act)[V], float (&rew)[V])
{
    static_assert(V % 2 == 0, "pairs");
    if constexpr (GYMRS_FAST_LANE == 1) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float sj[E::SD], oj[E::OD];
#pragma unroll
            for (int r = 0; r < E::SD; ++r) sj[r] = s[r][j];
            E::template advance<Lane1>(p, sj, E::pre(p, act[j]), oj, rew[j]);
#pragma unroll
            for (int r = 0; r < E::SD; ++r) s[r][j] = sj[r];
            if (!E::OBS_IS_STATE) {
#pragma unroll
                for (int r = 0; r < E::OD; ++r) o[r][j] = oj[r];
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < V / 2; ++q) {
            typename Lane2::T ps[E::SD], po[E::OD], pr;
#pragma unroll
            for (int r = 0; r < E::SD; ++r) ps[r] = Lane2::pack(s[r][2 * q], s[r][2 * q + 1]);
            E::template advance<Lane2>(p, ps, Lane2::pack(E::pre(p, act[2 * q]), E::pre(p, act[2 * q + 1])), po, pr);
#pragma unroll
            for (int r = 0; r < E::SD; ++r) Lane2::unpack(ps[r], s[r][2 * q], s[r][2 * q + 1]);
            if (!E::OBS_IS_STATE) {
#pragma unroll
                for (int r = 0; r < E::OD; ++r) Lane2::unpack(po[r], o[r][2 * q], o[r][2 * q + 1]);
            }
            Lane2::unpack(pr, rew[2 * q], rew[2 * q + 1]);
        }
    }
}

// Per-env bookkeeping after the dynamics: reward beyond termination, time limit, flags.
// Returns true when the env ended its episode in this step (done or truncated).
template <class E, bool SBT, bool TL>
__device__ __forceinline__ bool settle(const BatchArgs &a, float reward, bool done, int32_t &sbt, uint32_t &el,
                                       float &rew, uint8_t &dn, uint8_t &tr)
{
    bool trunc = false;
    if (SBT && E::HAS_SBT) {
        // reward 1.0 while alive and on the FIRST terminal step, 0.0 afterwards
        // (cartpole.rs:455-464); the state keeps integrating.
        if (done) {
            if (sbt < 0) { sbt = 0; }
            else { sbt += 1; reward = 0.0f; }
        }
    }
    if (TL) {
        el += 1u;
        trunc = el >= a.max_steps;
    }
    rew = reward;
    dn = done ? 1 : 0;
    tr = trunc ? 1 : 0;
    return done | trunc;
}

// The rare per-env path of a fused rollout (full-range libm trig for a far-out angle or position),
// out of line so that its ~100-instruction range reduction does not sit inside every env slot.
template <class E>
__device__ __noinline__ void step_out_of_line(const typename E::P &p, float (&s)[E::SD], typename E::Action a,
                                              float (&o)[E::OD], float &reward, bool &done)
{
    E::step(p, s, a, o, reward, done);
}

// One transition of the V envs held in registers.  Finished envs are NOT re-sampled here; the mask
// of them (bit j, only j < nvalid) is returned: a step overwrites them after its row stores
// (reset_patch), a rollout merges fresh states into its registers (reset_merge).
// ASSUME_FAST: the caller has already established all_fast().
template <class E, int V, bool AR, bool SBT, bool TL, bool ASSUME_FAST = false>
__device__ __forceinline__ uint32_t transition(const typename E::P &p, const BatchArgs &a,
                                               float (&s)[E::SD][V], float (&o)[E::OD][V],
                                               const typename E::Action (&act)[V],
                                               int32_t (&sbt)[V], uint32_t (&el)[V],
                                               float (&rew)[V], uint8_t (&dn)[V], uint8_t (&tr)[V],
                                               uint64_t gid0, int nvalid)
{
    uint32_t ended = 0;
    bool fast = false;
    if constexpr (V % 2 == 0) fast = ASSUME_FAST || all_fast<E, V>(p, s, act);
    if (fast) {
        if constexpr (V % 2 == 0) {
            advance_fast<E, V>(p, s, o, act, rew);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float sj[E::SD];
#pragma unroll
                for (int r = 0; r < E::SD; ++r) sj[r] = s[r][j];
                if (settle<E, SBT, TL>(a, rew[j], E::terminal(p, sj), sbt[j], el[j], rew[j], dn[j], tr[j])) ended |= 1u << j;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float sj[E::SD], oj[E::OD];
#pragma unroll
            for (int r = 0; r < E::SD; ++r) sj[r] = s[r][j];
#pragma unroll
            for (int r = 0; r < E::OD; ++r) oj[r] = o[r][j];
            float reward = 0.0f;
            bool done = false;
            if (E::valid(act[j])) {
                if (V % 2 == 0) step_out_of_line<E>(p, sj, act[j], oj, reward, done);
                else E::step(p, sj, act[j], oj, reward, done);
                if (settle<E, SBT, TL>(a, reward, done, sbt[j], el[j], rew[j], dn[j], tr[j])) ended |= 1u << j;
            } else {
                if (j < nvalid) report_invalid(a.err, gid0 + j, (uint32_t)act[j]); //   cartpole.rs:402-406
                rew[j] = 0.0f;
                dn[j] = 0;
                tr[j] = 0;
            }
#pragma unroll
            for (int r = 0; r < E::SD; ++r) s[r][j] = sj[r];
#pragma unroll
            for (int r = 0; r < E::OD; ++r) o[r][j] = oj[r];
        }
    }
    return AR ? (ended & ((1u << nvalid) - 1u)) : 0u;
}

// Same-launch auto-reset (examples/cartpole.rs:23-28 does it by hand).  Only a few percent of envs
// end per step, so instead of a predicated Philox per slot (which nearly every warp would execute V
// times) each thread draws once per finished env: the loop runs as often as the busiest lane of the
// warp needs, typically once.
//
// reset_merge: the fresh state goes into the registers (fused rollout: the envs keep stepping).
template <class E, int V, bool SBT, bool TL>
__device__ __forceinline__ void reset_merge(const typename E::P &p, const BatchArgs &a, float (&s)[E::SD][V],
                                            float (&o)[E::OD][V], int32_t (&sbt)[V], uint32_t (&el)[V],
                                            uint64_t gid0, uint64_t epoch, uint32_t ended)
{
    while (ended) {
        const int j = __ffs(ended) - 1;
        ended &= ended - 1;
        float sj[E::SD], oj[E::OD];
        E::reset(p, sj, oj, reset_words(a.rk, gid0 + j, epoch));
#pragma unroll
        for (int jj = 0; jj < V; ++jj) {
            if (jj == j) {
#pragma unroll
                for (int r = 0; r < E::SD; ++r) s[r][jj] = sj[r];
                if (!E::OBS_IS_STATE) {
#pragma unroll
                    for (int r = 0; r < E::OD; ++r) o[r][jj] = oj[r];
                }
                if (SBT) sbt[jj] = -1; //                              cartpole.rs:504
                if (TL) el[jj] = 0u;
            }
        }
    }
}

// reset_patch: a single step stores its rows first and then overwrites each finished env i0 + j with
// its fresh state (same thread, same address: program order).  The state does not have to stay live
// across the Philox evaluation (fewer registers) and the row stores are issued before it; the lines
// are still in L2, so the 4-byte stores cost no extra DRAM traffic.
// FROM_THREAD: i0 = (blockIdx.x * blockDim.x + threadIdx.x) * V is re-derived from the special
// registers inside the loop (volatile, so that it is not merged with the first evaluation) rather
// than kept live across the row stores -- at 48 registers ptxas would otherwise spill it.
template <class E, bool SBT, bool TL, int FROM_THREAD = 0>
__device__ __forceinline__ void reset_patch(const typename E::P &p, const BatchArgs &a, idx_t i0, uint64_t epoch,
                                            uint32_t ended)
{
    while (ended) {
        const int j = __ffs(ended) - 1;
        ended &= ended - 1;
        if (FROM_THREAD) {
            uint32_t t, c, n;
            asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
            asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(c));
            asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(n));
            i0 = (c * n + t) * (idx_t)FROM_THREAD;
        }
        float sj[E::SD], oj[E::OD];
        const idx_t i = i0 + (idx_t)j;
        E::reset(p, sj, oj, reset_words(a.rk, a.global_off + i, epoch));
#pragma unroll
        for (int r = 0; r < E::SD; ++r) a.state[r * a.ld + i] = sj[r];
        if (!E::OBS_IS_STATE) {
#pragma unroll
            for (int r = 0; r < E::OD; ++r) a.obs[r * a.ld + i] = oj[r];
        }
        if (SBT) a.sbt[i] = -1; //                                          cartpole.rs:504
        if (TL) a.elapsed[i] = 0u;
    }
}

// ---- step -------------------------------------------------------------------
#ifndef GYMRS_STEP_MIN_CTAS
#define GYMRS_STEP_MIN_CTAS 5 // <= 51 registers/thread: 5 CTAs of 256 threads resident per SM
#endif

// store the results of V envs (rows of the handle's SoA arrays)
template <class E, int V, bool SBT, bool TL, bool FULL>
__device__ __forceinline__ void store_rows(const BatchArgs &a, idx_t i0, int nvalid, const float (&s)[E::SD][V],
                                           const float (&o)[E::OD][V], const int32_t (&sbt)[V], const uint32_t (&el)[V],
                                           const float (&rew)[V], const uint8_t (&dn)[V], const uint8_t (&tr)[V])
{
#pragma unroll
    for (int r = 0; r < E::SD; ++r) st_row<V, FULL>(a.state + r * a.ld + i0, s[r], nvalid);
    if (!E::OBS_IS_STATE) {
#pragma unroll
        for (int r = 0; r < E::OD; ++r) st_row<V, FULL>(a.obs + r * a.ld + i0, o[r], nvalid);
    }
    st_row<V, FULL>(a.reward + i0, rew, nvalid);
    st_row<V, FULL>(a.done + i0, dn, nvalid);
    if (TL) st_row<V, FULL>(a.truncated + i0, tr, nvalid);
    if (SBT) st_row<V, FULL>(a.sbt + i0, sbt, nvalid);
    if (TL) st_row<V, FULL>(a.elapsed + i0, el, nvalid);
}

// One env at a time, everything from and to global memory, scalar lanes: the general form of a step
// (any action, any state, any alignment).  The V = 1 kernels run it inline; the vector kernels call
// the out-of-line copy below for the rare thread that cannot take the straight-line path and for the
// ragged tail of a batch, so that none of this shapes the hot path's registers.
template <class E, bool AR, bool SBT, bool TL>
__device__ __forceinline__ void step_env_scalar(const typename E::P &p, const BatchArgs &a, idx_t i, uint64_t epoch)
{
    using A = typename E::Action;
    A act[1] = {__ldg(reinterpret_cast<const A *>(a.actions) + i)};
    float s[E::SD][1], o[E::OD][1];
#pragma unroll
    for (int r = 0; r < E::SD; ++r) s[r][0] = __ldcg(a.state + r * a.ld + i);
    int32_t sbt[1] = {-1};
    uint32_t el[1] = {0u};
    if (SBT) sbt[0] = __ldcg(a.sbt + i);
    if (TL) el[0] = __ldcg(a.elapsed + i);
    float rew[1];
    uint8_t dn[1], tr[1];
    const uint32_t ended = transition<E, 1, AR, SBT, TL>(p, a, s, o, act, sbt, el, rew, dn, tr, a.global_off + i, 1);
    store_rows<E, 1, SBT, TL, true>(a, i, 1, s, o, sbt, el, rew, dn, tr);
    if (AR) reset_patch<E, SBT, TL>(p, a, i, epoch, ended);
}

template <class E, bool AR, bool SBT, bool TL>
__device__ __noinline__ void step_envs_out_of_line(const typename E::P &p, const BatchArgs &a, idx_t i0, int count,
                                                   uint64_t epoch)
{
    for (int j = 0; j < count; ++j) step_env_scalar<E, AR, SBT, TL>(p, a, i0 + j, epoch);
}

template <class E, int V, bool AR, bool SBT, bool TL, bool FULL>
__device__ __forceinline__ void step_body(const typename E::P &p, const BatchArgs &a, idx_t i0,
                                          int nvalid, typename E::Action (&act)[V], uint64_t epoch)
{
    if constexpr (V == 1) {
        step_env_scalar<E, AR, SBT, TL>(p, a, i0, epoch);
    } else if constexpr (!FULL) {
        step_envs_out_of_line<E, AR, SBT, TL>(p, a, i0, nvalid, epoch); // the one ragged group at the end of a batch
    } else {
        float s[E::SD][V], o[E::OD][V];
#pragma unroll
        for (int r = 0; r < E::SD; ++r) ld_row<V, true>(a.state + r * a.ld + i0, s[r], V, 0.0f);
        int32_t sbt[V];
        uint32_t el[V];
        if (SBT) ld_row<V, true>(a.sbt + i0, sbt, V, int32_t(-1));
        if (TL) ld_row<V, true>(a.elapsed + i0, el, V, uint32_t(0));
        if (!all_fast<E, V>(p, s, act)) {
            step_envs_out_of_line<E, AR, SBT, TL>(p, a, i0, V, epoch);
            return;
        }
        float rew[V];
        uint8_t dn[V], tr[V];
        const uint32_t ended = transition<E, V, AR, SBT, TL, true>(p, a, s, o, act, sbt, el, rew, dn, tr, a.global_off + i0, V);
#ifndef GYMRS_RESET_PATCH
// 0 (default): fresh states are merged into the registers before the row stores.  1: the rows are stored
// first and finished envs overwritten afterwards (reset_patch) -- measured slower: the scattered 4-byte
// stores cost more than they save once the state is L2-resident (profiles/r02_sweeps.md).
#define GYMRS_RESET_PATCH 0
#endif
        if (AR && !GYMRS_RESET_PATCH) reset_merge<E, V, SBT, TL>(p, a, s, o, sbt, el, a.global_off + i0, epoch, ended);
        store_rows<E, V, SBT, TL, true>(a, i0, V, s, o, sbt, el, rew, dn, tr);
        if (AR && GYMRS_RESET_PATCH) reset_patch<E, SBT, TL, V>(p, a, 0, epoch, ended);
    }
}

// MINB: minimum number of 256-thread CTAs per SM, i.e. the register budget (5 -> 48 registers, the
// default; E::WIDE_MIN_CTAS -> 40 / 32, the opt-in high-occupancy build, LaunchOpts::wide).  Same
// source, same arithmetic, same bits; only the occupancy differs (profiles/r02_sweeps.md).
template <class E, int V, bool AR, bool SBT, bool TL, bool DEVC, int MINB = GYMRS_STEP_MIN_CTAS>
__global__ void __launch_bounds__(256, MINB)
step_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a)
{
    using A = typename E::Action;
    const idx_t n = (idx_t)a.n;
    const idx_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * V; // < 2^31 + 1024: no wrap
    const bool live = i0 < n;
    const bool full = live && n - i0 >= (idx_t)V;
    const int nvalid = live ? (full ? V : (int)(n - i0)) : 0;
    const A *actp = reinterpret_cast<const A *>(a.actions) + i0;

    // Before any dependency is resolved: pull this CTA's input rows into L2 with one bulk prefetch
    // per row (cp.async.bulk.prefetch.L2).  Always safe -- L2 is the point of coherence, so a line
    // the previous launch is still writing is simply updated in place -- and it lets the HBM reads
    // of launch t + 1 overlap the tail of launch t even when the launch has to wait for the whole
    // previous grid (pdl = 1): after the wait the row loads are L2 hits.
    if (a.l2_prefetch && threadIdx.x <= E::SD) {
        const idx_t cta0 = blockIdx.x * blockDim.x * V;
        if (cta0 < n) {
            const idx_t left = n - cta0, span = blockDim.x * V;
            const uint32_t bytes = (uint32_t)((left < span ? left : span) * 4u) & ~15u;
            const void *src = threadIdx.x < E::SD
                                  ? static_cast<const void *>(a.state + threadIdx.x * a.ld + cta0)
                                  : static_cast<const void *>(reinterpret_cast<const A *>(a.actions) + cta0);
            if (bytes && (reinterpret_cast<uintptr_t>(src) & 15u) == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        }
    }

    A act[V];
    // pdl == 2: the caller guarantees the action batch predates the previous launch, so it can
    // be fetched before any dependency is resolved
    // (the scalar forms -- V = 1, the ragged group at the end -- fetch their own actions)
    if (V > 1 && a.early_actions && full) ld_stream<V, true>(actp, act, nvalid);

    // let the next launch in the stream get its CTAs scheduled as soon as SM slots free up
    pdl_launch_dependents();

    if (a.chain) {
        // fine-grained dependency: these envs were last written by CTA blockIdx.x of the previous
        // step of this handle; wait for that CTA only, not for the whole previous grid
        if (threadIdx.x == 0) {
            const uint32_t want = a.chain_seq - 1u;
            // chain_flags[-1] is the handle's "protocol broken" word: once any CTA has timed out,
            // nobody spins again (a bug must surface as an error code, never as a hung device)
            uint32_t spins = 0;
            while (ld_acquire_gpu(a.chain_flags + blockIdx.x) != want) {
                __nanosleep(32);
                if ((++spins & 1023u) == 0 && (spins > (1u << 17) || ld_acquire_gpu(a.chain_flags - 1) != 0u)) {
                    st_release_gpu(a.chain_flags - 1, 1u);
                    *reinterpret_cast<volatile uint32_t *>(a.err + 3) = 1u;
                    break;
                }
            }
        }
        __syncthreads();
    } else {
        // touch nothing the previous launch writes until it has completed and flushed
        pdl_wait();
    }

    const uint64_t epoch = first_epoch<DEVC>(a);
    if (live) {
        if (V > 1 && !a.early_actions && full) ld_stream<V, true>(actp, act, nvalid);
        if (full) step_body<E, V, AR, SBT, TL, true>(p, a, i0, nvalid, act, epoch);
        else step_body<E, V, AR, SBT, TL, false>(p, a, i0, nvalid, act, epoch);
    }
    advance_epoch<DEVC>(a, epoch, 1u);

    // publish "this CTA's envs are at step chain_seq" for the next chained launch.  The barrier
    // orders every thread's stores before thread 0's release (cumulativity), so one release store
    // per CTA is enough.  Skipped entirely outside pdl == 2 so plain launches pay nothing.
    if (a.publish) {
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(a.chain_flags + blockIdx.x, a.chain_seq);
    }
}

// ---- step, persistent TMA-staged stream ------------------------------------------
// Same transition, different data movement (opt-in: launch config vec = 8).  The grid is
// persistent: G = (#SMs x CTAs per SM) CTAs, CTA c walks the tiles c, c + G, c + 2G, ... of 1024
// envs.  Each CTA keeps a ring of STREAM_STAGES tiles in shared memory: its control warp issues a
// bulk async copy (cp.async.bulk, the 1-D TMA path -- UBLKCP in SASS) per input row of a tile,
// completing on that stage's `full` mbarrier, while the 8 compute warps consume landed tiles with
// conflict-free LDS.128 and store results straight to global.  Compared with step_kernel this
//   * always has tiles in flight per CTA, held in shared memory instead of registers, so the
//     bytes in flight per SM no longer depend on how many register-heavy CTAs fit,
//   * occupies only part of each SM, so with PDL the next launch's CTAs are co-resident from the
//     start and stream in right behind (its trigger fires immediately: the grid is one wave),
//   * replaces per-thread 64-bit address arithmetic + LDG by one LDS per row,
//   * keeps flag waits and the release fence of chained launches off the compute warps.
// Chained-launch flags are per tile (1024 envs), the same granularity as step_kernel<V = 4> with
// 256-thread CTAs, so the two kernels can follow each other in a chain.  Measured:
// a little faster than step_kernel for chained launches on one stream, slower when the batch is
// L2-resident (a tile is worked on by only 8 warps) -- hence opt-in.
constexpr int TMA_TILE = 1024; // envs per tile = 256 threads x 4
#ifndef GYMRS_STREAM_STAGES
#define GYMRS_STREAM_STAGES 2 // measured, one stream chained, CartPole: 2 stages 7.5 us, 3: 8.1, 4: 9.8 (profiles/r01_sweeps.md)
#endif
constexpr int STREAM_STAGES = GYMRS_STREAM_STAGES;

template <class E, bool SBT, bool TL>
struct TmaCfg {
    static constexpr int ROWS = E::SD + 1 + (SBT ? 1 : 0) + (TL ? 1 : 0); // state rows, action row, optional rows
    static constexpr size_t SMEM = (size_t)STREAM_STAGES * ROWS * TMA_TILE * 4 + 2 * STREAM_STAGES * 8; // + full/done mbarriers
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// spin until chain_flags[t] == want (thread 0 only); bounded, see step_kernel
__device__ __forceinline__ void chain_wait(const BatchArgs &a, uint64_t t, uint32_t want)
{
    uint32_t spins = 0;
    while (ld_acquire_gpu(a.chain_flags + t) != want) {
        __nanosleep(32);
        if ((++spins & 1023u) == 0 && (spins > (1u << 17) || ld_acquire_gpu(a.chain_flags - 1) != 0u)) {
            st_release_gpu(a.chain_flags - 1, 1u);
            *reinterpret_cast<volatile uint32_t *>(a.err + 3) = 1u;
            break;
        }
    }
}

#ifndef GYMRS_STREAM_MIN_CTAS
#define GYMRS_STREAM_MIN_CTAS 4
#endif
constexpr int STREAM_COMPUTE_THREADS = 256;                      // 8 warps, 4 envs per thread = one tile
constexpr int STREAM_THREADS = STREAM_COMPUTE_THREADS + 32;      // + one control warp

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: warps 0-7 compute, warp 8 (one lane) is the control warp that owns every
// inter-CTA duty -- waiting on the previous step's tile flags, issuing the bulk copies, and
// publishing finished tiles with a release store.  The release fence and the flag spins therefore
// never stall a compute warp; compute warps only wait on the stage's `full` mbarrier and signal
// `done` (256 arrivals) when they have read the stage and issued their stores.
template <class E, bool AR, bool SBT, bool TL>
__global__ void __launch_bounds__(STREAM_THREADS, GYMRS_STREAM_MIN_CTAS)
step_stream_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a)
{
    using A = typename E::Action;
    constexpr int V = 4, S = STREAM_STAGES;
    constexpr int ROWS = TmaCfg<E, SBT, TL>::ROWS;
    constexpr int R_ACT = E::SD, R_SBT = E::SD + 1, R_EL = E::SD + 1 + (SBT ? 1 : 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float(*tile)[ROWS][TMA_TILE] = reinterpret_cast<float(*)[ROWS][TMA_TILE]>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)S * ROWS * TMA_TILE * 4);
    uint64_t *done = full + S;

    const uint64_t ntiles = (a.n + TMA_TILE - 1) / TMA_TILE;
    const uint64_t G = gridDim.x;
    // this CTA's k-th tile is tile_of(k); it owns K of them
    const uint32_t K = blockIdx.x < ntiles ? (uint32_t)((ntiles - blockIdx.x + G - 1) / G) : 0u;
    auto tile_of = [&](uint32_t k) -> uint64_t { return blockIdx.x + (uint64_t)k * G; };
    auto tile_envs = [&](uint64_t t) -> uint32_t { // envs in tile t (n % 4 == 0 is a launch precondition)
        const uint64_t left = a.n - t * TMA_TILE;
        return left < (uint64_t)TMA_TILE ? (uint32_t)left : (uint32_t)TMA_TILE;
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < S; ++st) {
            mbar_init(&full[st], 1);
            mbar_init(&done[st], STREAM_COMPUTE_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    if (threadIdx.x >= STREAM_COMPUTE_THREADS) {
        // ------------------------------- control warp -------------------------------
        if (threadIdx.x != STREAM_COMPUTE_THREADS) return;
        // arm the stage's barrier and (optionally) fetch the action row
        auto arm = [&](uint32_t k) {
            const uint64_t t = tile_of(k);
            const uint32_t bytes = tile_envs(t) * 4u;
            mbar_expect_tx(&full[k % S], ROWS * bytes);
            if (a.early_actions)
                bulk_g2s(tile[k % S][R_ACT], reinterpret_cast<const A *>(a.actions) + t * TMA_TILE, bytes, &full[k % S]);
        };
        // fetch everything the previous step of this handle may have written
        auto fetch = [&](uint32_t k) {
            const uint64_t t = tile_of(k), e0 = t * TMA_TILE;
            const uint32_t bytes = tile_envs(t) * 4u;
            const int st = k % S;
#pragma unroll
            for (int r = 0; r < E::SD; ++r) bulk_g2s(tile[st][r], a.state + r * a.ld + e0, bytes, &full[st]);
            if (!a.early_actions) bulk_g2s(tile[st][R_ACT], reinterpret_cast<const A *>(a.actions) + e0, bytes, &full[st]);
            if (SBT) bulk_g2s(tile[st][R_SBT], a.sbt + e0, bytes, &full[st]);
            if (TL) bulk_g2s(tile[st][R_EL], a.elapsed + e0, bytes, &full[st]);
        };
        // prologue: fill the ring
        const uint32_t K0 = K < (uint32_t)S ? K : (uint32_t)S;
        for (uint32_t k = 0; k < K0; ++k) arm(k);
        if (a.chain) {
            // per-tile dependency on the same tile of the handle's previous step; the rows were
            // written through the generic proxy and are read below through the async proxy
            for (uint32_t k = 0; k < K0; ++k) chain_wait(a, tile_of(k), a.chain_seq - 1u);
            asm volatile("fence.proxy.async;" ::: "memory");
        } else {
            pdl_wait();
        }
        for (uint32_t k = 0; k < K0; ++k) fetch(k);
        // steady state: when tile k is done, refill its stage, then publish it
        for (uint32_t k = 0; k < K; ++k) {
            mbar_wait(&done[k % S], (k / S) & 1u);
            if (k + S < K) {
                arm(k + S);
                if (a.chain) {
                    chain_wait(a, tile_of(k + S), a.chain_seq - 1u);
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                fetch(k + S);
            }
            // the mbarrier hand-over orders every compute thread's stores before this release
            if (a.publish) st_release_gpu(a.chain_flags + tile_of(k), a.chain_seq);
        }
        return;
    }

    // --------------------------------- compute warps ---------------------------------
    // without a per-tile chain, stores must come after the previous grid has completed
    if (!a.chain) pdl_wait();
    const uint32_t il = threadIdx.x * V; // first env of this thread inside a tile
#pragma unroll 1
    for (uint32_t k = 0; k < K; ++k) {
        const int st = k % S;
        const uint64_t t = tile_of(k);
        mbar_wait(&full[st], (k / S) & 1u);
        if (il < tile_envs(t)) {
            const idx_t i0 = (idx_t)t * TMA_TILE + il;
            float s[E::SD][V], o[E::OD][V];
            A act[V];
            int32_t sbt[V];
            uint32_t el[V];
#pragma unroll
            for (int r = 0; r < E::SD; ++r) unpack<V>(*reinterpret_cast<const float4 *>(&tile[st][r][il]), s[r]);
            unpack<V>(*reinterpret_cast<const uint4 *>(&tile[st][R_ACT][il]), act);
            if (SBT) unpack<V>(*reinterpret_cast<const int4 *>(&tile[st][R_SBT][il]), sbt);
            if (TL) unpack<V>(*reinterpret_cast<const uint4 *>(&tile[st][R_EL][il]), el);

            if (!all_fast<E, V>(p, s, act)) {
                step_envs_out_of_line<E, AR, SBT, TL>(p, a, i0, V, a.epoch); // re-reads this thread's envs from global
            } else {
                float rew[V];
                uint8_t dn[V], tr[V];
                const uint32_t ended = transition<E, V, AR, SBT, TL, true>(p, a, s, o, act, sbt, el, rew, dn, tr,
                                                                               a.global_off + i0, V);
                store_rows<E, V, SBT, TL, true>(a, i0, V, s, o, sbt, el, rew, dn, tr);
                if (AR) reset_patch<E, SBT, TL>(p, a, i0, a.epoch, ended);
            }
        }
        mbar_arrive(&done[st]); // this thread has read the stage and issued its stores for tile k
    }
}

// ---- fused rollout ------------------------------------------------------------
// n_steps transitions in one launch.  State lives in registers; per step the kernel
// reads one action row and streams out observation / reward / done.  The actions of
// step k+1 are loaded before the math of step k so their latency is covered.
template <class E, int V, bool AR, bool SBT, bool TL, bool FULL>
__device__ __forceinline__ void rollout_body(const typename E::P &p, const BatchArgs &a, idx_t i0, int nvalid,
                                             uint64_t epoch)
{
    using A = typename E::Action;
    const A *actp = reinterpret_cast<const A *>(a.actions) + i0;
    A act[V], act_next[V];
    ld_stream<V, FULL>(actp, act, nvalid);

    float s[E::SD][V], o[E::OD][V];
#pragma unroll
    for (int r = 0; r < E::SD; ++r) ld_row<V, FULL>(a.state + r * a.ld + i0, s[r], nvalid, 0.0f);
    int32_t sbt[V];
    uint32_t el[V];
    if (SBT) ld_row<V, FULL>(a.sbt + i0, sbt, nvalid, int32_t(-1));
    if (TL) ld_row<V, FULL>(a.elapsed + i0, el, nvalid, uint32_t(0));
    float rew[V];
    uint8_t dn[V], tr[V];

    for (uint32_t k = 0; k < a.n_steps; ++k) {
        if (k + 1 < a.n_steps) ld_stream<V, FULL>(actp + (uint64_t)(k + 1) * a.act_ld, act_next, nvalid);
        const uint32_t ended = transition<E, V, AR, SBT, TL>(p, a, s, o, act, sbt, el, rew, dn, tr, a.global_off + i0, nvalid);
        if (AR) reset_merge<E, V, SBT, TL>(p, a, s, o, sbt, el, a.global_off + i0, epoch + k, ended);
        if (a.obs_out) {
            float *ob = a.obs_out + (uint64_t)k * E::OD * a.out_ld + i0;
#pragma unroll
            for (int r = 0; r < E::OD; ++r) {
                if constexpr (E::OBS_IS_STATE) st_row<V, FULL, true>(ob + r * a.out_ld, s[r], nvalid);
                else st_row<V, FULL, true>(ob + r * a.out_ld, o[r], nvalid);
            }
        }
        if (a.reward_out) st_row<V, FULL, true>(a.reward_out + (uint64_t)k * a.out_ld + i0, rew, nvalid);
        if (a.done_out) st_row<V, FULL, true>(a.done_out + (uint64_t)k * a.out_ld + i0, dn, nvalid);
#pragma unroll
        for (int j = 0; j < V; ++j) act[j] = act_next[j];
    }

#pragma unroll
    for (int r = 0; r < E::SD; ++r) st_row<V, FULL>(a.state + r * a.ld + i0, s[r], nvalid);
    if (!E::OBS_IS_STATE) {
#pragma unroll
        for (int r = 0; r < E::OD; ++r) st_row<V, FULL>(a.obs + r * a.ld + i0, o[r], nvalid);
    }
    st_row<V, FULL>(a.reward + i0, rew, nvalid);
    st_row<V, FULL>(a.done + i0, dn, nvalid);
    if (TL) st_row<V, FULL>(a.truncated + i0, tr, nvalid);
    if (SBT) st_row<V, FULL>(a.sbt + i0, sbt, nvalid);
    if (TL) st_row<V, FULL>(a.elapsed + i0, el, nvalid);
}

template <class E, int V, bool AR, bool SBT, bool TL, bool DEVC>
__global__ void __launch_bounds__(256)
rollout_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a)
{
    const idx_t n = (idx_t)a.n;
    const idx_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * V;
    const uint64_t epoch = first_epoch<DEVC>(a);
    if (i0 < n && n - i0 >= (idx_t)V) rollout_body<E, V, AR, SBT, TL, true>(p, a, i0, V, epoch);
    else if (i0 < n) rollout_body<E, V, AR, SBT, TL, false>(p, a, i0, (int)(n - i0), epoch);
    advance_epoch<DEVC>(a, epoch, a.n_steps);
}

// ---- reset ----------------------------------------------------------------------
template <class E>
__global__ void __launch_bounds__(256)
reset_kernel(const __grid_constant__ typename E::P p, const __grid_constant__ BatchArgs a,
             const uint8_t *__restrict__ mask)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (!mask && a.epoch_from_dev) { // a full reset restarts the device step counters under this seed
        for (uint64_t c = i; c < a.epoch_slots; c += a.n) a.epoch_dev[4 + c] = 0;
        if (i == 0) a.epoch_dev[2] = (uint64_t)a.rk.k[0][0] | ((uint64_t)a.rk.k[0][1] << 32);
    }
    if (mask && !mask[i]) return;
    float s[E::SD], o[E::OD];
    E::reset(p, s, o, reset_words(a.rk, a.global_off + i, 0));
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
    if (!pdl) { a.early_actions = 0; a.chain = 0; a.l2_prefetch = 0; }
    return cudaLaunchKernelEx(&cfg, kernel, p, a);
}

template <class E, int V, bool ROLLOUT, bool DEVC, bool WIDE = false>
cudaError_t dispatch_flags(const typename E::P &p, const BatchArgs &a_in, const LaunchOpts &o, cudaStream_t s)
{
    constexpr int MINB = WIDE ? E::WIDE_MIN_CTAS : GYMRS_STEP_MIN_CTAS;
    const uint64_t threads = (a_in.n + V - 1) / V;
    const int block = pick_block(o);
    const bool sbt = E::HAS_SBT && o.use_sbt;
    BatchArgs a = a_in;
    a.early_actions = (o.pdl == 2);
    a.l2_prefetch = use_l2_prefetch(o);
    const int key = (o.autoreset ? 4 : 0) | (sbt ? 2 : 0) | (o.time_limit ? 1 : 0);
#define GYMRS_CASE(K, AR, SB, TL)                                                                   \
    case K:                                                                                         \
        return ROLLOUT ? launch_ex(rollout_kernel<E, V, AR, SB, TL, DEVC>, threads, block, false, s, p, a) \
                       : launch_ex(step_kernel<E, V, AR, SB, TL, DEVC, MINB>, threads, block, o.pdl != 0, s, p, a);
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

template <class K>
cudaError_t allow_big_smem(K kernel, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// persistent grid size: SMs of the current device x CTAs per SM (GYMRS_STREAM_CTAS_PER_SM, default 2)
inline int stream_grid_slots()
{
    static const int per_sm = [] {
        const char *e = std::getenv("GYMRS_STREAM_CTAS_PER_SM");
        const int v = e ? std::atoi(e) : 2;
        return v >= 1 && v <= 8 ? v : 2;
    }();
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * per_sm;
}

template <class E>
cudaError_t dispatch_tma(const typename E::P &p, const BatchArgs &a_in, const LaunchOpts &o, cudaStream_t s)
{
    const bool sbt = E::HAS_SBT && o.use_sbt;
    BatchArgs a = a_in;
    a.early_actions = (o.pdl == 2);
    if (o.pdl == 0) { a.early_actions = 0; a.chain = 0; }
    const uint64_t tiles = (a.n + TMA_TILE - 1) / TMA_TILE;
    cudaLaunchConfig_t cfg = {};
    const uint64_t slots = (uint64_t)stream_grid_slots();
    cfg.gridDim = dim3((unsigned)(tiles < slots ? tiles : slots));
    cfg.blockDim = dim3(STREAM_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = o.pdl != 0 ? 1 : 0;
    const int key = (o.autoreset ? 4 : 0) | (sbt ? 2 : 0) | (o.time_limit ? 1 : 0);
#define GYMRS_CASE(K, AR, SB, TL)                                                      \
    case K: {                                                                          \
        auto kern = step_stream_kernel<E, AR, SB, TL>;                                    \
        cfg.dynamicSmemBytes = TmaCfg<E, SB, TL>::SMEM;                                \
        cudaError_t e = allow_big_smem(kern, cfg.dynamicSmemBytes);                    \
        if (e != cudaSuccess) return e;                                                \
        return cudaLaunchKernelEx(&cfg, kern, p, a);                                   \
    }
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
    if (!ROLLOUT && use_tma_step(a, o) && !a.epoch_from_dev) return dispatch_tma<E>(p, a, o, s);
    const int v = pick_vec(a, o.vec, ROLLOUT);
    if (a.epoch_from_dev) { // device-counted (CUDA-graph) launches: the 128-bit path or the scalar one
        if (v == 4) return dispatch_flags<E, 4, ROLLOUT, true>(p, a, o, s);
        return dispatch_flags<E, 1, ROLLOUT, true>(p, a, o, s);
    }
    // the high-occupancy build exists for the 128-bit host-counted step only
    if (!ROLLOUT && v == 4 && o.wide) return dispatch_flags<E, 4, false, false, true>(p, a, o, s);
    switch (v) {
    case 4: return dispatch_flags<E, 4, ROLLOUT, false>(p, a, o, s);
    case 2: return dispatch_flags<E, 2, ROLLOUT, false>(p, a, o, s);
    default: return dispatch_flags<E, 1, ROLLOUT, false>(p, a, o, s);
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

// Each env is instantiated in its own translation unit (kernels_<env>.cu) so the three compile
// in parallel.
#define GYMRS_INSTANTIATE(E)                                                                              \
    template cudaError_t launch_step<E>(const E::P &, const BatchArgs &, const LaunchOpts &, cudaStream_t);   \
    template cudaError_t launch_rollout<E>(const E::P &, const BatchArgs &, const LaunchOpts &, cudaStream_t); \
    template cudaError_t launch_reset<E>(const E::P &, const BatchArgs &, const uint8_t *, cudaStream_t);

} // namespace gymrs
