// philox.cuh -- counter-based RNG for reset sampling (device side).
//
// Philox4x32-10 (Salmon et al., SC'11): the same function as curand's
// curand_Philox4x32_10, written out so that one call costs 20 IMAD.WIDE and no
// state has to live in HBM.  reset_words(seed, gid, epoch) is bit-identical to
// curand_init(seed, /*subsequence*/ epoch, /*offset*/ 4 * gid, &st); curand4(&st)
// (tests/test_curand_gpu.py checks it against cuRAND's own device generator).  The reference re-seeds a fresh PCG64 on every
// reset (cartpole.rs:491-494, mountain_car.rs:470-473, seeding.rs:21-26), i.e.
// a reset is a pure function of the seed; here it is a pure function of
// (seed, global env id, epoch), which keeps that property per env and makes
// results independent of how the batch is sharded over GPUs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gymrs {

// The ten round keys (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85).  The key is the seed, the same
// for every env of a launch, so the host expands it once (philox_round_keys) and the kernel reads
// the schedule from its parameter block instead of re-deriving it per reset.
struct PhiloxKeys {
    uint32_t k[10][2];
};

inline PhiloxKeys philox_round_keys(uint64_t seed)
{
    PhiloxKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.k[r][0] = k0;
        rk.k[r][1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return rk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ rk.k[r][0], lo1, hi0 ^ c.w ^ rk.k[r][1], lo0);
    }
    return c;
}

// The four words that env `gid` draws at `epoch` (0 = explicit reset,
// 1 + step index = auto-reset during that step).
__device__ __forceinline__ uint4 reset_words(const PhiloxKeys &rk, uint64_t gid, uint64_t epoch)
{
    return philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32),
                                    (uint32_t)epoch, (uint32_t)(epoch >> 32)), rk);
}

// U[low, low + scale) on a 2^-24 grid: the reference draws from the half-open
// range (rand Uniform::new, cartpole.rs:363, mountain_car.rs:189).  scale24 = scale * 2^-24
// (exact), so this is low + (w >> 8) * 2^-24 * scale in one FMA.  `cap` is the largest float
// below high, so rounding in the fma can never return high itself.
__device__ __forceinline__ float uniform_from_word(uint32_t w, float low, float scale24, float cap)
{
    return fminf(fmaf((float)(w >> 8), scale24, low), cap);
}

} // namespace gymrs
