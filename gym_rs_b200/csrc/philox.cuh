// philox.cuh -- counter-based RNG for reset sampling (device side).
//
// Philox4x32-10 (Salmon et al., SC'11): the same function as curand's
// curand_Philox4x32_10, written out so that one call costs 20 IMAD.WIDE and no
// state has to live in HBM.  The reference re-seeds a fresh PCG64 on every
// reset (cartpole.rs:491-494, mountain_car.rs:470-473, seeding.rs:21-26), i.e.
// a reset is a pure function of the seed; here it is a pure function of
// (seed, global env id, epoch), which keeps that property per env and makes
// results independent of how the batch is sharded over GPUs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gymrs {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// The four words that env `gid` draws at `epoch` (0 = explicit reset,
// 1 + step index = auto-reset during that step).
__device__ __forceinline__ uint4 reset_words(uint64_t seed, uint64_t gid, uint64_t epoch)
{
    return philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32),
                                    (uint32_t)epoch, (uint32_t)(epoch >> 32)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// U[low, low + scale) on a 2^-24 grid: the reference draws from the half-open
// range (rand Uniform::new, cartpole.rs:363, mountain_car.rs:189).  `cap` is the
// largest float below high, so rounding in the fma can never return high itself.
__device__ __forceinline__ float uniform_from_word(uint32_t w, float low, float scale, float cap)
{
    const float r = (float)(w >> 8) * 5.9604644775390625e-08f; // 2^-24
    return fminf(fmaf(r, scale, low), cap);
}

} // namespace gymrs
