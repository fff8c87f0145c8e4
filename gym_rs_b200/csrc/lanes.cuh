// lanes.cuh -- the two arithmetic "lanes" the env dynamics are written against.
//
// Lane2 carries TWO env instances per operation in one 64-bit register pair and issues Blackwell's
// packed fp32 instructions (PTX fma.rn.f32x2 / mul.rn.f32x2 -> SASS FFMA2 / FMUL2, sm_100+): the
// step kernels sit close to the issue limit once the state is L2-resident or the launches of one
// stream cannot overlap, and packing halves the issue slots of the physics.
// Lane1 is the scalar form (one lane per env, ragged tails, slow paths).  Both lanes execute the SAME
// sequence of individually rounded IEEE operations -- Lane1 through the never-contracted
// __fmaf_rn / __fmul_rn intrinsics -- so every vector width produces the same bits
// (tests/test_parity_gpu.py::test_vec_widths_and_ragged_sizes_are_bit_identical).
//
// Rule for code written against a lane: only fma / mul (no add of a mul result): ptxas contracts a
// packed mul followed by a packed add into one FFMA2, which would round differently from Lane1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gymrs {

struct Lane1 {
    using T = float;
    static constexpr int W = 1;
    __device__ __forceinline__ static T bc(float c) { return c; }
    __device__ __forceinline__ static T fma(T a, T b, T c) { return __fmaf_rn(a, b, c); }
    __device__ __forceinline__ static T mul(T a, T b) { return __fmul_rn(a, b); }
    __device__ __forceinline__ static T neg(T a) { return -a; }
    __device__ __forceinline__ static T rcp(T a)
    {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
        return r;
    }
    __device__ __forceinline__ static T floor(T a) { return floorf(a); }
    // element access (slot 0 only)
    __device__ __forceinline__ static float get(T a, int) { return a; }
    __device__ __forceinline__ static uint32_t bits(T a, int) { return __float_as_uint(a); }
    template <class F> __device__ __forceinline__ static T map(T a, F f) { return f(a, 0); }
    template <class F> __device__ __forceinline__ static T map2(T a, T b, F f) { return f(a, b, 0); }
};

struct Lane2 {
    using T = uint64_t;
    static constexpr int W = 2;
    __device__ __forceinline__ static T pack(float lo, float hi)
    {
        T r;
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
        return r;
    }
    __device__ __forceinline__ static void unpack(T v, float &lo, float &hi)
    {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    }
    __device__ __forceinline__ static T bc(float c) { return pack(c, c); }
    __device__ __forceinline__ static T fma(T a, T b, T c)
    {
        T d;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
        return d;
    }
    __device__ __forceinline__ static T mul(T a, T b)
    {
        T d;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
        return d;
    }
    // ptxas folds this into the operand's negate modifier (FFMA2 -R.F32x2)
    __device__ __forceinline__ static T neg(T a)
    {
        float lo, hi;
        unpack(a, lo, hi);
        return pack(-lo, -hi);
    }
    __device__ __forceinline__ static T rcp(T a)
    {
        float lo, hi;
        unpack(a, lo, hi);
        return pack(Lane1::rcp(lo), Lane1::rcp(hi));
    }
    __device__ __forceinline__ static T floor(T a)
    {
        float lo, hi;
        unpack(a, lo, hi);
        return pack(floorf(lo), floorf(hi));
    }
    __device__ __forceinline__ static float get(T a, int i)
    {
        float lo, hi;
        unpack(a, lo, hi);
        return i ? hi : lo;
    }
    __device__ __forceinline__ static uint32_t bits(T a, int i) { return (uint32_t)(a >> (32 * i)); }
    // element-wise scalar functions (selects, clamps): f(value, element index)
    template <class F> __device__ __forceinline__ static T map(T a, F f)
    {
        float lo, hi;
        unpack(a, lo, hi);
        return pack(f(lo, 0), f(hi, 1));
    }
    template <class F> __device__ __forceinline__ static T map2(T a, T b, F f)
    {
        float alo, ahi, blo, bhi;
        unpack(a, alo, ahi);
        unpack(b, blo, bhi);
        return pack(f(alo, blo, 0), f(ahi, bhi, 1));
    }
};

// ---- sin / cos on a lane -----------------------------------------------------------------------
// Two minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf coefficients, <= 1 ulp there) ...
template <class L>
__device__ __forceinline__ void sincos_poly(typename L::T x, typename L::T &s, typename L::T &c)
{
    using T = typename L::T;
    const T z = L::mul(x, x);
    const T ps = L::fma(L::fma(L::bc(-1.9515295891e-4f), z, L::bc(8.3321608736e-3f)), z, L::bc(-1.6666654611e-1f));
    const T pc = L::fma(L::fma(L::bc(2.443315711809948e-5f), z, L::bc(-1.388731625493765e-3f)), z, L::bc(4.166664568298827e-2f));
    s = L::fma(L::mul(x, z), ps, x);
    c = L::fma(L::mul(z, z), pc, L::fma(z, L::bc(-0.5f), L::bc(1.0f)));
}

// ... and, for arguments up to TRIG_FAST_MAX, a quadrant reduction in front of them: k = rint(x * 2/pi)
// by the 1.5 * 2^23 magic add (the quadrant is then the low two bits of the biased float), r = x - k *
// pi/2 in three Cody-Waite steps (the first constant has 8 significant bits, so k * hi is exact).
// Absolute error <= ~1.5e-7 on the whole range -- the env tolerances are 1e-6.
constexpr float TRIG_FAST_MAX = 1024.0f;

template <class L>
__device__ __forceinline__ void sincos_reduced(typename L::T x, typename L::T &s, typename L::T &c)
{
    using T = typename L::T;
    const T biased = L::fma(x, L::bc(0.63661977236758134f), L::bc(12582912.0f));
    const T k = L::fma(biased, L::bc(1.0f), L::bc(-12582912.0f)); // exact: biased - magic
    T r = L::fma(k, L::bc(-1.5703125f), x);
    r = L::fma(k, L::bc(-4.837512969970703125e-4f), r);
    r = L::fma(k, L::bc(-7.54978995489188216e-8f), r);
    T ps, pc;
    sincos_poly<L>(r, ps, pc);
    // quadrant q = k mod 4:  sin x = {s, c, -s, -c}[q],  cos x = {c, -s, -c, s}[q]
    s = L::map2(ps, pc, [&](float sv, float cv, int i) {
        const uint32_t q = L::bits(biased, i);
        const float v = (q & 1u) ? cv : sv;
        return __uint_as_float(__float_as_uint(v) ^ ((q & 2u) << 30));
    });
    c = L::map2(ps, pc, [&](float sv, float cv, int i) {
        const uint32_t q = L::bits(biased, i);
        const float v = (q & 1u) ? sv : cv;
        return __uint_as_float(__float_as_uint(v) ^ (((q + 1u) & 2u) << 30));
    });
}

} // namespace gymrs
