// mode_floor: what does each STEPPING MODE cost for a kernel that only moves CartPole's bytes?
//
// The step path is HBM-bound, but a 1M-env step is only ~43 MB (~6.7 us of HBM time), so how a
// caller strings launches together matters as much as the kernel: two streams, one stream with a
// grid-wide dependency (pdl = 1), an L2-resident batch, a launch timed alone.  This tool times
// data-movement-only prototypes (same rows, same in-place pattern, a token FMA per element instead
// of the dynamics) next to the library's real CartPole step in every mode, on one box, so that
// "fraction of the mode's floor" can be stated and alternative data-movement schemes can be ranked
// before the dynamics are ported to them.
//
//   plain16   step_kernel's scheme (one thread = 4 envs, 128-bit LDG/STG, L2 prefetch before the
//             dependency wait), natural occupancy (16 CTAs of 128 per SM: the whole grid is one wave)
//   plain10   the same with shared-memory padding so that 10 CTAs fit per SM, step_kernel's occupancy
//             at 48 registers (1.38 waves)
//   chunk     persistent: one CTA per SM (or two), every input row of the CTA's env range fetched
//             into shared memory at the very start with bulk async copies (one mbarrier per
//             1024-env slice), warps consume slices as they land and store with STG.128
//   real      gymrs_step (auto-reset) through the C ABI
// --env mountain_car | pendulum: the plain variants for that env's rows (2 state rows; Pendulum + 3 obs rows)
//
// Build (here):  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/mode_floor \
//                  tools/mode_floor.cu -Iinclude -Lgym_rs_b200 -lgymrs_b200 -Xlinker -rpath='$ORIGIN/../../gym_rs_b200'
// Run (GPU box): tools/bin/mode_floor [--env cartpole|mountain_car|pendulum] [--iso-only] [--k 2000]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "gymrs_b200.h"

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            std::exit(1);                                                                         \
        }                                                                                         \
    } while (0)

struct Slot {
    float *state; // 4 rows, leading dimension n
    const int32_t *act;
    float *reward;
    uint8_t *done;
};

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float4 touch(float4 v, int4 a, int r)
{
    const float k = 0.999f;
    v.x = fmaf(v.x, k, a.x ? 1e-3f : -1e-3f);
    v.y = fmaf(v.y, k, a.y ? 1e-3f : -1e-3f);
    v.z = fmaf(v.z, k, a.z ? 1e-3f : -1e-3f);
    v.w = fmaf(v.w, k, (a.w + r) ? 1e-3f : -1e-3f);
    return v;
}

// ---- step_kernel's data movement ---------------------------------------------------------------
__global__ void __launch_bounds__(128) copy_plain(Slot s, uint32_t n, int prefetch)
{
    extern __shared__ unsigned char pad[];
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (prefetch && threadIdx.x <= 4) {
        const uint32_t cta0 = blockIdx.x * blockDim.x * 4u;
        const void *src = threadIdx.x < 4 ? (const void *)(s.state + (size_t)threadIdx.x * n + cta0) : (const void *)(s.act + cta0);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(blockDim.x * 16u) : "memory");
    }
    pdl_launch_dependents();
    pdl_wait();
    if (i0 >= n) return;
    float4 v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = __ldcg(reinterpret_cast<const float4 *>(s.state + (size_t)r * n + i0));
    const int4 a = __ldg(reinterpret_cast<const int4 *>(s.act + i0));
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = touch(v[r], a, r);
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<float4 *>(s.state + (size_t)r * n + i0) = v[r];
    *reinterpret_cast<float4 *>(s.reward + i0) = make_float4(v[0].x, v[1].y, v[2].z, v[3].w);
    *reinterpret_cast<uchar4 *>(s.done + i0) = make_uchar4(v[0].x > 0, v[0].y > 0, v[0].z > 0, v[0].w > 0);
}

// the same scheme for any env's rows: SR state rows updated in place, the action row read, OR observation rows
// written (Pendulum: cos, sin, theta_dot; 0 where the observation IS the state), reward and done written
template <int SR, int OR>
__global__ void __launch_bounds__(128) copy_rows(Slot s, float *obs, uint32_t n, int prefetch)
{
    extern __shared__ unsigned char pad[];
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (prefetch && threadIdx.x <= SR) {
        const uint32_t cta0 = blockIdx.x * blockDim.x * 4u;
        const void *src = threadIdx.x < SR ? (const void *)(s.state + (size_t)threadIdx.x * n + cta0) : (const void *)(s.act + cta0);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(blockDim.x * 16u) : "memory");
    }
    pdl_launch_dependents();
    pdl_wait();
    if (i0 >= n) return;
    float4 v[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) v[r] = __ldcg(reinterpret_cast<const float4 *>(s.state + (size_t)r * n + i0));
    const int4 a = __ldg(reinterpret_cast<const int4 *>(s.act + i0));
#pragma unroll
    for (int r = 0; r < SR; ++r) v[r] = touch(v[r], a, r);
#pragma unroll
    for (int r = 0; r < SR; ++r) *reinterpret_cast<float4 *>(s.state + (size_t)r * n + i0) = v[r];
#pragma unroll
    for (int r = 0; r < OR; ++r) *reinterpret_cast<float4 *>(obs + (size_t)r * n + i0) = touch(v[r % SR], a, r + 1);
    *reinterpret_cast<float4 *>(s.reward + i0) = make_float4(v[0].x, v[SR - 1].y, v[0].z, v[SR - 1].w);
    *reinterpret_cast<uchar4 *>(s.done + i0) = make_uchar4(v[0].x > 0, v[0].y > 0, v[0].z > 0, v[0].w > 0);
}

// ---- persistent, whole input range of the CTA staged in shared memory at t = 0 ----------------------
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
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int SLICE = 1024; // envs per slice: 256 threads x 4
constexpr int ROWS = 5;     // 4 state rows + the action row

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) copy_chunk(Slot s, uint32_t n, int max_slices, int prefetch)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float(*tile)[ROWS][SLICE] = reinterpret_cast<float(*)[ROWS][SLICE]>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)max_slices * ROWS * SLICE * 4);

    // env range of this CTA, in warp units of 128 envs (n % 128 == 0 here)
    const uint32_t units = n / 128u;
    const uint32_t u0 = (uint32_t)((uint64_t)blockIdx.x * units / gridDim.x);
    const uint32_t u1 = (uint32_t)((uint64_t)(blockIdx.x + 1) * units / gridDim.x);
    const uint32_t e0 = u0 * 128u, cnt = (u1 - u0) * 128u;
    const int nsl = (int)((cnt + SLICE - 1) / SLICE);

    if (prefetch && threadIdx.x < ROWS) {
        const void *src = threadIdx.x < 4 ? (const void *)(s.state + (size_t)threadIdx.x * n + e0) : (const void *)(s.act + e0);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(cnt * 4u) : "memory");
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < nsl; ++k) mbar_init(&bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    if ((int)threadIdx.x < nsl) { // one lane per slice issues that slice's bulk copies
        const int k = threadIdx.x;
        const uint32_t off = (uint32_t)k * SLICE, len = min((uint32_t)SLICE, cnt - off), bytes = len * 4u;
        mbar_expect_tx(&bar[k], ROWS * bytes);
#pragma unroll
        for (int r = 0; r < 4; ++r) bulk_g2s(tile[k][r], s.state + (size_t)r * n + e0 + off, bytes, &bar[k]);
        bulk_g2s(tile[k][4], s.act + e0 + off, bytes, &bar[k]);
    }
    const uint32_t il = (threadIdx.x % 256u) * 4u;
    for (int k = threadIdx.x / 256; k < nsl; k += THREADS / 256) {
        const uint32_t off = (uint32_t)k * SLICE, len = min((uint32_t)SLICE, cnt - off);
        mbar_wait(&bar[k], 0);
        if (il < len) {
            const uint32_t i0 = e0 + off + il;
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = *reinterpret_cast<const float4 *>(&tile[k][r][il]);
            const int4 a = *reinterpret_cast<const int4 *>(&tile[k][4][il]);
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = touch(v[r], a, r);
#pragma unroll
            for (int r = 0; r < 4; ++r) *reinterpret_cast<float4 *>(s.state + (size_t)r * n + i0) = v[r];
            *reinterpret_cast<float4 *>(s.reward + i0) = make_float4(v[0].x, v[1].y, v[2].z, v[3].w);
            *reinterpret_cast<uchar4 *>(s.done + i0) = make_uchar4(v[0].x > 0, v[0].y > 0, v[0].z > 0, v[0].w > 0);
        }
    }
}

__global__ void empty_kernel() { pdl_launch_dependents(); pdl_wait(); }

// ---- harness -------------------------------------------------------------------------------------
using Launch = std::function<void(int slot, cudaStream_t)>;

static double median(std::vector<double> v)
{
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

struct Harness {
    cudaStream_t s0, s1;
    cudaEvent_t e0, e1, fork, join;
    int K = 2000, reps = 5, ring = 16;

    // us per launch
    double two_streams(const Launch &f)
    {
        std::vector<double> t;
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s0));
            CK(cudaStreamWaitEvent(s1, e0));
            for (int i = 0; i < K; ++i) f(i % ring, (i & 1) ? s1 : s0);
            CK(cudaEventRecord(join, s1));
            CK(cudaStreamWaitEvent(s0, join));
            CK(cudaEventRecord(e1, s0));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r) t.push_back(ms * 1e3 / K);
        }
        return median(t);
    }
    double one_stream(const Launch &f, bool resident)
    {
        std::vector<double> t;
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s0));
            for (int i = 0; i < K; ++i) f(resident ? 0 : i % ring, s0);
            CK(cudaEventRecord(e1, s0));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r) t.push_back(ms * 1e3 / K);
        }
        return median(t);
    }
    // one launch at a time, drained before the next; events around the launch (includes ~1-2 us of event cost:
    // compare with the `empty` row, and with the ncu pass over --iso-only)
    double isolated(const Launch &f, int count)
    {
        std::vector<double> t;
        for (int i = 0; i < count + 8; ++i) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s0));
            f(i % ring, s0);
            CK(cudaEventRecord(e1, s0));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 8) t.push_back(ms * 1e3);
        }
        return median(t);
    }
};

template <class K, class... A>
static void launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
}

int main(int argc, char **argv)
{
    bool iso_only = false;
    std::string env_name = "cartpole";
    Harness h;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--iso-only")) iso_only = true;
        else if (!std::strcmp(argv[i], "--env") && i + 1 < argc) env_name = argv[++i];
        else if (!std::strcmp(argv[i], "--k") && i + 1 < argc) h.K = std::atoi(argv[++i]);
    }
    const uint32_t n = 1u << 20;
    const bool cartpole = env_name == "cartpole", pendulum = env_name == "pendulum";
    if (!cartpole && !pendulum && env_name != "mountain_car") { std::fprintf(stderr, "--env cartpole | mountain_car | pendulum\n"); return 2; }
    const int kind = cartpole ? GYMRS_CARTPOLE : pendulum ? GYMRS_PENDULUM : GYMRS_MOUNTAIN_CAR;
    const int state_rows = cartpole ? 4 : 2, obs_rows = pendulum ? 3 : 0;
    const double mb = (4.0 * (2 * state_rows + obs_rows + 2) + 1) * n / 1e6; // algorithmic bytes per launch
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaStreamCreateWithFlags(&h.s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h.s1, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h.e0));
    CK(cudaEventCreate(&h.e1));
    CK(cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming));

    // ring of independent batches (larger than L2 together)
    std::vector<Slot> slots(h.ring);
    std::vector<float *> obs(h.ring, nullptr);
    std::vector<int32_t> host_act(n);
    for (int i = 0; i < h.ring; ++i) {
        if (obs_rows) { CK(cudaMalloc(&obs[i], 4ull * obs_rows * n)); }
        float *st, *rw;
        int32_t *ac;
        uint8_t *dn;
        CK(cudaMalloc(&st, 16ull * n));
        CK(cudaMalloc(&ac, 4ull * n));
        CK(cudaMalloc(&rw, 4ull * n));
        CK(cudaMalloc(&dn, n));
        CK(cudaMemset(st, 0, 16ull * n));
        uint32_t x = 12345u + i;
        for (uint32_t j = 0; j < n; ++j) {
            x = x * 1664525u + 1013904223u;
            if (pendulum) { const float u = (float)((x >> 8) & 0xffff) / 65536.0f * 4.0f - 2.0f; std::memcpy(&host_act[j], &u, 4); }
            else host_act[j] = (x >> 16) % (cartpole ? 2 : 3);
        }
        CK(cudaMemcpy(ac, host_act.data(), 4ull * n, cudaMemcpyHostToDevice));
        slots[i] = Slot{st, ac, rw, dn};
    }
    // the library's handles over their own memory
    std::vector<gymrs_env *> envs(h.ring);
    for (int i = 0; i < h.ring; ++i) {
        if (gymrs_create(kind, n, 0, (uint64_t)i * n, nullptr, 0, &envs[i])) { std::fprintf(stderr, "%s\n", gymrs_last_error()); return 1; }
        uint64_t seed = 7;
        gymrs_reset(envs[i], &seed, nullptr, nullptr, nullptr, nullptr);
        gymrs_sync(envs[i], nullptr);
    }
    auto bind = [&](cudaStream_t st, int pdl) {
        for (auto *e : envs) { gymrs_set_stream(e, st); gymrs_set_launch_config(e, 0, 0, pdl); }
    };

    const unsigned grid_plain = n / 4 / 128;
    CK(cudaFuncSetAttribute(copy_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const size_t pad10 = 21 * 1024; // 10 CTAs x (21 KB + 1 KB reserved) <= 228 KB, 11 do not fit
    auto chunk_smem = [&](int max_slices) { return (size_t)max_slices * ROWS * SLICE * 4 + 8 * max_slices + 64; };
    const int sl1 = (int)(((n / 128 / sms + 1) * 128 + SLICE - 1) / SLICE), sl2 = (int)(((n / 128 / (2 * sms) + 1) * 128 + SLICE - 1) / SLICE);
    CK(cudaFuncSetAttribute(copy_chunk<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk_smem(sl1)));
    CK(cudaFuncSetAttribute(copy_chunk<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk_smem(sl2)));
    CK(cudaFuncSetAttribute(copy_chunk<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk_smem(sl1)));

    struct Variant { std::string name; Launch f; bool is_real; };
    std::vector<Variant> vs;
    vs.push_back({"empty", [&](int, cudaStream_t st) { launch_pdl(empty_kernel, dim3(1), dim3(32), 0, st, true); }, false});
    auto rows = [&](int i, cudaStream_t st, size_t smem, int pf) {
        if (pendulum) launch_pdl(copy_rows<2, 3>, dim3(grid_plain), dim3(128), smem, st, true, slots[i], obs[i], n, pf);
        else launch_pdl(copy_rows<2, 0>, dim3(grid_plain), dim3(128), smem, st, true, slots[i], obs[i], n, pf);
    };
    if (!cartpole) {
        CK(cudaFuncSetAttribute(copy_rows<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(copy_rows<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        vs.push_back({"plain16", [&](int i, cudaStream_t st) { rows(i, st, 0, 1); }, false});
        vs.push_back({"plain16-nopf", [&](int i, cudaStream_t st) { rows(i, st, 0, 0); }, false});
        vs.push_back({"plain10", [&](int i, cudaStream_t st) { rows(i, st, pad10, 1); }, false});
    }
    if (cartpole) vs.push_back({"plain16", [&](int i, cudaStream_t st) { launch_pdl(copy_plain, dim3(grid_plain), dim3(128), 0, st, true, slots[i], n, 1); }, false});
    if (cartpole) vs.push_back({"plain16-nopf", [&](int i, cudaStream_t st) { launch_pdl(copy_plain, dim3(grid_plain), dim3(128), 0, st, true, slots[i], n, 0); }, false});
    if (cartpole) vs.push_back({"plain10", [&](int i, cudaStream_t st) { launch_pdl(copy_plain, dim3(grid_plain), dim3(128), pad10, st, true, slots[i], n, 1); }, false});
    if (cartpole) vs.push_back({"chunk 1x1024", [&](int i, cudaStream_t st) { launch_pdl(copy_chunk<1024>, dim3(sms), dim3(1024), chunk_smem(sl1), st, true, slots[i], n, sl1, 0); }, false});
    if (cartpole) vs.push_back({"chunk 1x768", [&](int i, cudaStream_t st) { launch_pdl(copy_chunk<768>, dim3(sms), dim3(768), chunk_smem(sl1), st, true, slots[i], n, sl1, 0); }, false});
    if (cartpole) vs.push_back({"chunk 2x512", [&](int i, cudaStream_t st) { launch_pdl(copy_chunk<512>, dim3(2 * sms), dim3(512), chunk_smem(sl2), st, true, slots[i], n, sl2, 0); }, false});
    vs.push_back({"real step", [&](int i, cudaStream_t) { gymrs_step(envs[i], slots[i].act, GYMRS_STEP_AUTORESET); }, true});

    std::printf("%-14s %10s %12s %12s %10s   (%s, us per 1M-env launch; %.1f MB algorithmic per launch)\n", "variant", "2 streams",
                "1 str cold", "1 str L2-res", "isolated", env_name.c_str(), mb);
    for (auto &v : vs) {
        double a = 0, b = 0, c = 0, d = 0;
        if (v.is_real) bind(h.s0, 1);
        if (!iso_only) {
            if (v.is_real) {
                // handles are bound to a stream: alternate them over the two streams like bench.py does
                for (int i = 0; i < h.ring; ++i) { gymrs_set_stream(envs[i], (i & 1) ? h.s1 : h.s0); }
                a = h.two_streams(v.f);
                bind(h.s0, 1);
            } else {
                a = h.two_streams(v.f);
            }
            b = h.one_stream(v.f, false);
            c = h.one_stream(v.f, true);
        }
        d = h.isolated(v.f, iso_only ? 40 : 200);
        std::printf("%-14s %10.2f %12.2f %12.2f %10.2f\n", v.name.c_str(), a, b, c, d);
        std::fflush(stdout);
    }
    for (auto *e : envs) {
        uint64_t bad = 0;
        if (gymrs_sync(e, &bad)) std::fprintf(stderr, "library error: %s\n", gymrs_last_error());
        gymrs_destroy(e);
    }
    return 0;
}
