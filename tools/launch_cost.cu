// Host cost of one launch: bare cudaLaunchKernelEx (empty kernel, same parameter size as step_kernel,
// with / without the PDL attribute) vs gymrs_step through the C ABI.  Build here, run on the GPU box:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/bin/launch_cost tools/launch_cost.cu \
//        -Iinclude -Lgym_rs_b200 -lgymrs_b200 -Xlinker -rpath='$ORIGIN/../../gym_rs_b200'
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include "gymrs_b200.h"

struct Blob { char bytes[416]; };
__global__ void empty_kernel(const __grid_constant__ Blob b) { if (b.bytes[0] == 77 && threadIdx.x == 999) printf("x"); }

template <class F> static double per_call_us(int k, F f)
{
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < k; ++i) f();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / k;
}

int main()
{
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    Blob b = {};
    for (int pdl = 0; pdl <= 1; ++pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(128); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl;
        auto f = [&] { cudaLaunchKernelEx(&cfg, empty_kernel, b); };
        per_call_us(500, f); cudaStreamSynchronize(s);
        double us = per_call_us(400, f);
        cudaStreamSynchronize(s);
        std::printf("bare cudaLaunchKernelEx, 416 B params, pdl attr=%d: %5.2f us/call\n", pdl, us);
    }
    {
        cudaStreamCaptureStatus st;
        double us = per_call_us(100000, [&] { cudaStreamIsCapturing(s, &st); });
        int d;
        double us2 = per_call_us(100000, [&] { cudaGetDevice(&d); });
        std::printf("cudaStreamIsCapturing %5.3f us, cudaGetDevice %5.3f us\n", us, us2);
    }
    for (uint64_t n : {1024ull, 1ull << 20}) {
        gymrs_env *e = nullptr;
        if (gymrs_create(GYMRS_CARTPOLE, n, 0, 0, nullptr, 0, &e)) { std::printf("%s\n", gymrs_last_error()); return 1; }
        uint64_t seed = 0;
        gymrs_reset(e, &seed, nullptr, nullptr, nullptr, nullptr);
        int32_t *act;
        cudaMalloc(&act, 4 * n);
        cudaMemset(act, 0, 4 * n);
        for (int pdl = 0; pdl <= 2; ++pdl) {
            gymrs_set_launch_config(e, 0, 0, pdl);
            for (int i = 0; i < 200; ++i) gymrs_step(e, act, GYMRS_STEP_AUTORESET);
            gymrs_sync(e, nullptr);
            const int k = n == 1024 ? 400 : 4000;
            auto t0 = std::chrono::steady_clock::now();
            double us = per_call_us(k, [&] { gymrs_step(e, act, GYMRS_STEP_AUTORESET); });
            gymrs_sync(e, nullptr);
            auto t2 = std::chrono::steady_clock::now();
            std::printf("gymrs_step n=%8llu pdl=%d: host %5.2f us/call, until drained %5.2f us/step\n", (unsigned long long)n, pdl,
                        us, std::chrono::duration<double, std::micro>(t2 - t0).count() / k);
        }
        gymrs_destroy(e);
        cudaFree(act);
    }
    return 0;
}
