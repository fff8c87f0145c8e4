// Host cost of gymrs_step from C (no Python / ctypes): nvcc tools/launch_cost.cu -Iinclude -Lgym_rs_b200 -lgymrs_b200
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include "gymrs_b200.h"
int main()
{
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
            for (int i = 0; i < k; ++i) gymrs_step(e, act, GYMRS_STEP_AUTORESET);
            auto t1 = std::chrono::steady_clock::now();
            gymrs_sync(e, nullptr);
            auto t2 = std::chrono::steady_clock::now();
            std::printf("C  n=%8llu pdl=%d: host %5.2f us/call, until drained %5.2f us/step\n", (unsigned long long)n, pdl,
                        std::chrono::duration<double, std::micro>(t1 - t0).count() / k,
                        std::chrono::duration<double, std::micro>(t2 - t0).count() / k);
        }
        gymrs_destroy(e);
        cudaFree(act);
    }
    return 0;
}
