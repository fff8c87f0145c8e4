// curand_check.cu -- test helper (not product code): prints what cuRAND's own Philox4x32-10 device
// generator returns for the (seed, global env id, epoch) triples on stdin, so the tests can show
// that the reset stream of this repo (csrc/philox.cuh, restated in oracle/) IS "in-kernel curand":
//     curand_init(seed, /*subsequence*/ epoch, /*offset*/ 4 * gid, &state);  curand4(&state)
// == Philox4x32-10(key = seed, counter = (gid, epoch)).
// Build (done by __graft_entry__.build()):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/curand_check tests/cuda/curand_check.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <curand_kernel.h>

struct Triple { unsigned long long seed, gid, epoch; };

__global__ void draw(const Triple *t, uint4 *out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(t[i].seed, t[i].epoch, 4ull * t[i].gid, &st);
    out[i] = curand4(&st);
}

int main()
{
    std::vector<Triple> in;
    Triple t;
    while (std::scanf("%llu %llu %llu", &t.seed, &t.gid, &t.epoch) == 3) in.push_back(t);
    const int n = (int)in.size();
    if (n == 0) return 0;
    Triple *d_in;
    uint4 *d_out;
    if (cudaMalloc(&d_in, n * sizeof(Triple)) != cudaSuccess) { std::fprintf(stderr, "no CUDA device\n"); return 2; }
    cudaMalloc(&d_out, n * sizeof(uint4));
    cudaMemcpy(d_in, in.data(), n * sizeof(Triple), cudaMemcpyHostToDevice);
    draw<<<(n + 127) / 128, 128>>>(d_in, d_out, n);
    std::vector<uint4> out(n);
    if (cudaMemcpy(out.data(), d_out, n * sizeof(uint4), cudaMemcpyDeviceToHost) != cudaSuccess) return 3;
    for (const uint4 &w : out) std::printf("%u %u %u %u\n", w.x, w.y, w.z, w.w);
    return 0;
}
