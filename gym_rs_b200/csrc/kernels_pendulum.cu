// kernels_pendulum.cu -- Pendulum instantiation of the step / rollout / reset kernels, plus the
// kernel that rebuilds the observation rows (cos, sin, theta_dot) after gymrs_set_state.
#include "kernels_impl.cuh"
namespace gymrs {
GYMRS_INSTANTIATE(Pendulum)

namespace {
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
} // namespace

cudaError_t launch_pendulum_obs(const BatchArgs &a, cudaStream_t s)
{
    if (a.n == 0) return cudaSuccess;
    pendulum_obs_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}
} // namespace gymrs
