// kernels_cartpole.cu -- CartPole instantiation of the step / rollout / reset kernels.
#include "kernels_impl.cuh"
namespace gymrs {
GYMRS_INSTANTIATE(CartPole)
} // namespace gymrs
