// kernels_mountain_car.cu -- MountainCar instantiation of the step / rollout / reset kernels.
#include "kernels_impl.cuh"
namespace gymrs {
GYMRS_INSTANTIATE(MountainCar)
} // namespace gymrs
