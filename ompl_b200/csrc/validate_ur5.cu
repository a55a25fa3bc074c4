// Validation kernels of the ur5 robot (explicit instantiation of validate_launch.cuh).
#include "validate_launch.cuh"
#include "robots/ur5_gen.cuh"

namespace b200mp {
B200MP_INSTANTIATE_ROBOT(Ur5)
}
