// Validation kernels of the panda robot (explicit instantiation of validate_launch.cuh).
#include "validate_launch.cuh"
#include "robots/panda_gen.cuh"

namespace b200mp {
B200MP_INSTANTIATE_ROBOT(Panda)
}
