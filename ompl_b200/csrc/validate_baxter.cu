// Validation kernels of the baxter robot (explicit instantiation of validate_launch.cuh).
#include "validate_launch.cuh"
#include "robots/baxter_gen.cuh"

namespace b200mp {
B200MP_INSTANTIATE_ROBOT(Baxter)
}
