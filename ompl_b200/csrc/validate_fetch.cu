// Validation kernels of the fetch robot (explicit instantiation of validate_launch.cuh).
#include "validate_launch.cuh"
#include "robots/fetch_gen.cuh"

namespace b200mp {
B200MP_INSTANTIATE_ROBOT(Fetch)
}
