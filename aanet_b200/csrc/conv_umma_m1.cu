// Engine instantiations for MODE 1 (DEFORM, general); see conv_umma_kernel.cuh.
#include "conv_umma_kernel.cuh"

namespace aanet {
AANET_DEFINE_CONV_MODE(1)
AANET_DEFINE_PROFILE_READ(1)
}  // namespace aanet
