// Engine instantiations for MODE 3 (DENSE, 1-D tiles only (1x1 convolutions)); see conv_umma_kernel.cuh.
#include "conv_umma_kernel.cuh"

namespace aanet {
AANET_DEFINE_CONV_MODE(3)
AANET_DEFINE_PROFILE_READ(3)
}  // namespace aanet
