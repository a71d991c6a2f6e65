// Engine instantiations for MODE 0 (DENSE, 2-D tiles possible); see conv_umma_kernel.cuh.
#include "conv_umma_kernel.cuh"

namespace aanet {
AANET_DEFINE_CONV_MODE(0)
AANET_DEFINE_PROFILE_READ(0)
}  // namespace aanet
