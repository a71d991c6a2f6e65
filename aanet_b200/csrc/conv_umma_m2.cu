// Engine instantiations for MODE 2 (DEFORM, single-run K blocks); see conv_umma_kernel.cuh.
#include "conv_umma_kernel.cuh"

namespace aanet {
AANET_DEFINE_CONV_MODE(2)
AANET_DEFINE_PROFILE_READ(2)
}  // namespace aanet
