// Status strings / version for the C-ABI (include/aanet_b200.h).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace aanet {
static thread_local char g_last_err[256] = "";
void set_last_cuda_error(const char *msg) {
    strncpy(g_last_err, msg ? msg : "", sizeof(g_last_err) - 1);
    g_last_err[sizeof(g_last_err) - 1] = 0;
}
int num_sms() {
    static int cache[64] = {0};          // benign race: every thread writes the same value
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
        cache[dev] = n;
    }
    return cache[dev];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_f32(CUtensorMap *tm, const void *base, int rank, const uint64_t *dims,
                        const uint64_t *strides_bytes, const uint32_t *box, int swizzle) {
    static EncodeTiledFn encode = nullptr;      // benign race: every thread resolves the same entry point
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            set_last_cuda_error("cuTensorMapEncodeTiled is not available");
            return AANET_ERR_LAUNCH;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), d, s, b,
                              e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_cuda_error("cuTensorMapEncodeTiled failed");
        return AANET_ERR_LAUNCH;
    }
    return AANET_OK;
}

bool pdl_enabled() {
    static const bool on = getenv("AANET_NO_PDL") == nullptr;
    return on;
}
}  // namespace aanet

extern "C" int aanet_abi_version(void) { return AANET_B200_ABI_VERSION; }

extern "C" const char *aanet_status_string(int s) {
    switch (s) {
        case AANET_OK: return "ok";
        case AANET_ERR_NULL: return "required pointer is NULL";
        case AANET_ERR_SHAPE: return "invalid or inconsistent shape";
        case AANET_ERR_UNSUPPORTED: return "configuration not supported by the sm_100a kernels";
        case AANET_ERR_WORKSPACE: return "workspace missing or too small";
        case AANET_ERR_LAUNCH: return "CUDA launch failed";
        default: return "unknown status";
    }
}

extern "C" const char *aanet_last_cuda_error(void) { return aanet::g_last_err; }
