// Status strings / version for the C-ABI (include/aanet_b200.h).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

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
