// Shared helpers for the aanet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/aanet_b200.h"

namespace aanet {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs (sizing of static tables only)
// SM count of the current device (cached per device ordinal): grids and split heuristics use this, so a MIG
// slice or another SKU only changes the numbers, never correctness.
int num_sms();

// Records the CUDA error text of a failed launch for aanet_last_cuda_error().
void set_last_cuda_error(const char *msg);

// Checks cudaGetLastError() after a launch; never synchronises.
inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_cuda_error(cudaGetErrorString(e));
        return AANET_ERR_LAUNCH;
    }
    return AANET_OK;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch.  A kernel launched through launch_pdl may be scheduled while the previous
// kernel of the stream is still running (after that kernel's CTAs have all called pdl_trigger() or exited); it
// must call pdl_wait() before touching global memory.  AANET_NO_PDL=1 turns the attribute off.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    return check_launch();
}

template <typename T>
__host__ __device__ inline bool aligned16(const T *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Streaming (read-once) 128-bit and 32-bit loads that do not pollute L1.
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

}  // namespace aanet
