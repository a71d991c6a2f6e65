// Probe (round 2): issue/execution rate of tcgen05.mma kind::tf32 (M = 128, K = 8) and kind::f16 (bf16, K = 16) as
// a function of N, back to back from one thread on resident shared-memory operands, one CTA per SM.  Calibrates the
// tensor floor of the 3xTF32 convolution kernels (conv_umma_kernel.cuh, halo_engine.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mma_rate_probe mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../aanet_b200/csrc/umma.cuh"

using namespace aanet::umma;

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: tf32, one accumulator; 1: tf32, two accumulators alternating; 2: bf16 one accumulator
__global__ void __launch_bounds__(128)
rate_kernel(long long *out, int N, int iters, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&s_tmem);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint64_t a = make_desc_sw128(smem_u32(smem)), b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = mode == 2 ? idesc_bf16(128, N) : make_idesc_tf32(128, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = s_tmem + ((mode == 1 && (k & 1)) ? 256 : 0);
                if (mode == 2) mma_bf16(d, a + 2 * k, b + 2 * k, id, 1);
                else mma_tf32(d, a + 2 * k, b + 2 * k, id, 1);
            }
        }
        const long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

int main() {
    long long *d_out, h[2];
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    const int iters = 2000;
    const char *names[3] = {"tf32 K=8, one accumulator", "tf32 K=8, two accumulators", "bf16 K=16, one accumulator"};
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {32, 64, 128, 192, 256}) {
            for (int grid : {1, 148}) {
                rate_kernel<<<grid, 128, 52 * 1024>>>(d_out, N, iters, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
                const double per = (double)h[1] / (iters * 4.0);
                printf("%-28s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f MAC/clk/SM\n", names[mode], N,
                       grid, (double)h[0] / (iters * 4.0), per, 128.0 * N * (mode == 2 ? 16 : 8) / per);
            }
        }
    return 0;
}
