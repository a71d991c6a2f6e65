// Probe 2 (round 2): is the 93-cycle floor of a small-N tcgen05.mma (mma_rate_probe.cu) a property of the tensor
// pipe or of the issuing thread / CTA shape?  Measures, on resident zeroed operands, one CTA (or CTA pair) per SM:
//   A  cta_group::1, M = 128, one issuer thread                     (baseline)
//   B  cta_group::1, M = 128, TWO issuer threads (warps 0 and 1), one accumulator each
//   C  cta_group::1, M = 64
//   D  cta_group::2, M = 256 (128 rows per CTA), B split N/2 per CTA, issued by the leader CTA
//   E  cta_group::2, M = 128 (64 rows per CTA)
//   F  as A, issued by an elected lane of a warp-uniform loop
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mma_rate_probe2 mma_rate_probe2.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../aanet_b200/csrc/umma.cuh"

namespace cg = cooperative_groups;
using namespace aanet::umma;

__device__ __forceinline__ void mma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// mode F: the issuing warp runs warp-uniform code and one ELECTED lane issues (elect.sync): no compiler-generated
// ELECT / BRA.U.ANY retry loop around every UTCHMMA (what `if (tid == 0)` produces)
__global__ void __launch_bounds__(128)
rate_elect_kernel(long long *out, int M, int N, int iters) {
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
    if (warp == 0) {
        const uint64_t a = make_desc_sw128(smem_u32(smem)), b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = make_idesc_tf32(M, N);
        const uint32_t d = s_tmem;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_tf32(d, a + 2 * k, b + 2 * k, id, 1);
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

// mode T: as F with the A operand in TENSOR MEMORY (the TMEM-A kernels' form), single N and the kernels' alternating
// N = 2 BN / N = BN pair
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128)
rate_ts_kernel(long long *out, int N, int N2, int iters) {
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
    if (warp == 0) {
        const uint64_t b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = make_idesc_tf32(128, N), id2 = make_idesc_tf32(128, N2);
        const uint32_t d = s_tmem, a = s_tmem + 256;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_tf32_ts(d, a + k * 8, b + 2 * k, id, 1);
                    if (N2) mma_tf32_ts(d, a + 32 + k * 8, b + 2 * k, id2, 1);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

// mode Q: depth of the tensor pipe's instruction queue -- cycles the issuing lane needs to get n MMAs (N = 256, 128
// cycles of pipe time each) accepted, starting from an idle pipe
__global__ void __launch_bounds__(128)
queue_kernel(long long *out, int n) {
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
    if (warp == 0) {
        const uint64_t a = make_desc_sw128(smem_u32(smem)), b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = make_idesc_tf32(128, 256);
        const uint32_t d = s_tmem;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int k = 0; k < n; ++k) mma_tf32(d, a, b, id, 1);
            t1 = clock64();
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && t1 != 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

// modes A (n_issuers = 1), B (n_issuers = 2), C (M = 64)
__global__ void __launch_bounds__(128)
rate1_kernel(long long *out, int M, int N, int iters, int n_issuers) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ long long s_t[2][2];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&s_tmem);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if ((tid & 31) == 0 && warp < n_issuers) {
        const uint64_t a = make_desc_sw128(smem_u32(smem)), b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = make_idesc_tf32(M, N);
        const uint32_t d = s_tmem + warp * 256;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32(d, a + 2 * k, b + 2 * k, id, 1);
        }
        const long long t1 = clock64();
        tc_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0);
        const long long t2 = clock64();
        s_t[warp][0] = t1 - t0; s_t[warp][1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) {
        out[0] = s_t[0][0]; out[1] = s_t[0][1];
        if (n_issuers > 1) { out[0] = max(out[0], s_t[1][0]); out[1] = max(out[1], s_t[1][1]); }
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

// modes D / E: CTA pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
rate2_kernel(long long *out, int M, int N, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned rank = cluster.block_rank();
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    cluster.sync();
    tc_fence_after();
    if (tid == 0 && rank == 0) {
        const uint64_t a = make_desc_sw128(smem_u32(smem)), b = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = make_idesc_tf32(M, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_tf32_2cta(s_tmem, a + 2 * k, b + 2 * k, id, 1);
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    cluster.sync();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512) : "memory");
    }
}

static int report(const char *what, int M, int N, int grid, long long *d_out, int iters, int rows_per_sm) {
    long long h[2];
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s M=%d N=%d: error %s\n", what, M, N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    const double per = (double)h[1] / (iters * 4.0);
    printf("%-44s M=%3d N=%3d grid=%3d: issue %.1f, complete %.1f cyc/MMA -> %.0f MAC/clk/SM\n", what, M, N, grid,
           (double)h[0] / (iters * 4.0), per, (double)rows_per_sm * N * 8 / per);
    return 0;
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(rate1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    const int iters = 2000;
    for (int N : {32, 64, 128, 256})
        for (int grid : {1, 148}) {
            rate1_kernel<<<grid, 128, 52 * 1024>>>(d_out, 128, N, iters, 1);
            if (report("A cta_group::1, one issuer", 128, N, grid, d_out, iters, 128)) return 1;
        }
    cudaFuncSetAttribute(rate_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    for (int N : {32, 64, 128, 192}) {
        long long h[2];
        rate_ts_kernel<<<148, 128, 52 * 1024>>>(d_out, N, 0, iters);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("ts probe failed\n"); return 1; }
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("T A in tensor memory, elected lane                M=128 N=%3d: %.1f cyc/MMA\n", N, (double)h[1] / (iters * 4.0));
    }
    for (int N : {32, 64}) {
        long long h[2];
        rate_ts_kernel<<<148, 128, 52 * 1024>>>(d_out, 2 * N, N, iters);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("ts probe failed\n"); return 1; }
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("T A in tensor memory, the kernels' K step (N=%3d then N=%3d): %.1f cycles per K step of 8\n", 2 * N, N, (double)h[1] / (iters * 4.0));
    }
    cudaFuncSetAttribute(queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    for (int n : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32}) {
        long long h[2];
        queue_kernel<<<1, 128, 52 * 1024>>>(d_out, n);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("queue probe failed\n"); return 1; }
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("Q issue of %2d MMAs (N = 256, 128 cyc each) from an idle pipe: %5lld cycles until accepted, %5lld until retired\n", n, h[0], h[1]);
    }
    cudaFuncSetAttribute(rate_elect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    for (int N : {32, 64, 96, 128, 192, 256}) {
        rate_elect_kernel<<<148, 128, 52 * 1024>>>(d_out, 128, N, iters);
        if (report("F cta_group::1, elected lane of a uniform warp", 128, N, 148, d_out, iters, 128)) return 1;
    }
    for (int N : {32, 64, 128, 256}) {
        rate1_kernel<<<148, 128, 52 * 1024>>>(d_out, 128, N, iters, 2);
        // two issuers: per-issuer cycles per MMA; the SM retires 2 MMAs in that time
        if (report("B cta_group::1, two issuers (x2 MMAs)", 128, N, 148, d_out, iters, 256)) return 1;
    }
    for (int N : {32, 64, 128, 256}) {
        rate1_kernel<<<148, 128, 52 * 1024>>>(d_out, 64, N, iters, 1);
        if (report("C cta_group::1, M = 64", 64, N, 148, d_out, iters, 64)) return 1;
    }
    for (int N : {32, 64, 128, 256}) {
        rate2_kernel<<<148, 128, 52 * 1024>>>(d_out, 256, N, iters);
        if (report("D cta_group::2, M = 256 (128 rows per SM)", 256, N, 148, d_out, iters, 128)) return 1;
    }
    for (int N : {32, 64, 128, 256}) {
        rate2_kernel<<<148, 128, 52 * 1024>>>(d_out, 128, N, iters);
        if (report("E cta_group::2, M = 128 (64 rows per SM)", 128, N, 148, d_out, iters, 64)) return 1;
    }
    return 0;
}
