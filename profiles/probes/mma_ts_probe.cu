// Probe (round 2): tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st, thread = row) --
// (1) correctness against an identity B, (2) cycles per MMA vs N, to compare with the shared-memory-A numbers of
// mma_rate_probe.cu (104 / 93 cycles at N = 128 / <= 64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mma_ts_probe mma_ts_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../aanet_b200/csrc/umma.cuh"

using namespace aanet::umma;

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// A[r][k] (128 x 32) lives in TMEM columns [256, 288); B = 32 x 32 identity (K-major SW128) -> D[r][n] = A[r][n].
__global__ void __launch_bounds__(128)
ts_check(float *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *bt = reinterpret_cast<float *>(smem);
    for (int i = tid; i < 32 * 32; i += 128) {
        const int n = i >> 5, k = i & 31;
        bt[n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3))] = (n == k) ? 1.f : 0.f;
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&s_tmem);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = s_tmem;
    float v[16];
    for (int h = 0; h < 2; ++h) {
        for (int i = 0; i < 16; ++i) v[i] = (float)((tid * 3 + (h * 16 + i) * 5) % 1000 + 1);
        tmem_st16(tb + ((uint32_t)(warp * 32) << 16) + 256 + h * 16, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint64_t b = make_desc_sw128(smem_u32(bt));
        for (int k = 0; k < 4; ++k) mma_tf32_ts(tb, tb + 256 + k * 8, b + 2 * k, make_idesc_tf32(128, 32), k != 0);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int h = 0; h < 2; ++h) {
        tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + h * 16, v);
        for (int i = 0; i < 16; ++i) out[tid * 32 + h * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
    (void)lane;
}

__global__ void __launch_bounds__(128)
ts_rate(long long *out, int N, int iters, int ts) {
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
        const uint32_t id = make_idesc_tf32(128, N), tb = s_tmem;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ts) mma_tf32_ts(tb, tb + 256 + 8 * k, b + 2 * k, id, 1);
                else mma_tf32(tb, a + 2 * k, b + 2 * k, id, 1);
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(s_tmem); }
}

int main() {
    float *d_out;
    long long *d_t, ht;
    cudaMalloc(&d_out, 128 * 32 * 4);
    cudaMalloc(&d_t, 8);
    cudaFuncSetAttribute(ts_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    ts_check<<<1, 128, 8192>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts_check: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 32];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 32; ++c) if (h[r * 32 + c] != (float)((r * 3 + c * 5) % 1000 + 1)) ++bad;
    printf("tcgen05.st (32x32b, thread = row) + TMEM-A MMA: %d / 4096 mismatches\n", bad);
    const int iters = 2000;
    for (int ts = 0; ts < 2; ++ts)
        for (int N : {16, 32, 64, 128, 192, 256}) {
            ts_rate<<<148, 128, 52 * 1024>>>(d_t, N, iters, ts);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ts_rate: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&ht, d_t, 8, cudaMemcpyDeviceToHost);
            printf("A in %s, tf32 M=128 K=8 N=%3d: %.1f cycles per MMA\n", ts ? "TMEM" : "smem", N, (double)ht / (iters * 4.0));
        }
    return 0;
}
