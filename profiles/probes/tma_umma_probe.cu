// Probe (round 2): the two hardware behaviours the halo-staged kernels rely on, checked on the B200 itself.
//   1. cp.async.bulk.tensor.3d (tensor-map TMA) of a channels-last halo box with out-of-image coordinates:
//      zero fill, element order in shared memory (plain and SWIZZLE_128B).
//   2. tcgen05.mma reading its A operand through a descriptor that is a WINDOW into a SWIZZLE_128B halo buffer:
//      start address shifted by whole 128-byte lines (not 1024-byte aligned), stride-byte-offset = halo pitch.
//      For each shift the matrix-descriptor "base offset" field is tried as 0 and as (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_umma_probe tma_umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../aanet_b200/csrc/umma.cuh"

using namespace aanet::umma;

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__global__ void tma_probe(const __grid_constant__ CUtensorMap tm, float *out, int c0, int x0, int y0, int n) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, n * 4);
        tma_load_3d(smem, &tm, c0, x0, y0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// A = window into the swizzled halo: row r -> line start_line + (r / 8) * pitch + (r % 8); B = 32 x 32 identity.
__global__ void __launch_bounds__(128)
window_probe(const __grid_constant__ CUtensorMap tm, float *out, int box_lines, int start_line, int pitch,
             int use_base_off) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar, bar_mma;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    float *halo = reinterpret_cast<float *>(smem);
    float *bt = reinterpret_cast<float *>(smem + ((box_lines * 128 + 1023) & ~1023));
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<32>(&s_tmem);
    for (int i = tid; i < 32 * 32; i += 128) {
        const int n = i >> 5, k = i & 31;
        bt[n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3))] = (n == k) ? 1.f : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(&bar, box_lines * 128);
        tma_load_3d(smem, &tm, 0, 0, 0, &bar);
        mbar_wait(&bar, 0);
        const uint32_t a0 = smem_u32(halo) + start_line * 128;
        const uint64_t a = desc_sw128(a0, pitch * 128, use_base_off ? (a0 >> 7) & 7 : 0);
        const uint64_t b = desc_sw128(smem_u32(bt), 1024, 0);
        const uint32_t idesc = make_idesc_tf32(128, 32);
        for (int k = 0; k < 4; ++k) mma_tf32(s_tmem, a + ((k * 32) >> 4), b + ((k * 32) >> 4), idesc, k != 0);
        tc_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    float v[16];
    for (int h = 0; h < 2; ++h) {
        tmem_ld16(s_tmem + ((uint32_t)(warp * 32) << 16) + h * 16, v);
        for (int i = 0; i < 16; ++i) out[tid * 32 + h * 16 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(s_tmem); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int H = 20, W = 30, C = 64;
    std::vector<float> h((size_t)H * W * C);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
        h[((size_t)y * W + x) * C + c] = (float)((y * 7 + x * 3 + c) % 1000 + 1);
    float *d_x, *d_out;
    CK(cudaMalloc(&d_x, h.size() * 4));
    CK(cudaMemcpy(d_x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, 1 << 20));
    std::vector<float> o(1 << 18);
    int bad_total = 0;

    for (int swz = 0; swz < 2; ++swz) {
        const int HB = 9, WB = 11, CB = 32;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
        cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4};
        cuuint32_t box[3] = {CB, WB, HB}, es[3] = {1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_x, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int n = HB * WB * CB;
        for (int t = 0; t < 3; ++t) {
            const int c0 = t == 1 ? 32 : 0, x0 = t == 0 ? -3 : (t == 1 ? 25 : 4), y0 = t == 0 ? -2 : (t == 1 ? 15 : 3);
            CK(cudaMemset(d_out, 0xff, n * 4));
            tma_probe<<<1, 128, n * 4 + 1024>>>(tm, d_out, c0, x0, y0, n);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(o.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int y = 0; y < HB; ++y) for (int x = 0; x < WB; ++x) for (int c = 0; c < CB; ++c) {
                const int gy = y0 + y, gx = x0 + x, gc = c0 + c;
                const float want = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[((size_t)gy * W + gx) * C + gc] : 0.f;
                const int line = y * WB + x;
                const int idx = swz ? line * 32 + ((((c >> 2) ^ (line & 7)) << 2) | (c & 3)) : line * 32 + c;
                if (o[idx] != want) ++bad;
            }
            printf("TMA 3-D box swizzle=%s origin(c=%d,x=%d,y=%d): %d / %d mismatches\n", swz ? "128B" : "none", c0, x0,
                   y0, bad, n);
            bad_total += bad;
        }
    }

    // window probe: halo box = whole-width rows of the image (pitch = WB lines), SWIZZLE_128B
    {
        const int WB = 13, HB = 20, CB = 32;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
        cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4};
        cuuint32_t box[3] = {CB, WB, HB}, es[3] = {1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_x, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int lines = WB * HB;
        const size_t smem = ((lines * 128 + 1023) & ~1023) + 4096 + 1024;
        CK(cudaFuncSetAttribute(window_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int shifts[6][2] = {{0, 0}, {0, 2}, {1, 0}, {2, 4}, {0, 5}, {3, 3}};   // (wy, wx): start = wy*pitch + wx
        for (int s = 0; s < 6; ++s) for (int bo = 0; bo < 2; ++bo) {
            const int start = shifts[s][0] * WB + shifts[s][1];
            CK(cudaMemset(d_out, 0xff, 128 * 32 * 4));
            window_probe<<<1, 128, smem>>>(tm, d_out, lines, start, WB, bo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("window start=%d base_off=%d: %s\n", start, bo, cudaGetErrorString(e)); return 1; }
            CK(cudaMemcpy(o.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int rr = 0; rr < 128; ++rr) for (int c = 0; c < 32; ++c) {
                const int line = start + (rr / 8) * WB + (rr % 8);
                const int y = line / WB, x = line % WB;
                const float want = h[((size_t)y * W + x) * C + c];
                if (o[rr * 32 + c] != want) ++bad;
            }
            printf("UMMA window: start line %3d (addr %% 1024 = %4d), pitch %d lines, base_offset %s: %d / 4096 mismatches\n",
                   start, (start * 128) % 1024, WB, bo ? "(addr>>7)&7" : "0", bad);
        }
    }
    printf("tma box mismatches total: %d\n", bad_total);
    return 0;
}
