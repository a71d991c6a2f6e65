// Correlation cost volume (reference nets/cost.py:40-48), round-2 kernel: TMA-staged feature tiles feeding tcgen05.
//
//     cost[b,d,h,w] = (1/C) * sum_c L[b,c,h,w] * R[b,c,h,w-d]   (0 for w < d)
// Per image row the band is a slice of a GEMM, G[w, w'] = sum_c L[c,w] R[c,w'], cost[d,w] = G[w, w-d] (see
// correlation_umma.cu).  The round-1 kernel staged its operands with LDG.32 + a register transposition + STS into
// K-major tiles, single-buffered, with one __syncthreads between store and MMA (31.7 us at the 1/3 scale of config 2
// against a 10.6 us HBM floor).  Here:
//   * the NCHW rows are contiguous along w, which is exactly an MN-major operand: cp.async.bulk.tensor (4-D tensor
//     map {W, H, C, B}, box 32 w x 32 c, SWIZZLE_128B_ATOM_32B -- the one layout tcgen05 takes for MN-major tf32) drops [32 channels][32 w] boxes into shared memory and the
//     tensor core reads them through MN-major descriptors -- no transposition, no staging through registers;
//   * kind::tf32 ignores the 13 low mantissa bits, so the raw fp32 box IS the hi operand; converter warps write only
//     lo = x - trunc_tf32(x) next to it (3xTF32: raw x raw + raw x lo + lo x raw, the parity bar is 1e-4);
//   * four operand stages of 16 channels with full/empty mbarriers and two TMEM accumulators: the loads of the next
//     K blocks run under the MMAs of the current one, the epilogue of tile i under the main loop of tile i+1; persistent CTAs walk (b, h, 128 w)
//     tiles.
// Roles (320 threads): warps 0-3 converters, 4-7 epilogue (TMEM lane quarters), 8 TMA, 9 MMA issuer.
// Supported: D <= 64, W % 4 == 0, 16-byte aligned tensors; everything else takes correlation_umma.cu.
#include <stdlib.h>
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kCtM = 128;            // w per tile (UMMA M)
constexpr int kCtThreads = 320;
constexpr int kCtKC = 16;            // channels per pipeline stage (two k steps of 8)
constexpr int kCtStages = 4;         // the loads are latency-bound (HBM round trip ~ 2 stages of MMA time): 4 in flight
constexpr int kCtBox = kCtKC * 32 * 4;   // one TMA box: 16 channels x 32 w x 4 B

struct CorrTmaParams {
    float *cost;
    int C, H, W, D, B;
    int tiles_w, total_tiles, KB;
};

// N = 128 + Dp columns w' in [w0 - Dp, w0 + 128); Dp in {32, 64} (window margin >= D).
template <int N, bool NHWC>
__global__ void __launch_bounds__(kCtThreads, 1)
corr_tma_kernel(const __grid_constant__ CorrTmaParams cp, const __grid_constant__ CUtensorMap tmL,
                const __grid_constant__ CUtensorMap tmR) {
    constexpr int Dp = N - kCtM, NB = N / 32;
    constexpr int kRawBytes = (4 + NB) * kCtBox;          // A boxes then B boxes
    constexpr int kStageBytes = 2 * kRawBytes;            // raw, then lo at + kRawBytes
    constexpr int kPitch = 32 + Dp + 1;                   // epilogue staging row pitch in floats (odd: conflict-free writes)
    extern __shared__ uint8_t smem_raw[];
    constexpr int S = kCtStages;
    __shared__ __align__(8) uint64_t bar_raw[S], bar_lo[S], bar_empty[S], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);
    float *stage_mem = reinterpret_cast<float *>(smem + S * kStageBytes);      // [128][kPitch]
    const int KB = cp.KB, total = cp.total_tiles;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            umma::mbar_init(&bar_raw[s], 1);              // expect_tx + TMA bytes
            umma::mbar_init(&bar_lo[s], 4);               // converter warps
            umma::mbar_init(&bar_empty[s], 1);            // tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&bar_acc_full[s], 1);
            umma::mbar_init(&bar_acc_empty[s], 4);        // epilogue warps
        }
        umma::fence_mbar_init();
    }
    if (warp == 9) umma::tmem_alloc<512>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    pdl_wait();
    bool triggered = false;

    auto tile_of = [&](int t, int &b, int &h, int &w0) {
        const int tw = t % cp.tiles_w;
        const int r = t / cp.tiles_w;
        h = r % cp.H; b = r / cp.H;
        w0 = tw * kCtM;
    };

    if (warp < 4) {
        // ================================ converters: lo = x - trunc_tf32(x) ===================
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x)
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % S;
                const float4 *raw = reinterpret_cast<const float4 *>(smem + (size_t)s * kStageBytes);
                float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * kStageBytes + kRawBytes);
                umma::mbar_wait_sleep(&bar_raw[s], (it / S) & 1);
#pragma unroll 4
                for (int i = tid; i < kRawBytes / 16; i += 128) {
                    const float4 v = raw[i];
                    float4 l;
                    float h;
                    umma::split_tf32(v.x, h, l.x); umma::split_tf32(v.y, h, l.y);
                    umma::split_tf32(v.z, h, l.z); umma::split_tf32(v.w, h, l.w);
                    lo[i] = l;
                }
                umma::fence_proxy_async();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_lo[s]);
            }
    } else if (warp < 8) {
        // ================================ epilogue: TMEM -> staging rows -> diagonals -> global ==
        const int q = warp & 3;
        const int m = q * 32 + lane;                       // row of the tile (= w - w0) this lane extracts from
        float *my_stage = stage_mem + (size_t)q * 32 * kPitch;
        const float inv = 1.f / (float)cp.C;
        const long HW = (long)cp.H * cp.W;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
            if (t + (int)gridDim.x >= total) { pdl_trigger(); triggered = true; }
            int b, h, w0;
            tile_of(t, b, h, w0);
            const int a = ti & 1;
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
            // lane (row m) needs columns m + Dp - d, d = 0..D-1: the warp copies the 32 + Dp columns [32 q, 32 q + 32 + Dp)
            constexpr int kChunks = (32 + Dp) / 16;
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                float v[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * 256 + q * 32 + ch * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) my_stage[lane * kPitch + ch * 16 + i] = v[i];
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);      // the accumulator is free for tile i + 2
            if (!NHWC) {
                const int w = w0 + m;
                if (w < cp.W) {
                    float *orow = cp.cost + ((long)b * cp.D * cp.H + h) * cp.W + w;
                    for (int d = 0; d < cp.D; ++d) {
                        const float v = my_stage[lane * kPitch + lane + Dp - d];
                        orow[(long)d * HW] = (w >= d) ? v * inv : 0.f;
                    }
                }
            } else {
                // channels-last: the D values of a pixel are contiguous; 16 lanes x float4 cover 64 disparities of
                // one pixel, two pixels per instruction
                const int sub = lane >> 4, l16 = lane & 15;
                for (int itp = 0; itp < 16; ++itp) {
                    const int pl = itp * 2 + sub;                     // pixel (row) within this warp's quarter
                    const int w = w0 + q * 32 + pl;
                    if (w >= cp.W) continue;
                    float *orow = cp.cost + (((long)b * cp.H + h) * cp.W + w) * cp.D;
                    for (int d = l16 * 4; d < cp.D; d += 64) {
                        const float *sp = my_stage + pl * kPitch + pl + Dp - d;
                        float4 o;
                        o.x = (w >= d) ? sp[0] * inv : 0.f;
                        o.y = (w >= d + 1) ? sp[-1] * inv : 0.f;
                        o.z = (w >= d + 2) ? sp[-2] * inv : 0.f;
                        o.w = (w >= d + 3) ? sp[-3] * inv : 0.f;
                        *reinterpret_cast<float4 *>(orow + d) = o;
                    }
                }
            }
            __syncwarp();                                  // staging rows are rewritten by the next tile
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ================================ TMA producer =========================================
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int b, h, w0;
                tile_of(t, b, h, w0);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % S;
                    umma::mbar_wait_sleep(&bar_empty[s], ((it / S) & 1) ^ 1);
                    umma::mbar_expect_tx(&bar_raw[s], kRawBytes);
                    uint8_t *dst = smem + (size_t)s * kStageBytes;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        umma::tma_load_4d(dst + j * kCtBox, &tmL, w0 + 32 * j, h, kb * kCtKC, b, &bar_raw[s]);
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        umma::tma_load_4d(dst + (4 + j) * kCtBox, &tmR, w0 - Dp + 32 * j, h, kb * kCtKC, b, &bar_raw[s]);
                }
            }
        }
    } else if (warp == 9) {
        {
            // ================================ MMA issuer ============================================
            // warp-uniform loop, one elected lane issues (no per-instruction ELECT retry loops), wrapping stage counter
            constexpr uint32_t idesc = umma::make_idesc_tf32_major(kCtM, N, 1, 1);      // both operands MN-major
            uint32_t ti = 0, stage = 0, phase = 0;
            const uint32_t smem0 = umma::smem_u32(smem);
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * 256;
#pragma unroll 1
                for (int kb = 0; kb < KB; ++kb) {
                    umma::mbar_wait_sleep(&bar_lo[stage], phase);         // raw landed and lo written
                    umma::tc_fence_after();
                    if (umma::elect_one()) {
                        const uint32_t raw = smem0 + stage * (uint32_t)kStageBytes;
                        const uint32_t lo = raw + kRawBytes;
#pragma unroll
                        for (int k = 0; k < kCtKC / 8; ++k) {
                            // one k step = 8 channels = 8 lines of 128 B = two 4-line swizzle atoms (512 B apart) of every 32-w box
                            const uint64_t a_raw = umma::make_desc_mn_tf32(raw + k * 1024, kCtBox, 512);
                            const uint64_t a_lo = umma::make_desc_mn_tf32(lo + k * 1024, kCtBox, 512);
                            const uint64_t b_raw = umma::make_desc_mn_tf32(raw + 4 * kCtBox + k * 1024, kCtBox, 512);
                            const uint64_t b_lo = umma::make_desc_mn_tf32(lo + 4 * kCtBox + k * 1024, kCtBox, 512);
                            if (k == 0) umma::mma_tf32(d_tmem, a_raw, b_raw, idesc, kb != 0);
                            else umma::mma_tf32(d_tmem, a_raw, b_raw, idesc, 1);
                            umma::mma_tf32(d_tmem, a_raw, b_lo, idesc, 1);
                            umma::mma_tf32(d_tmem, a_lo, b_raw, idesc, 1);
                        }
                        umma::tc_commit(&bar_empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
                }
                if (umma::elect_one()) umma::tc_commit(&bar_acc_full[a]);
                __syncwarp();
            }
        }
    }
    if (!triggered) pdl_trigger();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        umma::tc_fence_after();
        umma::tmem_dealloc<512>(tmem_base);
    }
}

template <int N>
static int launch_corr_tma(const CorrTmaParams &cp, const CUtensorMap &tmL, const CUtensorMap &tmR, bool nhwc,
                           cudaStream_t stream) {
    constexpr size_t smem = kCtStages * 2 * (size_t)(4 + N / 32) * kCtBox + (size_t)kCtM * (32 + (N - kCtM) + 1) * 4 + 1024;
    const int grid = cp.total_tiles < num_sms() ? cp.total_tiles : num_sms();
    if (nhwc) {
        cudaFuncSetAttribute(corr_tma_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return launch_pdl(corr_tma_kernel<N, true>, dim3(grid), dim3(kCtThreads), smem, stream, cp, tmL, tmR);
    }
    cudaFuncSetAttribute(corr_tma_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return launch_pdl(corr_tma_kernel<N, false>, dim3(grid), dim3(kCtThreads), smem, stream, cp, tmL, tmR);
}

// Returns AANET_ERR_UNSUPPORTED when the shape should take correlation_umma.cu instead.
int corr_tma_launch(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D, bool nhwc,
                    cudaStream_t stream) {
    { const char *e = getenv("AANET_CORR_TMA"); if (e && e[0] == '0') return AANET_ERR_UNSUPPORTED; }   // A/B switch
    if (D > 64 || W % 4 || !aligned16(L) || !aligned16(R) || !aligned16(cost)) return AANET_ERR_UNSUPPORTED;
    if (nhwc && D % 4) return AANET_ERR_UNSUPPORTED;
    CorrTmaParams cp;
    cp.cost = cost; cp.C = C; cp.H = H; cp.W = W; cp.D = D; cp.B = B;
    cp.tiles_w = ceil_div(W, kCtM);
    const long total = (long)B * H * cp.tiles_w;
    if (total > 0x3fffffffL) return AANET_ERR_UNSUPPORTED;
    cp.total_tiles = (int)total;
    cp.KB = ceil_div(C, kCtKC);
    CUtensorMap tmL, tmR;
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
    const uint32_t box[4] = {32, 1, kCtKC, 1};
    int rc = make_tensor_map_f32(&tmL, L, 4, dims, strides, box, 2);
    if (!rc) rc = make_tensor_map_f32(&tmR, R, 4, dims, strides, box, 2);
    if (rc) return rc;
    if (D <= 32) return launch_corr_tma<160>(cp, tmL, tmR, nhwc, stream);
    return launch_corr_tma<192>(cp, tmL, tmR, nhwc, stream);
}

}  // namespace aanet
