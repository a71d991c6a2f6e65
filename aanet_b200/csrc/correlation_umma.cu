// Correlation cost volume on the tensor cores (tcgen05, 3xTF32).
//
// Same operator as correlation.cu (reference nets/cost.py:40-48):
//     cost[b,d,h,w] = (1/C) * sum_c L[b,c,h,w] * R[b,c,h,w-d]   (0 for w < d)
// For one image row the band is a slice of a plain GEMM:  G[w, w'] = sum_c L[c,w] R[c,w'],  cost[d,w] = G[w, w-d].
// A CTA owns (b, h, 128 consecutive w): M = 128 pixels (one TMEM lane each), N = 128 + Dp columns w' in
// [w0 - Dp, w0 + 128) (Dp = D rounded up to a supported width), K = C in blocks of 32 channels.  Two thirds of G
// are outside the band and every product is issued three times (hi*hi + hi*lo + lo*hi, the parity bar is 1e-4),
// and it is still ~4x faster than the FFMA kernel, which is shared-memory bound (6 LDS.128 per 64 FMA).
//
//   staging   all 256 threads: LDG.32 coalesced along w (the tensors are NCHW), 4 channels per item, split into
//             tf32 hi + lo, two STS.128 into K-major SWIZZLE_128B tiles ([row = pixel][32 channels]) -- the
//             NCHW -> K-major transposition happens in registers.  The loads of K block k+1 are issued before the
//             MMAs of block k are waited for; the second CTA of the SM (80 KB of smem, 256 TMEM columns each)
//             fills the remaining bubbles.
//   MMA       one thread, 12 tcgen05.mma.kind::tf32 (M = 128, N, K = 8) per K block, tcgen05.commit -> mbarrier.
//   epilogue  lane w needs columns w + Dp - d, a per-lane window: every warp copies the 32 + Dp columns its 32
//             lanes can need from TMEM to a shared staging row (pitch = 1 mod 32: conflict-free), then reads the
//             diagonals back and writes cost[d][w0 + 32 q .. + 31] as 128-byte rows, scaled by 1/C, zero for w < d.
#include "common.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kCuM = 128;          // w per CTA (UMMA M)
constexpr int kCuThreads = 256;
constexpr int kCuTmemCols = 256;   // one accumulator of N <= 256 columns; two CTAs per SM

// NHWC = false: cost is [B,D,H,W] (the reference's layout); true: channels-last [B,H,W,D] for the fused aggregation
// executor (D % 4 == 0), which saves the layout kernel in front of the first 1x1 convolution.
template <int N, bool NHWC>
__global__ void __launch_bounds__(kCuThreads, 2)
corr_umma_kernel(const float *__restrict__ L, const float *__restrict__ R, float *__restrict__ cost, int C, int H,
                 int W, int D) {
    constexpr int Dp = N - kCuM;                       // window margin (>= D)
    constexpr int kABytes = kCuM * 32 * 4, kBBytes = N * 32 * 4;
    constexpr int kPitch = 32 + Dp + 1;                // staging row pitch in floats
    constexpr int kBItems = (N * 8 + kCuThreads - 1) / kCuThreads;
    static_assert(kCuM * kPitch * 4 <= 2 * kABytes + 2 * kBBytes, "epilogue staging reuses the operand tiles");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);
    float *a_hi = reinterpret_cast<float *>(smem), *a_lo = a_hi + kABytes / 4;
    float *b_hi = a_lo + kABytes / 4, *b_lo = b_hi + kBBytes / 4;

    const int w0 = blockIdx.x * kCuM, h = blockIdx.y, b = blockIdx.z;
    const int wb0 = w0 - Dp;                           // global column of B row 0
    const long HW = (long)H * W;
    const float *Lrow = L + (long)b * C * HW + (long)h * W;
    const float *Rrow = R + (long)b * C * HW + (long)h * W;

    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc<kCuTmemCols>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    // item = (row, channel quad cq): 4 channels of one column.  A: 128 x 8 items, 4 per thread; B: N x 8 items.
    float ra[4][4], rb[kBItems][4];
    auto load = [&](int kb) {
        const int c0 = kb * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = tid & 127, cq = (tid >> 7) + 2 * j;
            const int gw = w0 + r;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + cq * 4 + e;
                ra[j][e] = (c < C && gw < W) ? __ldg(Lrow + (long)c * HW + gw) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < kBItems; ++j) {
            const int i = tid + j * kCuThreads;
            const int cq = i / N, r = i - cq * N;
            const int gw = wb0 + r;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + cq * 4 + e;
                rb[j][e] = (i < N * 8 && c < C && gw >= 0 && gw < W) ? __ldg(Rrow + (long)c * HW + gw) : 0.f;
            }
        }
    };
    auto put = [&](float *hi_t, float *lo_t, int r, int cq, const float (&v)[4]) {
        float4 h4, l4;
        umma::split_tf32(v[0], h4.x, l4.x); umma::split_tf32(v[1], h4.y, l4.y);
        umma::split_tf32(v[2], h4.z, l4.z); umma::split_tf32(v[3], h4.w, l4.w);
        const int at = r * 32 + ((cq ^ (r & 7)) << 2);
        *reinterpret_cast<float4 *>(hi_t + at) = h4;
        *reinterpret_cast<float4 *>(lo_t + at) = l4;
    };
    auto store = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) put(a_hi, a_lo, tid & 127, (tid >> 7) + 2 * j, ra[j]);
#pragma unroll
        for (int j = 0; j < kBItems; ++j) {
            const int i = tid + j * kCuThreads;
            const int cq = i / N, r = i - cq * N;
            if (i < N * 8) put(b_hi, b_lo, r, cq, rb[j]);
        }
    };

    const int KB = (C + 31) / 32;
    constexpr uint32_t idesc = umma::make_idesc_tf32(kCuM, N);
    load(0);
    for (int kb = 0; kb < KB; ++kb) {
        if (kb > 0) {                                   // MMAs of block kb-1 have read the tiles
            umma::mbar_wait(&bar, (kb - 1) & 1);
            umma::tc_fence_after();
        }
        store();
        umma::fence_proxy_async();
        umma::tc_fence_before();
        __syncthreads();
        if (kb + 1 < KB) load(kb + 1);                  // in flight while the tensor core works on block kb
        if (tid == 0) {
            umma::tc_fence_after();
            const uint32_t s0 = umma::smem_u32(smem);
            const uint64_t dah = umma::make_desc_sw128(s0), dal = umma::make_desc_sw128(s0 + kABytes);
            const uint64_t dbh = umma::make_desc_sw128(s0 + 2 * kABytes);
            const uint64_t dbl = umma::make_desc_sw128(s0 + 2 * kABytes + kBBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t adv = k * 32;
                umma::mma_tf32(tmem_base, umma::desc_advance(dah, adv), umma::desc_advance(dbh, adv), idesc, (kb | k) != 0);
                umma::mma_tf32(tmem_base, umma::desc_advance(dah, adv), umma::desc_advance(dbl, adv), idesc, 1);
                umma::mma_tf32(tmem_base, umma::desc_advance(dal, adv), umma::desc_advance(dbh, adv), idesc, 1);
            }
            umma::tc_commit(&bar);
        }
    }
    umma::mbar_wait(&bar, (KB - 1) & 1);
    umma::tc_fence_after();

    // ---- epilogue: TMEM -> staging rows -> diagonals -> global
    float *stage = reinterpret_cast<float *>(smem);     // operand tiles are dead now
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    constexpr int kChunks = (32 + Dp) / 16;
    {
        const int ch_lo = half ? kChunks / 2 : 0, ch_hi = half ? kChunks : kChunks / 2;
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
            float v[16];
            umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + q * 32 + ch * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) stage[m * kPitch + ch * 16 + i] = v[i];
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (!NHWC) {
        const float inv = 1.f / (float)C;
        const int w = w0 + m;
        if (w < W) {
            float *orow = cost + ((long)b * D * H + h) * W + w;
            for (int d = half; d < D; d += 2) {
                const float v = stage[m * kPitch + lane + Dp - d];
                orow[(long)d * HW] = (w >= d) ? v * inv : 0.f;
            }
        }
    } else {
        // channels-last: the D values of a pixel are contiguous.  16 lanes x float4 cover 64 disparities of one pixel
        // (two pixels per instruction); a warp takes 16 of its quarter's 32 pixels.
        const float inv = 1.f / (float)C;
        const int sub = lane >> 4, l16 = lane & 15;
        for (int it = 0; it < 8; ++it) {
            const int pl = half * 16 + it * 2 + sub;          // pixel within the quarter (= its lane index in stage)
            const int mm = q * 32 + pl, w = w0 + mm;
            if (w >= W) continue;
            float *orow = cost + (((long)b * H + h) * W + w) * D;
            for (int d = l16 * 4; d < D; d += 64) {
                const float *sp = stage + mm * kPitch + pl + Dp - d;
                float4 o;
                o.x = (w >= d) ? sp[0] * inv : 0.f;
                o.y = (w >= d + 1) ? sp[-1] * inv : 0.f;
                o.z = (w >= d + 2) ? sp[-2] * inv : 0.f;
                o.w = (w >= d + 3) ? sp[-3] * inv : 0.f;
                *reinterpret_cast<float4 *>(orow + d) = o;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        umma::tc_fence_after();
        umma::tmem_dealloc<kCuTmemCols>(tmem_base);
    }
}

template <int N>
static int launch_corr_umma(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D,
                            bool nhwc, cudaStream_t stream) {
    constexpr size_t smem = 2 * (kCuM * 32 * 4) + 2 * (N * 32 * 4) + 1024;
    const dim3 grid(ceil_div(W, kCuM), H, B);
    if (nhwc) {
        cudaFuncSetAttribute(corr_umma_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        corr_umma_kernel<N, true><<<grid, kCuThreads, smem, stream>>>(L, R, cost, C, H, W, D);
    } else {
        cudaFuncSetAttribute(corr_umma_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        corr_umma_kernel<N, false><<<grid, kCuThreads, smem, stream>>>(L, R, cost, C, H, W, D);
    }
    return check_launch();
}

bool corr_umma_supported(int B, int C, int H, int W, int D) {
    return D <= 128 && H <= 65535 && B <= 65535 && (long)C * H * W < 0x7fffffffL;
}

int corr_umma_launch(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D, bool nhwc,
                     cudaStream_t stream) {
    if (nhwc && (D % 4 || !aligned16(cost))) return AANET_ERR_UNSUPPORTED;
    if (D <= 16) return launch_corr_umma<144>(L, R, cost, B, C, H, W, D, nhwc, stream);
    if (D <= 32) return launch_corr_umma<160>(L, R, cost, B, C, H, W, D, nhwc, stream);
    if (D <= 64) return launch_corr_umma<192>(L, R, cost, B, C, H, W, D, nhwc, stream);
    if (D <= 96) return launch_corr_umma<224>(L, R, cost, B, C, H, W, D, nhwc, stream);
    return launch_corr_umma<256>(L, R, cost, B, C, H, W, D, nhwc, stream);
}

}  // namespace aanet
