// Correlation cost volume, forward and backward.
//
// Replaces CostVolume.forward, correlation branch (reference nets/cost.py:40-48): the reference
// loops over the D disparities in Python, each iteration materialising a [B,C,H,W-d] product, a
// channel mean and a strided store (~100x the algorithmic HBM traffic, ~340 launches per pair).
// Here one launch per scale computes, per image row, the banded contraction
//     cost[d,w] = (1/C) * sum_c L[c,w] * R[c,w-d],   0 <= w-d,  0 <= d < D
// reading L and R once.
//
// Forward tiling: a CTA owns (b, h, 128 w's, 64 d's).  Channels are streamed through shared
// memory in chunks of 8; the R window for the tile is 128+64 wide so every (w,d) pair in the
// tile finds R[w-d] in shared memory.  A thread accumulates an 8(w) x 8(d) register tile: per
// channel it needs 8 L values and the 15-wide R diagonal band, fetched as 2+4 LDS.128 for 64 FMAs.
// The w<d triangle is written as exact zeros (the reference's new_zeros, cost.py:41).
#include <cuda_bf16.h>
#include "common.cuh"

namespace aanet {

constexpr int kTW = 128;            // w per CTA
constexpr int kTD = 64;             // d per CTA
constexpr int kCK = 8;              // channels per smem stage
constexpr int kRW = kTW + kTD;      // R window width (192)
constexpr int kCorrThreads = 128;   // 16 (w groups of 8) x 8 (d groups of 8)

__global__ void __launch_bounds__(kCorrThreads)
corr_fwd_kernel(const float *__restrict__ L, const float *__restrict__ R, float *__restrict__ cost,
                int C, int H, int W, int D, int n_wtiles) {
    __shared__ __align__(16) float sL[2][kCK][kTW];
    __shared__ __align__(16) float sR[2][kCK][kRW];

    const int wt = blockIdx.x % n_wtiles, dt = blockIdx.x / n_wtiles;
    const int h = blockIdx.y, b = blockIdx.z;
    const int w0 = wt * kTW, d0 = dt * kTD;
    const int tid = threadIdx.x;
    const int tw = (tid & 15) * 8, td = (tid >> 4) * 8;
    const long HW = (long)H * W;
    const float *Lrow = L + (long)b * C * HW + (long)h * W;
    const float *Rrow = R + (long)b * C * HW + (long)h * W;
    // smem R index r  <->  global column  w0 - d0 - kTD + r
    const int rbase = w0 - d0 - kTD;

    // A whole tile above the diagonal band (all w < d) is zeros.
    const bool all_zero = (w0 + kTW - 1) < d0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_stage = [&](int buf, int c0) {
        // L: kCK x kTW floats, R: kCK x kRW floats; coalesced along w, zero-filled out of range
        for (int i = tid; i < kCK * kTW; i += kCorrThreads) {
            const int c = i / kTW, x = i % kTW, w = w0 + x;
            sL[buf][c][x] = (c0 + c < C && w < W) ? Lrow[(long)(c0 + c) * HW + w] : 0.f;
        }
        for (int i = tid; i < kCK * kRW; i += kCorrThreads) {
            const int c = i / kRW, x = i % kRW, w = rbase + x;
            sR[buf][c][x] = (c0 + c < C && w >= 0 && w < W) ? Rrow[(long)(c0 + c) * HW + w] : 0.f;
        }
    };

    if (!all_zero) {
        const int nchunk = ceil_div(C, kCK);
        load_stage(0, 0);
        __syncthreads();
        for (int ck = 0; ck < nchunk; ++ck) {
            const int buf = ck & 1;
            if (ck + 1 < nchunk) load_stage(buf ^ 1, (ck + 1) * kCK);
#pragma unroll
            for (int c = 0; c < kCK; ++c) {
                float l[8], r[16];
                *reinterpret_cast<float4 *>(&l[0]) = *reinterpret_cast<const float4 *>(&sL[buf][c][tw]);
                *reinterpret_cast<float4 *>(&l[4]) = *reinterpret_cast<const float4 *>(&sL[buf][c][tw + 4]);
                // thread's R band: columns (tw+i) - (td+j) + kTD, i,j in 0..7  ->  [tw-td+kTD-7, tw-td+kTD+7]
                const int rb = tw - td + kTD - 8;       // multiple of 8 -> 16B aligned
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4 *>(&r[4 * q]) = *reinterpret_cast<const float4 *>(&sR[buf][c][rb + 4 * q]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(l[i], r[8 + i - j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    const float inv = 1.f / (float)C;
    const bool vec_ok = (W % 4 == 0) && aligned16(cost);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = d0 + td + j;
        if (d >= D) continue;
        float *orow = cost + (((long)b * D + d) * H + h) * W;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int w = w0 + tw + i;
            o[i] = (w >= d) ? acc[i][j] * inv : 0.f;
        }
        const int w = w0 + tw;
        if (vec_ok && w + 8 <= W) {
            *reinterpret_cast<float4 *>(orow + w) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(orow + w + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (w + i < W) orow[w + i] = o[i];
        }
    }
}

// Pipelined variant (W % 4 == 0, 16-byte aligned rows): same tiling, but the channel chunks arrive through a
// 3-stage cp.async ring (16-byte copies, zero-filled outside the image), so the global-load latency of
// chunk k+2 overlaps the FFMAs of chunk k.  ncu on the synchronous version: 36 % issue active, long- and
// short-scoreboard bound.
constexpr int kCorrStages = 3;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int src_bytes = valid ? 16 : 0;                 // 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kCorrThreads)
corr_fwd_pipelined_kernel(const float *__restrict__ L, const float *__restrict__ R, float *__restrict__ cost,
                          int C, int H, int W, int D, int n_wtiles) {
    __shared__ __align__(16) float sL[kCorrStages][kCK][kTW];
    __shared__ __align__(16) float sR[kCorrStages][kCK][kRW];

    const int wt = blockIdx.x % n_wtiles, dt = blockIdx.x / n_wtiles;
    const int h = blockIdx.y, b = blockIdx.z;
    const int w0 = wt * kTW, d0 = dt * kTD;
    const int tid = threadIdx.x;
    const int tw = (tid & 15) * 8, td = (tid >> 4) * 8;
    const long HW = (long)H * W;
    const float *Lrow = L + (long)b * C * HW + (long)h * W;
    const float *Rrow = R + (long)b * C * HW + (long)h * W;
    const int rbase = w0 - d0 - kTD;            // smem R index r <-> global column rbase + r (multiple of 64)
    const bool all_zero = (w0 + kTW - 1) < d0;  // tile entirely above the diagonal band

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nchunk = all_zero ? 0 : ceil_div(C, kCK);
    auto issue = [&](int ck) {
        if (ck < nchunk) {
            const int buf = ck % kCorrStages, c0 = ck * kCK;
            for (int i = tid; i < kCK * (kTW / 4); i += kCorrThreads) {           // 256 copies
                const int c = i / (kTW / 4), x = (i % (kTW / 4)) * 4, w = w0 + x;
                const bool ok = (c0 + c < C) && (w < W);
                cp_async16(&sL[buf][c][x], ok ? Lrow + (long)(c0 + c) * HW + w : Lrow, ok);
            }
            for (int i = tid; i < kCK * (kRW / 4); i += kCorrThreads) {           // 384 copies
                const int c = i / (kRW / 4), x = (i % (kRW / 4)) * 4, w = rbase + x;
                const bool ok = (c0 + c < C) && (w >= 0) && (w < W);
                cp_async16(&sR[buf][c][x], ok ? Rrow + (long)(c0 + c) * HW + w : Rrow, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");     // empty groups keep the wait counts uniform
    };
    issue(0);
    issue(1);
    for (int ck = 0; ck < nchunk; ++ck) {
        issue(ck + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");     // chunk ck has landed
        __syncthreads();
        const int buf = ck % kCorrStages;
#pragma unroll
        for (int c = 0; c < kCK; ++c) {
            float l[8], r[16];
            *reinterpret_cast<float4 *>(&l[0]) = *reinterpret_cast<const float4 *>(&sL[buf][c][tw]);
            *reinterpret_cast<float4 *>(&l[4]) = *reinterpret_cast<const float4 *>(&sL[buf][c][tw + 4]);
            const int rb = tw - td + kTD - 8;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4 *>(&r[4 * q]) = *reinterpret_cast<const float4 *>(&sR[buf][c][rb + 4 * q]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(l[i], r[8 + i - j], acc[i][j]);
        }
        __syncthreads();        // everyone is done with `buf` before chunk ck+3 overwrites it
    }

    const float inv = 1.f / (float)C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = d0 + td + j;
        if (d >= D) continue;
        float *orow = cost + (((long)b * D + d) * H + h) * W;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int w = w0 + tw + i;
            o[i] = (w >= d) ? acc[i][j] * inv : 0.f;
        }
        const int w = w0 + tw;
        if (w + 8 <= W) {
            *reinterpret_cast<float4 *>(orow + w) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(orow + w + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (w + i < W) orow[w + i] = o[i];
        }
    }
}

// bf16-feature variant (BASELINE config 5: "bf16 cost-volume variant with stated EPE tolerance"): L and R are
// bf16, products and the channel sum are fp32, the volume is fp32.  Same tiling and cp.async ring; a 16-byte
// copy now carries 8 features, halving both the HBM reads and the shared-memory operand traffic.  Requires
// W % 8 == 0.
__device__ __forceinline__ void unpack_bf16x8(const uint4 &q, float *f) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__global__ void __launch_bounds__(kCorrThreads)
corr_fwd_bf16_kernel(const __nv_bfloat16 *__restrict__ L, const __nv_bfloat16 *__restrict__ R,
                     float *__restrict__ cost, int C, int H, int W, int D, int n_wtiles) {
    __shared__ __align__(16) __nv_bfloat16 sL[kCorrStages][kCK][kTW];
    __shared__ __align__(16) __nv_bfloat16 sR[kCorrStages][kCK][kRW];

    const int wt = blockIdx.x % n_wtiles, dt = blockIdx.x / n_wtiles;
    const int h = blockIdx.y, b = blockIdx.z;
    const int w0 = wt * kTW, d0 = dt * kTD;
    const int tid = threadIdx.x;
    const int tw = (tid & 15) * 8, td = (tid >> 4) * 8;
    const long HW = (long)H * W;
    const __nv_bfloat16 *Lrow = L + (long)b * C * HW + (long)h * W;
    const __nv_bfloat16 *Rrow = R + (long)b * C * HW + (long)h * W;
    const int rbase = w0 - d0 - kTD;
    const bool all_zero = (w0 + kTW - 1) < d0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nchunk = all_zero ? 0 : ceil_div(C, kCK);
    auto issue = [&](int ck) {
        if (ck < nchunk) {
            const int buf = ck % kCorrStages, c0 = ck * kCK;
            for (int i = tid; i < kCK * (kTW / 8); i += kCorrThreads) {
                const int c = i / (kTW / 8), x = (i % (kTW / 8)) * 8, w = w0 + x;
                const bool ok = (c0 + c < C) && (w < W);
                cp_async16(&sL[buf][c][x], ok ? Lrow + (long)(c0 + c) * HW + w : Lrow, ok);
            }
            for (int i = tid; i < kCK * (kRW / 8); i += kCorrThreads) {
                const int c = i / (kRW / 8), x = (i % (kRW / 8)) * 8, w = rbase + x;
                const bool ok = (c0 + c < C) && (w >= 0) && (w < W);
                cp_async16(&sR[buf][c][x], ok ? Rrow + (long)(c0 + c) * HW + w : Rrow, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    issue(1);
    for (int ck = 0; ck < nchunk; ++ck) {
        issue(ck + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();
        const int buf = ck % kCorrStages;
#pragma unroll
        for (int c = 0; c < kCK; ++c) {
            float l[8], r[16];
            unpack_bf16x8(*reinterpret_cast<const uint4 *>(&sL[buf][c][tw]), l);
            const int rb = tw - td + kTD - 8;
            unpack_bf16x8(*reinterpret_cast<const uint4 *>(&sR[buf][c][rb]), r);
            unpack_bf16x8(*reinterpret_cast<const uint4 *>(&sR[buf][c][rb + 8]), r + 8);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(l[i], r[8 + i - j], acc[i][j]);
        }
        __syncthreads();
    }

    const float inv = 1.f / (float)C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = d0 + td + j;
        if (d >= D) continue;
        float *orow = cost + (((long)b * D + d) * H + h) * W;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int w = w0 + tw + i;
            o[i] = (w >= d) ? acc[i][j] * inv : 0.f;
        }
        const int w = w0 + tw;
        if (w + 8 <= W) {
            *reinterpret_cast<float4 *>(orow + w) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(orow + w + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (w + i < W) orow[w + i] = o[i];
        }
    }
}

// Backward (autograd of cost.py:45-48):
//   gL[c,w]  = (1/C) sum_{d<=w}    g[d,w]    * R[c,w-d]
//   gR[c,w'] = (1/C) sum_{w'+d<W}  g[d,w'+d] * L[c,w'+d]
// A CTA owns (b, h, 64 w's): it stages g[0..D) for the columns [w0, w0+64+D) once in shared memory
// and then walks the channels, each thread producing gL and gR for one (c, w).
constexpr int kBW = 64;
constexpr int kBwdThreads = 256;   // 64 w x 4 channel lanes

__global__ void __launch_bounds__(kBwdThreads)
corr_bwd_kernel(const float *__restrict__ L, const float *__restrict__ R, const float *__restrict__ g,
                float *__restrict__ gL, float *__restrict__ gR, int C, int H, int W, int D) {
    extern __shared__ float sg[];          // [D][kBW + D]  (row stride GW)
    const int GW = kBW + D;
    const int w0 = blockIdx.x * kBW, h = blockIdx.y, b = blockIdx.z;
    const long HW = (long)H * W;
    for (int i = threadIdx.x; i < D * GW; i += kBwdThreads) {
        const int d = i / GW, x = i % GW, w = w0 + x;
        sg[i] = (w < W) ? g[(((long)b * D + d) * H + h) * W + w] : 0.f;
    }
    __syncthreads();
    const int x = threadIdx.x % kBW, cl = threadIdx.x / kBW;
    const int w = w0 + x;
    if (w >= W) return;
    const float inv = 1.f / (float)C;
    for (int c = cl; c < C; c += kBwdThreads / kBW) {
        const float *lrow = L + ((long)b * C + c) * HW + (long)h * W;
        const float *rrow = R + ((long)b * C + c) * HW + (long)h * W;
        float al = 0.f, ar = 0.f;
        const int dl = min(D - 1, w);            // d <= w
        for (int d = 0; d <= dl; ++d) al = fmaf(sg[d * GW + x], rrow[w - d], al);
        const int dr = min(D - 1, W - 1 - w);    // w + d < W
        for (int d = 0; d <= dr; ++d) ar = fmaf(sg[d * GW + x + d], lrow[w + d], ar);
        gL[((long)b * C + c) * HW + (long)h * W + w] = al * inv;
        gR[((long)b * C + c) * HW + (long)h * W + w] = ar * inv;
    }
}

}  // namespace aanet

namespace aanet {
// correlation_umma.cu
bool corr_umma_supported(int B, int C, int H, int W, int D);
int corr_umma_launch(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D, bool nhwc,
                     cudaStream_t stream);
// correlation_tma.cu (AANET_ERR_UNSUPPORTED = take correlation_umma.cu)
int corr_tma_launch(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D, bool nhwc,
                    cudaStream_t stream);
}  // namespace aanet

using namespace aanet;

extern "C" int aanet_corr_fwd(const float *L, const float *R, float *cost, int B, int C, int H, int W,
                              int D, void *stream) {
    if (!L || !R || !cost) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (H > 65535 || B > 65535) return AANET_ERR_UNSUPPORTED;
    // tensor-core paths (D <= 128): TMA-staged tiles where the shape allows, else the register-staged kernel; the
    // FFMA kernels below remain for wider searches
    if (corr_umma_supported(B, C, H, W, D)) {
        const int rc = corr_tma_launch(L, R, cost, B, C, H, W, D, false, as_stream(stream));
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
        return corr_umma_launch(L, R, cost, B, C, H, W, D, false, as_stream(stream));
    }
    const int n_wtiles = ceil_div(W, kTW), n_dtiles = ceil_div(D, kTD);
    const dim3 grid(n_wtiles * n_dtiles, H, B);
    if (W % 4 == 0 && aligned16(L) && aligned16(R) && aligned16(cost))
        corr_fwd_pipelined_kernel<<<grid, kCorrThreads, 0, as_stream(stream)>>>(L, R, cost, C, H, W, D, n_wtiles);
    else
        corr_fwd_kernel<<<grid, kCorrThreads, 0, as_stream(stream)>>>(L, R, cost, C, H, W, D, n_wtiles);
    return check_launch();
}

extern "C" int aanet_corr_fwd_nhwc(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D,
                                   void *stream) {
    if (!L || !R || !cost) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (!corr_umma_supported(B, C, H, W, D)) return AANET_ERR_UNSUPPORTED;
    if (D % 4 == 0) {
        const int rc = corr_tma_launch(L, R, cost, B, C, H, W, D, true, as_stream(stream));
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
    }
    return corr_umma_launch(L, R, cost, B, C, H, W, D, true, as_stream(stream));
}

extern "C" int aanet_corr_fwd_bf16(const void *L, const void *R, float *cost, int B, int C, int H, int W, int D,
                                   void *stream) {
    if (!L || !R || !cost) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (H > 65535 || B > 65535) return AANET_ERR_UNSUPPORTED;
    if (W % 8 || !aligned16(static_cast<const char *>(L)) || !aligned16(static_cast<const char *>(R)) ||
        !aligned16(cost))
        return AANET_ERR_UNSUPPORTED;
    const int n_wtiles = ceil_div(W, kTW), n_dtiles = ceil_div(D, kTD);
    const dim3 grid(n_wtiles * n_dtiles, H, B);
    corr_fwd_bf16_kernel<<<grid, kCorrThreads, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16 *>(L), static_cast<const __nv_bfloat16 *>(R), cost, C, H, W, D, n_wtiles);
    return check_launch();
}

extern "C" int aanet_corr_bwd(const float *L, const float *R, const float *gcost, float *gL, float *gR,
                              int B, int C, int H, int W, int D, void *stream) {
    if (!L || !R || !gcost || !gL || !gR) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (H > 65535 || B > 65535) return AANET_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * (size_t)D * (kBW + D);
    if (smem > 200 * 1024) return AANET_ERR_UNSUPPORTED;
    // per launch: the attribute is per device (a process driving several GPUs, e.g. nn.DataParallel threads)
    cudaFuncSetAttribute(corr_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const dim3 grid(ceil_div(W, kBW), H, B);
    corr_bwd_kernel<<<grid, kBwdThreads, smem, as_stream(stream)>>>(L, R, gcost, gL, gR, C, H, W, D);
    return check_launch();
}
