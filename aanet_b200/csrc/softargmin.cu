// Soft-argmin disparity regression, forward and backward (HBM-bound).
//
// Replaces DisparityEstimation.forward (reference nets/estimation.py:13-30): softmax over the
// disparity axis of a [B,D,H,W] volume followed by sum_d d*p_d.  The reference runs softmax, arange,
// broadcast-mul and sum as four torch kernels (~5x the algorithmic traffic); here the volume is
// read exactly once (forward) with a single-pass online softmax.
//
// Mapping: for fixed (b,d) the H*W plane is contiguous, so a warp reads 32 x float4 = 512
// contiguous bytes per disparity.  The D axis is split across the 8 warps of the CTA (warp y
// handles d = y, y+8, ...) so that even the B=1 1/3-scale volume (53 248 pixels) yields 416
// CTAs / 3 328 warps -- enough bytes in flight to cover HBM latency on 148 SMs.  The partial
// (max, sum, weighted sum) triples are merged through shared memory.
#include "common.cuh"

namespace aanet {

constexpr int kDS = 8;      // D slices (warps) per CTA
constexpr int kUnroll = 4;  // disparities per online-softmax update

template <int VEC> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4], bool stream) {
        float4 t = stream ? ldg_stream4(p) : *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1], bool stream) {
        v[0] = stream ? ldg_stream(p) : *p;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
};

// Running (m, s, ws) over this thread's share of the D axis.
template <int VEC, bool STREAM>
__device__ __forceinline__ void partial_softmax(const float *__restrict__ c, long HW, int D, float sgn,
                                                float (&m)[VEC], float (&s)[VEC], float (&ws)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) { m[i] = -INFINITY; s[i] = 0.f; ws[i] = 0.f; }
    int d = threadIdx.y;
    for (; d + (kUnroll - 1) * kDS < D; d += kUnroll * kDS) {
        float v[kUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) Vec<VEC>::load(c + (long)(d + u * kDS) * HW, v[u], STREAM);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float mx = m[i];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) mx = fmaxf(mx, sgn * v[u][i]);
            const float r = __expf(m[i] - mx);
            float ss = s[i] * r, ww = ws[i] * r;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const float e = __expf(sgn * v[u][i] - mx);
                ss += e;
                ww = fmaf(e, (float)(d + u * kDS), ww);
            }
            m[i] = mx; s[i] = ss; ws[i] = ww;
        }
    }
    for (; d < D; d += kDS) {
        float v[VEC];
        Vec<VEC>::load(c + (long)d * HW, v, STREAM);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float x = sgn * v[i];
            const float mx = fmaxf(m[i], x);
            const float r = __expf(m[i] - mx), e = __expf(x - mx);
            s[i] = s[i] * r + e;
            ws[i] = fmaf(e, (float)d, ws[i] * r);
            m[i] = mx;
        }
    }
}

// Merge the kDS partial triples of each pixel; on return every thread with threadIdx.y == 0 holds
// the pixel's (M, S, WS) in (m, s, ws).
template <int VEC>
__device__ __forceinline__ void merge_slices(float (&m)[VEC], float (&s)[VEC], float (&ws)[VEC],
                                             float (*sm)[3][32 * VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        sm[threadIdx.y][0][threadIdx.x * VEC + i] = m[i];
        sm[threadIdx.y][1][threadIdx.x * VEC + i] = s[i];
        sm[threadIdx.y][2][threadIdx.x * VEC + i] = ws[i];
    }
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float M = m[i];
#pragma unroll
            for (int y = 1; y < kDS; ++y) M = fmaxf(M, sm[y][0][threadIdx.x * VEC + i]);
            float S = 0.f, WS = 0.f;
#pragma unroll
            for (int y = 0; y < kDS; ++y) {
                const float r = __expf(sm[y][0][threadIdx.x * VEC + i] - M);
                S = fmaf(sm[y][1][threadIdx.x * VEC + i], r, S);
                WS = fmaf(sm[y][2][threadIdx.x * VEC + i], r, WS);
            }
            m[i] = M; s[i] = S; ws[i] = WS;
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(32 * kDS)
softargmin_fwd_kernel(const float *__restrict__ cost, float *__restrict__ disp, int D, long HW, float sgn) {
    __shared__ float sm[kDS][3][32 * VEC];
    const long p0 = ((long)blockIdx.x * 32 + threadIdx.x) * VEC;
    const bool live = p0 < HW;      // HW % VEC == 0 by construction
    const float *c = cost + (long)blockIdx.y * D * HW + (live ? p0 : 0);
    float m[VEC], s[VEC], ws[VEC];
    partial_softmax<VEC, true>(c, HW, D, sgn, m, s, ws);
    merge_slices<VEC>(m, s, ws, sm);
    if (threadIdx.y == 0 && live) {
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = ws[i] / s[i];
        Vec<VEC>::store(disp + (long)blockIdx.y * HW + p0, o);
    }
}

template <int VEC>
__global__ void __launch_bounds__(32 * kDS)
softargmin_bwd_kernel(const float *__restrict__ cost, const float *__restrict__ gdisp,
                      float *__restrict__ gcost, int D, long HW, float sgn) {
    __shared__ float sm[kDS][3][32 * VEC];
    __shared__ float fin[3][32 * VEC];   // M, 1/S, disp per pixel
    const long p0 = ((long)blockIdx.x * 32 + threadIdx.x) * VEC;
    const bool live = p0 < HW;
    const float *c = cost + (long)blockIdx.y * D * HW + (live ? p0 : 0);
    float m[VEC], s[VEC], ws[VEC];
    partial_softmax<VEC, false>(c, HW, D, sgn, m, s, ws);
    merge_slices<VEC>(m, s, ws, sm);
    if (threadIdx.y == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            fin[0][threadIdx.x * VEC + i] = m[i];
            fin[1][threadIdx.x * VEC + i] = 1.f / s[i];
            fin[2][threadIdx.x * VEC + i] = ws[i] / s[i];
        }
    }
    __syncthreads();
    if (!live) return;
    float g[VEC];
    Vec<VEC>::load(gdisp + (long)blockIdx.y * HW + p0, g, false);
    float M[VEC], rS[VEC], dsp[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        M[i] = fin[0][threadIdx.x * VEC + i];
        rS[i] = fin[1][threadIdx.x * VEC + i] * sgn * g[i];   // sign * g / S
        dsp[i] = fin[2][threadIdx.x * VEC + i];
    }
    float *gc = gcost + (long)blockIdx.y * D * HW + p0;
    for (int d = threadIdx.y; d < D; d += kDS) {
        float v[VEC], o[VEC];
        Vec<VEC>::load(c + (long)d * HW, v, false);     // second touch: L1/L2 hit
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = __expf(sgn * v[i] - M[i]) * rS[i] * ((float)d - dsp[i]);
        Vec<VEC>::store(gc + (long)d * HW, o);
    }
}

}  // namespace aanet

using namespace aanet;

extern "C" int aanet_softargmin_fwd(const float *cost, float *disp, int B, int D, int H, int W,
                                    int similarity, void *stream) {
    if (!cost || !disp) return AANET_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return AANET_ERR_SHAPE;
    const long HW = (long)H * W;
    const float sgn = similarity ? 1.f : -1.f;
    const dim3 block(32, kDS);
    if (B > 65535) return AANET_ERR_UNSUPPORTED;
    if (HW % 4 == 0 && aligned16(cost) && aligned16(disp)) {
        const dim3 grid((unsigned)ceil_div_ll(HW / 4, 32), B);
        softargmin_fwd_kernel<4><<<grid, block, 0, as_stream(stream)>>>(cost, disp, D, HW, sgn);
    } else {
        const dim3 grid((unsigned)ceil_div_ll(HW, 32), B);
        softargmin_fwd_kernel<1><<<grid, block, 0, as_stream(stream)>>>(cost, disp, D, HW, sgn);
    }
    return check_launch();
}

extern "C" int aanet_softargmin_bwd(const float *cost, const float *gdisp, float *gcost, int B, int D,
                                    int H, int W, int similarity, void *stream) {
    if (!cost || !gdisp || !gcost) return AANET_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return AANET_ERR_SHAPE;
    if (B > 65535) return AANET_ERR_UNSUPPORTED;
    const long HW = (long)H * W;
    const float sgn = similarity ? 1.f : -1.f;
    const dim3 block(32, kDS);
    if (HW % 4 == 0 && aligned16(cost) && aligned16(gdisp) && aligned16(gcost)) {
        const dim3 grid((unsigned)ceil_div_ll(HW / 4, 32), B);
        softargmin_bwd_kernel<4><<<grid, block, 0, as_stream(stream)>>>(cost, gdisp, gcost, D, HW, sgn);
    } else {
        const dim3 grid((unsigned)ceil_div_ll(HW, 32), B);
        softargmin_bwd_kernel<1><<<grid, block, 0, as_stream(stream)>>>(cost, gdisp, gcost, D, HW, sgn);
    }
    return check_launch();
}
