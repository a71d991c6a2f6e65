// 5-D cost volumes of the non-correlation variants (nets/cost.py:22-38): 'difference' and 'concat'.
//   difference: out[b,c,d,h,w]   = L[b,c,h,w] - R[b,c,h,w-d]                      (w >= d), 0 otherwise
//   concat    : out[b,c,d,h,w]   = L[b,c,h,w],  out[b,C+c,d,h,w] = R[b,c,h,w-d]   (w >= d), 0 otherwise
// Pure data movement: the volume is D times the size of the features, so the kernel is bound by its HBM writes
// (4*B*Cout*D*H*W bytes; the features stay in L1/L2).  One thread produces 4 consecutive w of one (b, co, d, h)
// row -- a 128-bit store -- reading L aligned and R through scalar loads (w - d is not 16-byte aligned in general).
// The reference materialises D slices with a Python loop (2 kernels per disparity).  Results are bit-identical.
#include "common.cuh"

namespace aanet {

template <int MODE>   // 0 = difference, 1 = concat
__global__ void __launch_bounds__(256)
cost5d_fwd_kernel(const float *__restrict__ L, const float *__restrict__ R, float *__restrict__ out, int C, int H,
                  int W, int D, long n_vec, int Wv) {
    const int Cout = MODE == 0 ? C : 2 * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        const int wv = (int)(i % Wv);
        long r = i / Wv;
        const int h = (int)(r % H); r /= H;
        const int d = (int)(r % D); r /= D;
        const int co = (int)(r % Cout);
        const long b = r / Cout;
        const int w0 = wv * 4;
        const bool right = MODE == 1 && co >= C;
        const int c = right ? co - C : co;
        const float *Lr = L + ((b * C + c) * H + h) * (long)W;
        const float *Rr = R + ((b * C + c) * H + h) * (long)W;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int w = w0 + k;
            float x = 0.f;
            if (w < W && w >= d) {
                if (MODE == 0) x = __ldg(Lr + w) - __ldg(Rr + w - d);
                else x = right ? __ldg(Rr + w - d) : __ldg(Lr + w);
            }
            v[k] = x;
        }
        float *o = out + (((b * Cout + co) * D + d) * H + h) * (long)W + w0;
        if ((W & 3) == 0) {
            __stcs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));     // streaming: written once
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (w0 + k < W) o[k] = v[k];
        }
    }
}

// Gradients (autograd of cost.py:22-38): gL[b,c,h,w] = sum_{d <= w} g[b,c,d,h,w];
// gR[b,c,h,w'] = -+ sum_{d, w'+d < W} g[b,cR,d,h,w'+d]   (difference: minus, cR = c; concat: plus, cR = C + c).
// One thread per (b, c, h, w); the d loop strides by H*W, consecutive threads read consecutive w.
template <int MODE>
__global__ void __launch_bounds__(256)
cost5d_bwd_kernel(const float *__restrict__ g, float *__restrict__ gL, float *__restrict__ gR, int C, int H, int W,
                  int D, long n) {
    const int Cout = MODE == 0 ? C : 2 * C;
    const long HW = (long)H * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        long r = i / W;
        const int h = (int)(r % H); r /= H;
        const int c = (int)(r % C);
        const long b = r / C;
        const float *gl = g + ((b * Cout + c) * D) * HW + (long)h * W;
        const float *gr = g + ((b * Cout + (MODE == 0 ? c : C + c)) * D) * HW + (long)h * W;
        float sl = 0.f, sr = 0.f;
        for (int d = 0; d < D; ++d) {
            if (d <= w) sl += __ldg(gl + d * HW + w);
            if (w + d < W) sr += __ldg(gr + d * HW + w + d);
        }
        gL[i] = sl;
        gR[i] = MODE == 0 ? -sr : sr;
    }
}

}  // namespace aanet

using namespace aanet;

extern "C" int aanet_cost5d_fwd(const float *L, const float *R, float *out, int B, int C, int H, int W, int D,
                                int mode, void *stream) {
    if (!L || !R || !out) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (mode != 0 && mode != 1) return AANET_ERR_UNSUPPORTED;
    const int Wv = ceil_div(W, 4);
    const long n_vec = (long)B * (mode ? 2 * C : C) * D * H * Wv;
    const long blocks = ceil_div_ll(n_vec, 256);
    const int grid = (int)(blocks < 32L * num_sms() ? blocks : 32L * num_sms());
    if (mode == 0) cost5d_fwd_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(L, R, out, C, H, W, D, n_vec, Wv);
    else cost5d_fwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(L, R, out, C, H, W, D, n_vec, Wv);
    return check_launch();
}

extern "C" int aanet_cost5d_bwd(const float *gout, float *gL, float *gR, int B, int C, int H, int W, int D, int mode,
                                void *stream) {
    if (!gout || !gL || !gR) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return AANET_ERR_SHAPE;
    if (mode != 0 && mode != 1) return AANET_ERR_UNSUPPORTED;
    const long n = (long)B * C * H * W;
    const long blocks = ceil_div_ll(n, 256);
    const int grid = (int)(blocks < 32L * num_sms() ? blocks : 32L * num_sms());
    if (mode == 0) cost5d_bwd_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(gout, gL, gR, C, H, W, D, n);
    else cost5d_bwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(gout, gL, gR, C, H, W, D, n);
    return check_launch();
}
