// Refinement front end: the part of StereoDRNetRefinement / HourglassRefinement.forward
// that precedes the first convolution (reference nets/refinement.py:80-95 = :144-160, nets/warp.py:41-64):
//
//   disp   = bilinear(low_disp, (H,W), align_corners=False) * (W / w)      (low_disp itself when W == w)
//   warped = grid_sample(right, (x - disp, y), bilinear, padding_mode='border', align_corners=True)
//   concat = cat(warped - left, left)
//
// The reference spends ~15 launches on it (interpolate, mul, meshgrid, cat, two normalisations, two grid_samples --
// one only to build a mask it then drops -- sub, cat) plus a device synchronisation (`assert disp.min() >= 0`,
// warp.py:51, which also blocks CUDA-graph capture).  Here: one pass, one thread per output pixel, every image
// plane read and written once with consecutive lanes on consecutive pixels.
//
// The sampling position goes through the reference's normalise / un-normalise round trip in fp32
// (2*(g/(size-1)) - 1, then ((g+1)/2)*(size-1)) so that it rounds the same way.
#include "common.cuh"

namespace aanet {

__device__ __forceinline__ void up_index(int dst, int in, float scale, int &i0, int &i1, float &l0, float &l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;      // ATen area_pixel_compute_source_index, align_corners=False
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in - 1);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256)
refine_frontend_kernel(const float *__restrict__ low, const float *__restrict__ left, const float *__restrict__ right,
                       float *__restrict__ concat, float *__restrict__ disp, int C, int h, int w, int H, int W,
                       float scale) {
    pdl_wait();
    const long HW = (long)H * W;
    // grid = (image rows, 256-pixel segments of a row): no per-pixel index divisions
    const int x = blockIdx.y * blockDim.x + threadIdx.x;
    const int y = blockIdx.x % H;
    const long b = blockIdx.x / H;
    if (x < W) {
        const long i = (b * H + y) * W + x;
        float d;
        if (W == w) {
            d = __ldg(low + i);                        // same size: [B,h,w] == [B,H,W] only when h == H too
        } else {
            const float *src = low + b * h * w;
            int h0, h1, w0, w1; float a0, a1, b0, b1;
            up_index(y, h, (float)h / (float)H, h0, h1, a0, a1);
            up_index(x, w, (float)w / (float)W, w0, w1, b0, b1);
            d = a0 * (b0 * __ldg(src + (long)h0 * w + w0) + b1 * __ldg(src + (long)h0 * w + w1)) +
                a1 * (b0 * __ldg(src + (long)h1 * w + w0) + b1 * __ldg(src + (long)h1 * w + w1));
            d = d * scale;
        }
        disp[i] = d;
        float gx = (float)x - d, gy = (float)y;
        gx = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(gx, (float)(W - 1))), 1.f);      // warp.py:12-13, no contraction
        gy = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(gy, (float)(H - 1))), 1.f);
        float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(W - 1));  // grid_sampler_unnormalize
        float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(H - 1));
        ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));      // padding_mode = 'border'
        iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
        const float fx = floorf(ix), fy = floorf(iy);
        const int xw = (int)fx, yn = (int)fy;
        const float nw = (fx + 1.f - ix) * (fy + 1.f - iy), ne = (ix - fx) * (fy + 1.f - iy);
        const float sw = (fx + 1.f - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
        const bool x1 = xw + 1 < W, y1 = yn + 1 < H;
        const long o00 = (long)yn * W + xw;
        for (int c = 0; c < C; ++c) {
            const float *img = right + (b * C + c) * HW;
            float v = __ldg(img + o00) * nw;
            if (x1) v += __ldg(img + o00 + 1) * ne;
            if (y1) v += __ldg(img + o00 + W) * sw;
            if (x1 && y1) v += __ldg(img + o00 + W + 1) * se;
            const long pl = (b * C + c) * HW + (long)y * W + x;
            const float l = __ldg(left + pl);
            concat[(b * 2 * C + c) * HW + (long)y * W + x] = v - l;
            concat[(b * 2 * C + C + c) * HW + (long)y * W + x] = l;
        }
    }
    pdl_trigger();
}

}  // namespace aanet

using namespace aanet;

extern "C" int aanet_refine_frontend_fwd(const float *low_disp, const float *left, const float *right, float *concat,
                                         float *disp, int B, int C, int h, int w, int H, int W, void *stream) {
    if (!low_disp || !left || !right || !concat || !disp) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 1 || W <= 1) return AANET_ERR_SHAPE;
    if (W == w && H != h) return AANET_ERR_SHAPE;       // refinement.py:84-85 takes low_disp as is when the widths agree
    if ((long)B * H > 0x7fffffffL || ceil_div(W, 256) > 65535) return AANET_ERR_UNSUPPORTED;
    const float scale = (float)((double)W / (double)w);
    return launch_pdl(refine_frontend_kernel, dim3((unsigned)(B * H), ceil_div(W, 256)), dim3(256), 0,
                      as_stream(stream), low_disp, left, right, concat, disp, C, h, w, H, W, scale);
}
