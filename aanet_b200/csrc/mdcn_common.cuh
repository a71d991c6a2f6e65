// Geometry shared by the modulated-deformable-conv kernels.
#pragma once
#include "common.cuh"

namespace aanet {

struct MdcnDims {
    int B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg;
    int Ho, Wo, K, Cg, Og, Cd;   // derived: taps, channels per conv group / out per group / per dg
    long P, HW;                  // output / input pixels per image
};

// Validates like deform_conv_cuda.cpp:497-516 and fills the derived fields.
inline int mdcn_make_dims(MdcnDims &d, int B, int Cin, int H, int W, int Cout, int kh, int kw,
                          int stride, int pad, int dil, int groups, int dg) {
    if (B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 ||
        pad < 0 || dil <= 0 || groups <= 0 || dg <= 0)
        return AANET_ERR_SHAPE;
    if (Cin % groups || Cout % groups || Cin % dg) return AANET_ERR_SHAPE;
    d = MdcnDims{B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg};
    d.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    d.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (d.Ho <= 0 || d.Wo <= 0) return AANET_ERR_SHAPE;
    d.K = kh * kw; d.Cg = Cin / groups; d.Og = Cout / groups; d.Cd = Cin / dg;
    d.P = (long)d.Ho * d.Wo; d.HW = (long)H * W;
    if (B > 65535 || d.K > 64) return AANET_ERR_UNSUPPORTED;
    return AANET_OK;
}

// One bilinear sampling point: four clamped (always in-bounds) plane indices and four weights.
// A corner outside the image, or a point failing the (-1,H)x(-1,W) test, gets weight 0
// (reference deform_conv_cuda_kernel.cu:467-497, :618).
struct Sample {
    int i[4];      // (h0,w0) (h0,w1) (h1,w0) (h1,w1) as h*W+w, clamped to 0 when the corner is out
    float w[4];    // bilinear corner weights, 0 for dropped corners
    float lh, lw;  // fractional parts (needed by the coordinate gradient)
    int ok;        // bit c set when corner c lies inside the image (and the point is valid)
    bool valid;
};

__device__ __forceinline__ Sample make_sample(float h, float w, int H, int W) {
    Sample s;
    s.valid = (h > -1.f && w > -1.f && h < (float)H && w < (float)W);
    const float hf = floorf(h), wf = floorf(w);
    const int h0 = (int)hf, w0 = (int)wf;
    s.lh = h - hf; s.lw = w - wf;
    const float hh = 1.f - s.lh, hw = 1.f - s.lw;
    const bool h0ok = s.valid && h0 >= 0, h1ok = s.valid && h0 + 1 <= H - 1;
    const bool w0ok = w0 >= 0, w1ok = w0 + 1 <= W - 1;
    const bool ok0 = h0ok && w0ok, ok1 = h0ok && w1ok, ok2 = h1ok && w0ok, ok3 = h1ok && w1ok;
    s.ok = (ok0 ? 1 : 0) | (ok1 ? 2 : 0) | (ok2 ? 4 : 0) | (ok3 ? 8 : 0);
    s.i[0] = ok0 ? h0 * W + w0 : 0;
    s.i[1] = ok1 ? h0 * W + w0 + 1 : 0;
    s.i[2] = ok2 ? (h0 + 1) * W + w0 : 0;
    s.i[3] = ok3 ? (h0 + 1) * W + w0 + 1 : 0;
    s.w[0] = ok0 ? hh * hw : 0.f;
    s.w[1] = ok1 ? hh * s.lw : 0.f;
    s.w[2] = ok2 ? s.lh * hw : 0.f;
    s.w[3] = ok3 ? s.lh * s.lw : 0.f;
    return s;
}

// Sampling position of tap k for output pixel (ho, wo): cu:607-616 (int arithmetic first, then
// the float offset is added).
__device__ __forceinline__ Sample sample_at(const MdcnDims &d, const float *__restrict__ off_b,
                                            int g, int k, int ho, int wo, long p) {
    const int ki = k / d.kw, kj = k % d.kw;
    const float oh = off_b[((long)(g * d.K + k) * 2 + 0) * d.P + p];
    const float ow = off_b[((long)(g * d.K + k) * 2 + 1) * d.P + p];
    const float h = (float)(ho * d.stride - d.pad + ki * d.dil) + oh;
    const float w = (float)(wo * d.stride - d.pad + kj * d.dil) + ow;
    return make_sample(h, w, d.H, d.W);
}

}  // namespace aanet
