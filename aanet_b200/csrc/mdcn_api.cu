// C-ABI entry points of the convolution engine and the modulated deformable convolution
// (include/aanet_b200.h): reference-shaped NCHW calls and the channels-last calls the fused inference
// path uses.
#include "mdcn_common.cuh"

namespace aanet {
// mdcn_fwd.cu / mdcn_bwd.cu
int mdcn_fwd_generic(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, const MdcnDims &d, const float *post_scale,
                     const float *post_shift, int relu, cudaStream_t stream);
size_t mdcn_bwd_workspace_bytes(const MdcnDims &d);
int mdcn_bwd_launch(const float *x, const float *offset, const float *mask, const float *weight,
                    const float *gout, float *gx, float *goffset, float *gmask, float *gweight,
                    float *gbias, const MdcnDims &d, void *ws, size_t ws_bytes, cudaStream_t stream);

// conv_umma.cu (keep in sync)
enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_OFFSET_MASK = 3 };
struct ConvParams {
    const float *x;
    const float *offset, *mask;
    long off_bs, off_ps, off_cs;
    long mask_bs, mask_ps, mask_cs;
    const float *wpack;
    float *out;
    int out_nchw;
    const float *bias, *scale, *shift;
    const float *residual;
    int act; float slope; int n_offset_ch; float mask_scale;
    MdcnDims d;
    int K, KB, n_tiles_n, tiles_per_img, n_ptiles, total_tiles;
};
bool conv_umma_supported(const MdcnDims &d, bool deform);
size_t conv_umma_wpack_bytes(const MdcnDims &d);
int conv_umma_pack(const float *weight, void *wpack, const MdcnDims &d, cudaStream_t stream);
int conv_umma_transpose(const float *src, float *dst, int B, int R, long Cc, cudaStream_t stream);
int conv_umma_launch(ConvParams p, bool deform, cudaStream_t stream);

static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static inline size_t nhwc_bytes(const MdcnDims &d) { return (size_t)d.B * d.Cin * d.HW * sizeof(float); }
}  // namespace aanet

using namespace aanet;

// ------------------------------------------------------------------------------ reference-shaped (NCHW)
extern "C" size_t aanet_mdcn_workspace_bytes(int backward, int B, int Cin, int H, int W, int Cout, int kh,
                                             int kw, int stride, int pad, int dil, int groups, int dg) {
    MdcnDims d;
    if (mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg) != AANET_OK) return 0;
    if (backward) return mdcn_bwd_workspace_bytes(d);
    return conv_umma_supported(d, true) ? align256(conv_umma_wpack_bytes(d)) + nhwc_bytes(d) : 0;
}

extern "C" int aanet_mdcn_fwd(const float *x, const float *offset, const float *mask, const float *weight,
                              const float *bias, float *out, int B, int Cin, int H, int W, int Cout, int kh,
                              int kw, int stride, int pad, int dil, int groups, int dg,
                              const float *post_scale, const float *post_shift, int relu, void *ws,
                              size_t ws_bytes, void *stream) {
    if (!x || !offset || !weight || !out) return AANET_ERR_NULL;
    if ((post_scale == nullptr) != (post_shift == nullptr)) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    // tcgen05 path: needs the workspace (packed weights + channels-last copy of x).  ws == NULL selects the
    // shape-generic FFMA kernel.
    if (ws && conv_umma_supported(d, true) && aligned16(ws)) {
        const size_t wbytes = align256(conv_umma_wpack_bytes(d));
        if (ws_bytes < wbytes + nhwc_bytes(d)) return AANET_ERR_WORKSPACE;
        float *xt = reinterpret_cast<float *>(static_cast<char *>(ws) + wbytes);
        int prc = conv_umma_pack(weight, ws, d, as_stream(stream));
        if (prc) return prc;
        prc = conv_umma_transpose(x, xt, d.B, d.Cin, d.HW, as_stream(stream));      // NCHW -> NHWC
        if (prc) return prc;
        ConvParams p{};
        p.x = xt;
        p.offset = offset; p.off_bs = (long)d.dg * 2 * d.K * d.P; p.off_ps = 1; p.off_cs = d.P;
        p.mask = mask; p.mask_bs = (long)d.dg * d.K * d.P; p.mask_ps = 1; p.mask_cs = d.P;
        p.wpack = static_cast<const float *>(ws); p.out = out; p.out_nchw = 1;
        p.bias = bias; p.scale = post_scale; p.shift = post_shift; p.residual = nullptr;
        p.act = relu ? ACT_RELU : ACT_NONE; p.slope = 0.f; p.n_offset_ch = 0; p.mask_scale = 1.f;
        p.d = d;
        return conv_umma_launch(p, true, as_stream(stream));
    }
    return mdcn_fwd_generic(x, offset, mask, weight, bias, out, d, post_scale, post_shift, relu,
                            as_stream(stream));
}

extern "C" int aanet_mdcn_bwd(const float *x, const float *offset, const float *mask, const float *weight,
                              const float *gout, float *gx, float *goffset, float *gmask, float *gweight,
                              float *gbias, int B, int Cin, int H, int W, int Cout, int kh, int kw,
                              int stride, int pad, int dil, int groups, int dg, void *ws, size_t ws_bytes,
                              void *stream) {
    if (!x || !offset || !weight || !gout || !gx || !goffset || !gweight) return AANET_ERR_NULL;
    if (mask && !gmask) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    return mdcn_bwd_launch(x, offset, mask, weight, gout, gx, goffset, mask ? gmask : nullptr, gweight,
                           gbias, d, ws, ws_bytes, as_stream(stream));
}

extern "C" size_t aanet_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw,
                                               int stride, int pad, int dil, int groups) {
    MdcnDims d;
    if (mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, 1) != AANET_OK) return 0;
    return conv_umma_supported(d, false) ? align256(conv_umma_wpack_bytes(d)) + nhwc_bytes(d) : 0;
}

extern "C" int aanet_conv2d_fwd(const float *x, const float *weight, const float *bias, const float *scale,
                                const float *shift, const float *residual, int act, float slope, float *out,
                                int B, int Cin, int H, int W, int Cout, int kh, int kw, int stride, int pad,
                                int dil, int groups, void *ws, size_t ws_bytes, void *stream) {
    if (!x || !weight || !out) return AANET_ERR_NULL;
    if ((scale == nullptr) != (shift == nullptr)) return AANET_ERR_NULL;
    if (act < ACT_NONE || act > ACT_LEAKY) return AANET_ERR_UNSUPPORTED;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, 1);
    if (rc) return rc;
    if (!conv_umma_supported(d, false)) return AANET_ERR_UNSUPPORTED;
    const size_t wbytes = align256(conv_umma_wpack_bytes(d));
    if (!ws || !aligned16(ws) || ws_bytes < wbytes + nhwc_bytes(d)) return AANET_ERR_WORKSPACE;
    float *xt = reinterpret_cast<float *>(static_cast<char *>(ws) + wbytes);
    int prc = conv_umma_pack(weight, ws, d, as_stream(stream));
    if (prc) return prc;
    prc = conv_umma_transpose(x, xt, d.B, d.Cin, d.HW, as_stream(stream));
    if (prc) return prc;
    ConvParams p{};
    p.x = xt; p.wpack = static_cast<const float *>(ws); p.out = out; p.out_nchw = 1;
    p.bias = bias; p.scale = scale; p.shift = shift; p.residual = residual;
    p.act = act; p.slope = slope; p.mask_scale = 1.f;
    p.d = d;
    return conv_umma_launch(p, false, as_stream(stream));
}

// ------------------------------------------------------------------------------ channels-last engine calls
extern "C" size_t aanet_conv_wpack_bytes(int Cout, int Cin, int kh, int kw, int groups) {
    MdcnDims d;
    if (mdcn_make_dims(d, 1, Cin, kh, kw, Cout, kh, kw, 1, 0, 1, groups, 1) != AANET_OK) return 0;
    return conv_umma_supported(d, false) ? conv_umma_wpack_bytes(d) : 0;
}

extern "C" int aanet_conv_pack_weights(const float *weight, void *wpack, int Cout, int Cin, int kh, int kw,
                                       int groups, void *stream) {
    if (!weight || !wpack) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, 1, Cin, kh, kw, Cout, kh, kw, 1, 0, 1, groups, 1);
    if (rc) return rc;
    if (!conv_umma_supported(d, false)) return AANET_ERR_UNSUPPORTED;
    return conv_umma_pack(weight, wpack, d, as_stream(stream));
}

extern "C" int aanet_nchw_to_nhwc(const float *src, float *dst, int B, int C, int HW, void *stream) {
    if (!src || !dst) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0 || B > 65535) return AANET_ERR_SHAPE;
    return conv_umma_transpose(src, dst, B, C, HW, as_stream(stream));
}

extern "C" int aanet_nhwc_to_nchw(const float *src, float *dst, int B, int C, int HW, void *stream) {
    if (!src || !dst) return AANET_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0 || B > 65535) return AANET_ERR_SHAPE;
    return conv_umma_transpose(src, dst, B, HW, C, as_stream(stream));
}

extern "C" int aanet_conv2d_nhwc(const float *x, const void *wpack, const float *bias, const float *scale,
                                 const float *shift, const float *residual, int act, float slope,
                                 int n_offset_ch, float mask_scale, float *out, int out_nchw, int B, int Cin,
                                 int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil,
                                 int groups, void *stream) {
    if (!x || !wpack || !out) return AANET_ERR_NULL;
    if ((scale == nullptr) != (shift == nullptr)) return AANET_ERR_NULL;
    if (act < ACT_NONE || act > ACT_OFFSET_MASK) return AANET_ERR_UNSUPPORTED;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, 1);
    if (rc) return rc;
    if (!conv_umma_supported(d, false) || !aligned16(x) || !aligned16(wpack)) return AANET_ERR_UNSUPPORTED;
    ConvParams p{};
    p.x = x; p.wpack = static_cast<const float *>(wpack); p.out = out; p.out_nchw = out_nchw ? 1 : 0;
    p.bias = bias; p.scale = scale; p.shift = shift; p.residual = residual;
    p.act = act; p.slope = slope; p.n_offset_ch = n_offset_ch; p.mask_scale = mask_scale;
    p.d = d;
    return conv_umma_launch(p, false, as_stream(stream));
}

extern "C" int aanet_mdcn_nhwc(const float *x, const float *offmask, int om_channels, const void *wpack,
                               const float *bias, const float *post_scale, const float *post_shift, int relu,
                               float *out, int out_nchw, int B, int Cin, int H, int W, int Cout, int kh, int kw,
                               int stride, int pad, int dil, int groups, int dg, void *stream) {
    if (!x || !offmask || !wpack || !out) return AANET_ERR_NULL;
    if ((post_scale == nullptr) != (post_shift == nullptr)) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
    const int n_off = dg * 2 * d.K, n_mask = dg * d.K;
    if (om_channels != n_off && om_channels != n_off + n_mask) return AANET_ERR_SHAPE;
    if (!conv_umma_supported(d, true) || !aligned16(x) || !aligned16(wpack)) return AANET_ERR_UNSUPPORTED;
    ConvParams p{};
    p.x = x;
    p.offset = offmask; p.off_bs = (long)d.P * om_channels; p.off_ps = om_channels; p.off_cs = 1;
    p.mask = (om_channels == n_off) ? nullptr : offmask + n_off;
    p.mask_bs = p.off_bs; p.mask_ps = om_channels; p.mask_cs = 1;
    p.wpack = static_cast<const float *>(wpack); p.out = out; p.out_nchw = out_nchw ? 1 : 0;
    p.bias = bias; p.scale = post_scale; p.shift = post_shift;
    p.act = relu ? ACT_RELU : ACT_NONE; p.mask_scale = 1.f;
    p.d = d;
    return conv_umma_launch(p, true, as_stream(stream));
}
