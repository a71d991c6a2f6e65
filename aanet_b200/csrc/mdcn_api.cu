// C-ABI entry points of the convolution engine and the modulated deformable convolution
// (include/aanet_b200.h): reference-shaped NCHW calls and the channels-last calls the fused inference
// path uses.
#include "conv_engine.cuh"

namespace aanet {
// mdcn_fwd.cu / mdcn_bwd.cu
int mdcn_fwd_generic(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, const MdcnDims &d, const float *post_scale,
                     const float *post_shift, int relu, cudaStream_t stream);
size_t mdcn_bwd_workspace_bytes(const MdcnDims &d);
int mdcn_bwd_launch(const float *x, const float *offset, const float *mask, const float *weight,
                    const float *gout, float *gx, float *goffset, float *gmask, float *gweight,
                    float *gbias, const MdcnDims &d, void *ws, size_t ws_bytes, cudaStream_t stream);

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
    return conv_umma_supported(d, true) ? align256(conv_umma_wpack_bytes(d, 0)) + nhwc_bytes(d) : 0;
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
        const size_t wbytes = align256(conv_umma_wpack_bytes(d, 0));
        if (ws_bytes < wbytes + nhwc_bytes(d)) return AANET_ERR_WORKSPACE;
        float *xt = reinterpret_cast<float *>(static_cast<char *>(ws) + wbytes);
        int prc = conv_umma_pack(weight, ws, d, 0, as_stream(stream));
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
    return conv_umma_supported(d, false) ? align256(conv_umma_wpack_bytes(d, 0)) + nhwc_bytes(d) : 0;
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
    const size_t wbytes = align256(conv_umma_wpack_bytes(d, 0));
    if (!ws || !aligned16(ws) || ws_bytes < wbytes + nhwc_bytes(d)) return AANET_ERR_WORKSPACE;
    float *xt = reinterpret_cast<float *>(static_cast<char *>(ws) + wbytes);
    int prc = conv_umma_pack(weight, ws, d, 0, as_stream(stream));
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
extern "C" size_t aanet_conv_wpack_bytes(int Cout, int Cin, int kh, int kw, int groups, int bn) {
    MdcnDims d;
    if (mdcn_make_dims(d, 1, Cin, kh, kw, Cout, kh, kw, 1, 0, 1, groups, 1) != AANET_OK) return 0;
    return conv_umma_supported(d, false) ? conv_umma_wpack_bytes(d, bn) : 0;
}

extern "C" int aanet_conv_pack_weights(const float *weight, void *wpack, int Cout, int Cin, int kh, int kw,
                                       int groups, int bn, void *stream) {
    if (!weight || !wpack) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, 1, Cin, kh, kw, Cout, kh, kw, 1, 0, 1, groups, 1);
    if (rc) return rc;
    if (!conv_umma_supported(d, false)) return AANET_ERR_UNSUPPORTED;
    return conv_umma_pack(weight, wpack, d, bn, as_stream(stream));
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

// Fill a ConvParams from a descriptor; returns an aanet_status.
static int problem_from_desc(const aanet_conv_desc &c, bool deform, ConvParams &p) {
    if (!c.x || !c.wpack || !c.out) return AANET_ERR_NULL;
    if ((c.scale == nullptr) != (c.shift == nullptr)) return AANET_ERR_NULL;
    if (c.act < ACT_NONE || c.act > ACT_SOFTARGMIN) return AANET_ERR_UNSUPPORTED;
    // soft-argmin epilogue: every candidate in one N tile of a dense, ungrouped problem; `out` is [B][P]
    if (c.act == ACT_SOFTARGMIN && (deform || c.groups != 1 || c.Cout > 64 || c.residual || c.tail_wpack)) return AANET_ERR_UNSUPPORTED;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, c.B, c.Cin, c.H, c.W, c.Cout, c.kh, c.kw, c.stride, c.pad, c.dil, c.groups,
                                  deform ? c.dg : 1);
    if (rc) return rc;
    if (!conv_umma_supported(d, deform) || !aligned16(c.x) || !aligned16(c.wpack) || !aligned16(c.out))
        return AANET_ERR_UNSUPPORTED;
    p = ConvParams{};
    p.x = c.x; p.wpack = static_cast<const float *>(c.wpack); p.out = c.out; p.out_nchw = c.out_nchw ? 1 : 0;
    p.bias = c.bias; p.scale = c.scale; p.shift = c.shift; p.residual = c.residual;
    p.act = c.act; p.slope = c.slope; p.n_offset_ch = c.n_offset_ch; p.mask_scale = c.mask_scale;
    p.d = d;
    if (c.tail_wpack) {
        if (c.tail_cout <= 0 || (c.tail_scale == nullptr) != (c.tail_shift == nullptr)) return AANET_ERR_SHAPE;
        p.tail_wpack = static_cast<const float *>(c.tail_wpack); p.tail_scale = c.tail_scale; p.tail_shift = c.tail_shift;
        p.tail_residual = c.tail_residual; p.tail_cout = c.tail_cout; p.tail_act = c.tail_act;
    }
    if (deform) {
        if (!c.offmask) return AANET_ERR_NULL;
        const int n_off = c.dg * 2 * d.K, n_mask = c.dg * d.K;
        if (c.om_channels != n_off && c.om_channels != n_off + n_mask) return AANET_ERR_SHAPE;
        p.offset = c.offmask; p.off_bs = (long)d.P * c.om_channels;
        if (c.om_nchw) {        // channel planes: consecutive pixels are consecutive addresses
            p.off_ps = 1; p.off_cs = d.P;
            p.mask = (c.om_channels == n_off) ? nullptr : c.offmask + (long)n_off * d.P;
        } else {
            p.off_ps = c.om_channels; p.off_cs = 1;
            p.mask = (c.om_channels == n_off) ? nullptr : c.offmask + n_off;
        }
        p.mask_bs = p.off_bs; p.mask_ps = p.off_ps; p.mask_cs = p.off_cs;
    }
    return AANET_OK;
}

extern "C" int aanet_conv_tail_supported(const aanet_conv_desc *desc, int deform) {
    if (!desc || !desc->tail_wpack) return 0;
    ConvParams p;
    if (problem_from_desc(*desc, deform != 0, p) != AANET_OK) return 0;
    return tmem_tail_supported(p, deform != 0) ? 1 : 0;
}

extern "C" int aanet_conv_batch_nhwc(const aanet_conv_desc *descs, int n, int deform, int bn, void *stream) {
    if (!descs) return AANET_ERR_NULL;
    if (n < 1 || n > AANET_CONV_MAX_BATCH) return AANET_ERR_SHAPE;
    ConvParams probs[AANET_CONV_MAX_BATCH];
    for (int i = 0; i < n; ++i) {
        const int rc = problem_from_desc(descs[i], deform != 0, probs[i]);
        if (rc) return rc;
        // a fused tail only exists in the tensor-memory kernels (single problem)
        if (probs[i].tail_wpack && (n != 1 || !tmem_tail_supported(probs[i], deform != 0))) return AANET_ERR_UNSUPPORTED;
    }
    return conv_umma_launch_batch(probs, n, deform != 0, bn, as_stream(stream));
}

extern "C" int aanet_csa_conv1_nhwc(const float *const *terms, const int *th, const int *tw, int n_terms, float slope,
                                    float *fused_out, const void *wpack, const float *bias, const float *scale,
                                    const float *shift, int act, float *out, int B, int C, int Cout, int H, int W,
                                    void *stream) {
    if (!terms || !th || !tw || !wpack || !out) return AANET_ERR_NULL;
    if ((scale == nullptr) != (shift == nullptr)) return AANET_ERR_NULL;
    if (n_terms < 1 || n_terms > AANET_CSA_MAX_TERMS) return AANET_ERR_SHAPE;
    if ((act < ACT_NONE || act > ACT_LEAKY) && act != ACT_SOFTARGMIN) return AANET_ERR_UNSUPPORTED;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, C, H, W, Cout, 1, 1, 1, 0, 1, 1, 1);
    if (rc) return rc;
    if (!aligned16(wpack) || (act != ACT_SOFTARGMIN && !aligned16(out))) return AANET_ERR_UNSUPPORTED;
    ConvParams p{};
    p.wpack = static_cast<const float *>(wpack); p.out = out; p.bias = bias; p.scale = scale; p.shift = shift;
    p.act = act; p.slope = slope; p.mask_scale = 1.f;
    p.d = d;
    return csa_conv1_tmem_launch(terms, th, tw, n_terms, slope, fused_out, p, as_stream(stream));
}
