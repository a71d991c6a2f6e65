// C-ABI entry points of the modulated deformable convolution (include/aanet_b200.h).
#include "mdcn_common.cuh"

namespace aanet {
int mdcn_fwd_generic(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, const MdcnDims &d, const float *post_scale,
                     const float *post_shift, int relu, cudaStream_t stream);
size_t mdcn_bwd_workspace_bytes(const MdcnDims &d);
int mdcn_bwd_launch(const float *x, const float *offset, const float *mask, const float *weight,
                    const float *gout, float *gx, float *goffset, float *gmask, float *gweight,
                    float *gbias, const MdcnDims &d, void *ws, size_t ws_bytes, cudaStream_t stream);
// conv_umma.cu
enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_OFFSET_MASK = 3 };
struct ConvParams {
    const float *x, *offset, *mask;
    const float *wpack;
    float *out;
    const float *bias, *scale, *shift, *residual;
    int act; float slope; int n_offset_ch;
    float mask_scale;
    MdcnDims d;
    int K, KB, n_tiles_n, tiles_per_img;
};
bool conv_umma_supported(const MdcnDims &d, bool deform);
size_t conv_umma_wpack_bytes(const MdcnDims &d);
int conv_umma_pack(const float *weight, void *wpack, const MdcnDims &d, cudaStream_t stream);
int conv_umma_launch(ConvParams p, bool deform, cudaStream_t stream);
}  // namespace aanet

using namespace aanet;

extern "C" size_t aanet_mdcn_workspace_bytes(int backward, int B, int Cin, int H, int W, int Cout, int kh,
                                             int kw, int stride, int pad, int dil, int groups, int dg) {
    MdcnDims d;
    if (mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg) != AANET_OK) return 0;
    if (backward) return mdcn_bwd_workspace_bytes(d);
    return conv_umma_supported(d, true) ? conv_umma_wpack_bytes(d) : 0;
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
    // tcgen05 path: needs the packed-weight workspace.  ws == NULL selects the shape-generic FFMA kernel.
    if (ws && conv_umma_supported(d, true)) {
        if (ws_bytes < conv_umma_wpack_bytes(d)) return AANET_ERR_WORKSPACE;
        int prc = conv_umma_pack(weight, ws, d, as_stream(stream));
        if (prc) return prc;
        ConvParams p{};
        p.x = x; p.offset = offset; p.mask = mask; p.wpack = static_cast<const float *>(ws); p.out = out;
        p.bias = bias; p.scale = post_scale; p.shift = post_shift; p.residual = nullptr;
        p.act = relu ? ACT_RELU : ACT_NONE; p.slope = 0.f; p.n_offset_ch = 0; p.mask_scale = 1.f;
        p.d = d;
        return conv_umma_launch(p, true, as_stream(stream));
    }
    return mdcn_fwd_generic(x, offset, mask, weight, bias, out, d, post_scale, post_shift, relu,
                            as_stream(stream));
}

extern "C" size_t aanet_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw,
                                               int stride, int pad, int dil, int groups) {
    MdcnDims d;
    if (mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, 1) != AANET_OK) return 0;
    return conv_umma_supported(d, false) ? conv_umma_wpack_bytes(d) : 0;
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
    if (!ws || ws_bytes < conv_umma_wpack_bytes(d)) return AANET_ERR_WORKSPACE;
    int prc = conv_umma_pack(weight, ws, d, as_stream(stream));
    if (prc) return prc;
    ConvParams p{};
    p.x = x; p.offset = nullptr; p.mask = nullptr; p.wpack = static_cast<const float *>(ws); p.out = out;
    p.bias = bias; p.scale = scale; p.shift = shift; p.residual = residual;
    p.act = act; p.slope = slope; p.n_offset_ch = 0; p.mask_scale = 1.f;
    p.d = d;
    return conv_umma_launch(p, false, as_stream(stream));
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
