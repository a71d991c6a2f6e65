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
}  // namespace aanet

using namespace aanet;

extern "C" size_t aanet_mdcn_workspace_bytes(int backward, int B, int Cin, int H, int W, int Cout, int kh,
                                             int kw, int stride, int pad, int dil, int groups, int dg) {
    MdcnDims d;
    if (mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg) != AANET_OK) return 0;
    return backward ? mdcn_bwd_workspace_bytes(d) : 0;
}

extern "C" int aanet_mdcn_fwd(const float *x, const float *offset, const float *mask, const float *weight,
                              const float *bias, float *out, int B, int Cin, int H, int W, int Cout, int kh,
                              int kw, int stride, int pad, int dil, int groups, int dg,
                              const float *post_scale, const float *post_shift, int relu, void *ws,
                              size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    if (!x || !offset || !weight || !out) return AANET_ERR_NULL;
    if ((post_scale == nullptr) != (post_shift == nullptr)) return AANET_ERR_NULL;
    MdcnDims d;
    const int rc = mdcn_make_dims(d, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg);
    if (rc) return rc;
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
