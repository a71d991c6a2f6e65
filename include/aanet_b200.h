/*
 * aanet_b200 -- C-ABI of the B200-native AANet hot path (libaanet_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 *   - takes raw DEVICE pointers to caller-owned, contiguous NCHW fp32 buffers,
 *   - writes its outputs completely (no pre-zero requirement, unlike the reference whose
 *     callee zeroes `output`, deform_conv_cuda.cpp:530, and whose caller passes zeros_like
 *     gradients, deform_conv.py:156-160),
 *   - never allocates, never synchronises the device, keeps no global mutable state,
 *   - enqueues its kernels on `stream` (a cudaStream_t / CUstream passed as void*; NULL is the
 *     legacy default stream) and is safe under CUDA-graph capture,
 *   - returns an aanet_status (0 = ok).  Launch failures are returned, not printf'd and
 *     swallowed as in deform_conv_cuda_kernel.cu:794-798.
 *
 * Reference citations (path:line) are relative to the upstream repository
 * wuzhongwulidong/aanet.  INTEGRATION.md shows the reference-side binding.
 */
#ifndef AANET_B200_H_
#define AANET_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AANET_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define AANET_API __attribute__((visibility("default")))
#else
#define AANET_API
#endif

enum aanet_status {
    AANET_OK = 0,
    AANET_ERR_NULL = 1,        /* a required pointer is NULL                                   */
    AANET_ERR_SHAPE = 2,       /* non-positive / inconsistent dimensions (cpp:497-516 checks)  */
    AANET_ERR_UNSUPPORTED = 3, /* legal for the reference but outside what the kernels cover   */
    AANET_ERR_WORKSPACE = 4,   /* workspace NULL or smaller than *_workspace_bytes()           */
    AANET_ERR_LAUNCH = 5       /* cudaGetLastError() != cudaSuccess after a launch             */
};

/* ABI version of the loaded library, and a static string for a status code. */
AANET_API int aanet_abi_version(void);
AANET_API const char *aanet_status_string(int status);
/* Last CUDA error string recorded by a failing launch on this host thread ("" if none). */
AANET_API const char *aanet_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------------
 * Correlation cost volume.  Replaces CostVolume.forward, correlation branch
 * (nets/cost.py:40-48), called per scale by CostVolumePyramid.forward (nets/cost.py:64-76).
 *   cost[b,d,h,w] = (1/C) * sum_c L[b,c,h,w] * R[b,c,h,w-d]   for w >= d,  exactly 0 for w < d
 * L, R: [B,C,H,W]   cost: [B,D,H,W]
 * ------------------------------------------------------------------------------------------- */
AANET_API int aanet_corr_fwd(const float *L, const float *R, float *cost,
                   int B, int C, int H, int W, int D, void *stream);

/* Same volume written channels-last, cost: [B,H,W,D] (D % 4 == 0, D <= 128), for the fused aggregation executor:
 * its first 1x1 convolutions read channels-last activations.  AANET_ERR_UNSUPPORTED otherwise. */
AANET_API int aanet_corr_fwd_nhwc(const float *L, const float *R, float *cost, int B, int C, int H, int W, int D,
                                  void *stream);

/* bf16-feature variant (BASELINE config 5): L, R are bfloat16 [B,C,H,W]; products, channel sum and the
 * volume stay fp32.  Not bit-comparable with the reference (features are rounded to 8 mantissa bits); the
 * tests state the resulting end-point error.  Requires W % 8 == 0 (else AANET_ERR_UNSUPPORTED).  Inference
 * only. */
AANET_API int aanet_corr_fwd_bf16(const void *L, const void *R, float *cost,
                                  int B, int C, int H, int W, int D, void *stream);

/* Gradient of the above (the reference gets it from autograd of cost.py:45-48).
 * gcost: [B,D,H,W]   gL, gR: [B,C,H,W] (overwritten) */
AANET_API int aanet_corr_bwd(const float *L, const float *R, const float *gcost, float *gL, float *gR,
                   int B, int C, int H, int W, int D, void *stream);

/* ---------------------------------------------------------------------------------------------
 * 5-D cost volumes of the non-correlation variants.  Replaces CostVolume.forward, 'difference' and 'concat'
 * branches (nets/cost.py:22-38; used by the StereoNet / PSMNet / GC-Net style models of nets/aanet.py).
 *   mode 0 (difference): out[b,c,d,h,w] = L[b,c,h,w] - R[b,c,h,w-d]                       out: [B,C,D,H,W]
 *   mode 1 (concat)    : out[b,c,d,h,w] = L[b,c,h,w], out[b,C+c,d,h,w] = R[b,c,h,w-d]     out: [B,2C,D,H,W]
 * for w >= d, exactly 0 for w < d.  Bit-identical to the reference (one subtraction / a copy per element).
 * aanet_cost5d_bwd: gL, gR [B,C,H,W] (overwritten) from gout of the forward's shape.
 * ------------------------------------------------------------------------------------------- */
AANET_API int aanet_cost5d_fwd(const float *L, const float *R, float *out, int B, int C, int H, int W, int D,
                               int mode, void *stream);
AANET_API int aanet_cost5d_bwd(const float *gout, float *gL, float *gR, int B, int C, int H, int W, int D, int mode,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * Soft-argmin disparity regression.  Replaces DisparityEstimation.forward
 * (nets/estimation.py:13-30):  disp[b,h,w] = sum_d d * softmax_d(sign * cost[b,:,h,w]),
 * sign = +1 when `similarity` != 0 (match_similarity=True, aanet.py:113), else -1.
 * cost: [B,D,H,W]   disp: [B,H,W]
 * ------------------------------------------------------------------------------------------- */
AANET_API int aanet_softargmin_fwd(const float *cost, float *disp,
                         int B, int D, int H, int W, int similarity, void *stream);

/* gcost[b,d,h,w] = sign * gdisp[b,h,w] * p_d * (d - disp[b,h,w])  (overwritten) */
AANET_API int aanet_softargmin_bwd(const float *cost, const float *gdisp, float *gcost,
                         int B, int D, int H, int W, int similarity, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Modulated deformable convolution (DCNv2), the ISA operator.  Replaces the pybind entry points
 * modulated_deform_conv_cuda_forward / _backward (nets/deform_conv/src/deform_conv_cuda.cpp:490-
 * 569, :571-685, bound at :687-701) and the three kernels behind them
 * (deform_conv_cuda_kernel.cu:570-633, :635-693, :695-767).  With mask == NULL it is DCNv1
 * (deform_conv_forward_cuda & co., cpp:152-488; mask == 1).
 *
 *   x      [B, Cin, H, W]
 *   offset [B, dg*2*kh*kw, Ho, Wo]   channel (g*kh*kw + k)*2 + {0: dh, 1: dw}
 *   mask   [B, dg*kh*kw,   Ho, Wo]   or NULL
 *   weight [Cout, Cin/groups, kh, kw]
 *   bias   [Cout] or NULL (the reference passes a fake 1-element tensor, deform_conv.py:133)
 *   out    [B, Cout, Ho, Wo],  Ho = (H + 2*pad - (dil*(kh-1)+1))/stride + 1, same for Wo
 * A sampling point contributes 0 unless -1 < h < H and -1 < w < W; each of its four corners is
 * zero-padded individually (cu:467-497, :618).
 *
 * Epilogue extension (not in the reference op; used by the drop-in's fused inference path,
 * SURVEY.md 8f rank 1): when post_scale/post_shift are non-NULL the result is
 * out = act(out * post_scale[o] + post_shift[o]) with act = ReLU if relu != 0.  Pass
 * NULL, NULL, 0 for the reference semantics.
 *
 * Workspace: query with aanet_mdcn_workspace_bytes().  Forward: the workspace receives the tf32
 * hi/lo-split, swizzled weights that the tcgen05 kernel streams and a channels-last copy of x
 * (the engine gathers 128-byte channel rows); passing ws == NULL is legal and
 * selects the shape-generic FFMA kernel instead (same results within fp32 rounding).  Backward: the
 * workspace holds grad_weight partials and is required.
 * ------------------------------------------------------------------------------------------- */
AANET_API size_t aanet_mdcn_workspace_bytes(int backward, int B, int Cin, int H, int W, int Cout,
                                  int kh, int kw, int stride, int pad, int dil,
                                  int groups, int dg);

AANET_API int aanet_mdcn_fwd(const float *x, const float *offset, const float *mask,
                   const float *weight, const float *bias, float *out,
                   int B, int Cin, int H, int W, int Cout, int kh, int kw,
                   int stride, int pad, int dil, int groups, int dg,
                   const float *post_scale, const float *post_shift, int relu,
                   void *ws, size_t ws_bytes, void *stream);

/* All five gradients are overwritten.  gmask may be NULL iff mask is NULL; gbias may be NULL.
 * gx's summation order is not fixed (float atomics, as in the reference, cu:688). */
AANET_API int aanet_mdcn_bwd(const float *x, const float *offset, const float *mask,
                   const float *weight, const float *gout,
                   float *gx, float *goffset, float *gmask, float *gweight, float *gbias,
                   int B, int Cin, int H, int W, int Cout, int kh, int kw,
                   int stride, int pad, int dil, int groups, int dg,
                   void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dense convolution with fused epilogue on the same tcgen05 engine (the
 * 1x1 / 3x3 / strided / dilated / grouped convolutions of the ISA block and the CSA exchange paths,
 * nets/deform.py:216-236, nets/aggregation.py:346-371, :443-450, which the reference runs as
 * cuDNN conv + BatchNorm + activation kernels).
 *   out = act( (conv(x, weight) + bias[o]) * scale[o] + shift[o] + residual )
 * bias, scale/shift (both or neither), residual ([B,Cout,Ho,Wo]) may be NULL.
 * act: 0 none, 1 ReLU, 2 LeakyReLU(slope).  pad is explicit (not derived from dil).
 * Requires Cin/groups % 4 == 0 (else AANET_ERR_UNSUPPORTED) and the workspace.
 * ------------------------------------------------------------------------------------------- */
AANET_API size_t aanet_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw,
                                              int stride, int pad, int dil, int groups);

AANET_API int aanet_conv2d_fwd(const float *x, const float *weight, const float *bias,
                               const float *scale, const float *shift, const float *residual,
                               int act, float slope, float *out,
                               int B, int Cin, int H, int W, int Cout, int kh, int kw,
                               int stride, int pad, int dil, int groups,
                               void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Channels-last engine calls (used by the drop-in's fused inference path; no counterpart in the
 * reference, which keeps NCHW throughout).  Activations are [B][H*W][C]; weights are pre-packed once
 * (tf32 hi/lo split + 128-byte swizzle) with aanet_conv_pack_weights into a buffer of
 * aanet_conv_wpack_bytes() bytes and reused until the weights change.
 *
 * dense problem:  out = act((conv(x, w) + bias) * scale + shift + residual), out/residual channels-last
 *   (out_nchw == 0) or NCHW (out_nchw != 0).  act 0-2 as above; act == 3 is the DeformConv2d offset/mask
 *   head (nets/deform.py:80-89): channels >= n_offset_ch get mask_scale * sigmoid(.), the rest pass through.
 * deform problem:  DCNv2 with x channels-last and offsets+mask in ONE channels-last tensor
 *   offmask [B][Ho*Wo][om_channels], offsets in channels [0, dg*2*kh*kw) and the (already activated) mask
 *   behind them, i.e. the positional split of nets/deform.py:82-85; om_channels == dg*2*kh*kw means no
 *   mask (DCNv1).
 * ------------------------------------------------------------------------------------------- */
/* bn: N-tile width the weights are packed for -- 0 = the layer's own (multiple of 16, <= 64), or the width
 * of the batch the layer will run in (aanet_conv_batch_nhwc). */
AANET_API size_t aanet_conv_wpack_bytes(int Cout, int Cin, int kh, int kw, int groups, int bn);
AANET_API int aanet_conv_pack_weights(const float *weight, void *wpack, int Cout, int Cin, int kh, int kw,
                                      int groups, int bn, void *stream);
AANET_API int aanet_nchw_to_nhwc(const float *src, float *dst, int B, int C, int HW, void *stream);
AANET_API int aanet_nhwc_to_nchw(const float *src, float *dst, int B, int C, int HW, void *stream);

/* One convolution problem of a batch.  DENSE: offmask/om_channels/dg ignored.  DEFORM: act must be 0 or 1. */
typedef struct aanet_conv_desc {
    const float *x;          /* [B][H*W][Cin] */
    const void *wpack;       /* aanet_conv_pack_weights(..., bn) */
    const float *bias, *scale, *shift, *residual;   /* optional; residual has out's layout */
    float *out;              /* [B][Ho*Wo][Cout], or NCHW when out_nchw != 0 */
    const float *offmask;    /* DEFORM: [B][Ho*Wo][om_channels] */
    int om_channels;
    int B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg;
    int act;                 /* 0 none, 1 ReLU, 2 LeakyReLU(slope), 3 offset/mask head, 4 soft-argmin over the output
                              * channels (nets/estimation.py:19-28 fused into the final 1x1, nets/aggregation.py:443-450):
                              * `out` is then [B][Ho*Wo], one disparity per pixel; dense, groups == 1, Cout <= 64 */
    float slope;
    int n_offset_ch;         /* act == 3 */
    float mask_scale;        /* act == 3 */
    int out_nchw;
    int om_nchw;             /* DEFORM: offmask is [B][om_channels][Ho*Wo] (channel planes) instead of channels-last */
    /* Optional fused 1x1 tail (tail_wpack != NULL), the trailing conv3 + bn3 + identity + ReLU of a bottleneck
     * (nets/deform.py:177-183, :229-235) inside the same launch:
     *   out = tail_act( conv1x1(act(main result), tail weights) * tail_scale + tail_shift + tail_residual )
     * `out` and `tail_residual` are then [B][Ho*Wo][tail_cout]; bias/scale/shift/act describe the main convolution,
     * whose result never leaves the chip.  Only problems for which aanet_conv_tail_supported() returns 1. */
    const void *tail_wpack;  /* aanet_conv_pack_weights of the [tail_cout, Cout, 1, 1] weight */
    const float *tail_scale, *tail_shift, *tail_residual;
    int tail_cout, tail_act;
} aanet_conv_desc;

#define AANET_CONV_MAX_BATCH 3

/* Runs 1..3 problems of the same kind (deform != 0: all DCN, else all dense) as ONE persistent kernel whose
 * tile list spans all of them -- the three pyramid scales of one aggregation stage.  All problems use the
 * same N-tile width bn (0 = the widest natural width among them); their weights must be packed for it. */
AANET_API int aanet_conv_batch_nhwc(const aanet_conv_desc *descs, int n, int deform, int bn, void *stream);
/* 1 when `desc` (with its tail fields set) can run as ONE launch of the tensor-memory kernels, else 0: the caller
 * then issues the main convolution and the 1x1 as two problems. */
AANET_API int aanet_conv_tail_supported(const aanet_conv_desc *desc, int deform);
/* Channels-last twin of aanet_csa_fuse_fwd: terms and out are [B][h][w][C], C % 4 == 0. */
AANET_API int aanet_csa_fuse_nhwc(const float *const *terms, const int *th, const int *tw, int n_terms,
                                  float *out, int B, int C, int H, int W, float slope, void *stream);
/* aanet_csa_fuse_nhwc fused with the 1x1 convolution that follows it -- the tail of AdaptiveAggregationModule.forward
 * (nets/aggregation.py:387-400) and conv1 + bn1 + ReLU of the next module's bottleneck (nets/deform.py:164-170,
 * :216-222) as ONE launch:
 *   fused_out = LeakyReLU_slope( sum_k resize(terms[k]) )                       [B][H*W][C], written once; may be
 *                                                                               NULL (only the convolution needs it)
 *   out       = act( (conv1x1(fused_out, wpack) + bias) * scale + shift )       [B][H*W][Cout]
 * terms[k] is (H, W)-sized or smaller in both axes (bilinear, align_corners = False); C % 32 == 0, Cout 32 or 64,
 * wpack from aanet_conv_pack_weights (kh = kw = 1); bias / scale+shift optional.  act: 0 none, 1 ReLU,
 * 2 LeakyReLU(slope), 4 soft-argmin over the Cout channels (`out` is then [B][H*W]: the last aggregation module's
 * sum, the final 1x1 convolution (nets/aggregation.py:443-450) and DisparityEstimation (nets/estimation.py:19-28)
 * in one launch).  AANET_ERR_UNSUPPORTED otherwise (the caller then issues aanet_csa_fuse_nhwc and the convolution). */
AANET_API int aanet_csa_conv1_nhwc(const float *const *terms, const int *th, const int *tw, int n_terms, float slope,
                                   float *fused_out, const void *wpack, const float *bias, const float *scale,
                                   const float *shift, int act, float *out, int B, int C, int Cout, int H, int W,
                                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-scale aggregation fuse.  Replaces the resize + sum + LeakyReLU tail of
 * AdaptiveAggregationModule.forward (nets/aggregation.py:387-400):
 *   out = LeakyReLU_slope( (t0 + r(t1)) + r(t2) ... ),  r = bilinear resize to (H,W) with
 *   align_corners=False when the term's (th,tw) differs from (H,W) (:394-396), identity
 *   otherwise.  Summation order is the reference's (j = 0, 1, 2).
 * terms / gterms: HOST arrays of n_terms (<= AANET_CSA_MAX_TERMS) DEVICE pointers, term t being
 * [B, C, th[t], tw[t]];  th, tw: HOST int arrays.   out: [B,C,H,W].
 * ------------------------------------------------------------------------------------------- */
#define AANET_CSA_MAX_TERMS 4

AANET_API int aanet_csa_fuse_fwd(const float *const *terms, const int *th, const int *tw, int n_terms,
                       float *out, int B, int C, int H, int W, float slope, void *stream);

/* gterms[t] (NULL = not wanted) = adjoint of term t's path applied to gout * LeakyReLU'(pre-activation); the sign of
 * the pre-activation is read from `out` (slope must be > 0).  Deterministic (gather form). */
AANET_API int aanet_csa_fuse_bwd(const float *out, const float *gout, float *const *gterms,
                       const int *th, const int *tw, int n_terms,
                       int B, int C, int H, int W, float slope, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Refinement front end (widening step).  Replaces the lines of
 * StereoDRNetRefinement / HourglassRefinement.forward before the first convolution
 * (nets/refinement.py:80-95, :144-160) and disp_warp (nets/warp.py:41-64):
 *   disp   = bilinear(low_disp, (H,W), align_corners=False) * (W/w)    (low_disp itself when W == w)
 *   concat = cat(grid_sample(right, (x - disp, y), bilinear, border, align_corners=True) - left, left)
 * low_disp: [B,h,w]   left, right: [B,C,H,W]   concat: [B,2C,H,W]   disp: [B,1,H,W]   (all overwritten
 * outputs; the reference's `assert disp.min() >= 0` device synchronisation (warp.py:51) is not performed).
 * ------------------------------------------------------------------------------------------- */
AANET_API int aanet_refine_frontend_fwd(const float *low_disp, const float *left, const float *right, float *concat,
                                        float *disp, int B, int C, int h, int w, int H, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AANET_B200_H_ */
