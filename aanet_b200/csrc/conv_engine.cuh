// Descriptor types and host entry points of the tcgen05 convolution engine (conv_umma.cu), shared with the
// C-ABI layer (mdcn_api.cu).
#pragma once
#include "mdcn_common.cuh"

namespace aanet {

// ACT_SOFTARGMIN: the epilogue reduces the tile's output channels (= disparity candidates, one N tile) to the
// soft-argmin disparity sum_d d * softmax(out)[d] (nets/estimation.py:19-28) and writes ONE float per pixel to `out`
// ([B][P]): the final 1x1 convolution of the aggregation and DisparityEstimation in one launch.
enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_OFFSET_MASK = 3, ACT_SOFTARGMIN = 4 };

constexpr int kMaxProblems = 3;   // == AANET_CONV_MAX_BATCH

// One convolution problem.  A launch (ConvBatch) walks the tiles of up to kMaxProblems problems of the same
// kind (all DENSE or all DEFORM, same N-tile width): the three pyramid scales of one aggregation stage.
struct ConvParams {
    const float *x;                     // channels-last input [B][H*W][Cin]
    const float *offset, *mask;         // DEFORM only; mask may be NULL (DCNv1)
    long off_bs, off_ps, off_cs;        // offset strides in floats: batch, pixel, channel
    long mask_bs, mask_ps, mask_cs;
    const float *wpack;                 // packed weights, see conv_pack_weights_kernel
    float *out;                         // [B][P][Cout] (out_nchw == 0) or [B][Cout][P]
    int out_nchw;
    const float *bias, *scale, *shift;  // per output channel, optional
    const float *residual;              // same layout as out, optional
    int act; float slope; int n_offset_ch; float mask_scale;
    // Optional fused tail (TMEM kernels only): out = act_t(conv1x1(act(main)) * tail_scale + tail_shift + tail_residual),
    // `out` / `tail_residual` then are [B][P][tail_cout]; scale/shift/act above apply to the main convolution.
    const float *tail_wpack, *tail_scale, *tail_shift, *tail_residual;
    int tail_cout, tail_act;
    MdcnDims d;
    int K, KB;                          // K = kh*kw*Cg, KB = ceil(K / 32)
    int n_tiles_n;                      // ceil(Og / BN)
    int tile2d, tiles_x;                // 2-D tiles (kTileW x kTileH pixels) and their count per image row
    int tiles_per_img;                  // ceil(P / 128), or tiles_x * ceil(Ho / kTileH)
    int n_ptiles;                       // B * tiles_per_img
    int total_tiles;                    // groups * n_tiles_n * n_ptiles
    int tile_start;                     // first tile index of this problem in the batch's tile list
    int tbl_off;                        // first K-block row of this problem in the chunk tables
};

struct ConvBatch {
    ConvParams pr[kMaxProblems];
    int n;
    int total_tiles;
};

// TMEM geometry of an N tile with the stacked [B_hi | B_lo] operand (2 * BN accumulator columns, two accumulators).
template <int BN> struct EngineCfgLite {
    static constexpr int kAccCols = 2 * BN;
    static constexpr int kAccStride = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : 128;
    static constexpr uint32_t kTmemCols = 2 * kAccStride;
};

// halo_engine.cu: TMA-halo kernels for stride-1 dense convolutions (AANET_ERR_UNSUPPORTED = take the gather engine)
int conv_halo_launch(const ConvParams &p, int BN, cudaStream_t stream);
// deform_halo.cu: DCNv2 with the bilinear gather served from a TMA-staged halo (same convention)
int deform_halo_launch(const ConvParams &p, int BN, cudaStream_t stream);
// deform_tmem.cu: the same with the sampled operand written straight into tensor memory (tcgen05.st, TMEM-A MMA)
int deform_tmem_launch(const ConvParams &p, int BN, cudaStream_t stream);
int dense_tmem_launch(const ConvParams &p, int BN, cudaStream_t stream);
// can the (main + fused 1x1 tail) problem run as one TMEM-kernel launch?
bool tmem_tail_supported(const ConvParams &p, bool deform);
// CSA resize-and-sum + LeakyReLU as the A producer of the following 1x1 convolution (deform_tmem.cu, FUSE)
int csa_conv1_tmem_launch(const float *const *terms, const int *th, const int *tw, int n_terms, float slope,
                          float *fused_out, const ConvParams &conv, cudaStream_t stream);

bool conv_umma_supported(const MdcnDims &d, bool deform);
int conv_umma_pick_bn(int Og);
size_t conv_umma_wpack_bytes(const MdcnDims &d, int bn);
int conv_umma_pack(const float *weight, void *wpack, const MdcnDims &d, int bn, cudaStream_t stream);
int conv_umma_transpose(const float *src, float *dst, int B, int R, long Cc, cudaStream_t stream);
int conv_umma_launch(ConvParams p, bool deform, cudaStream_t stream);
int conv_umma_launch_batch(const ConvParams *probs, int n, bool deform, int bn, cudaStream_t stream);

}  // namespace aanet
