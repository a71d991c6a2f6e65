// Host side of the tcgen05 convolution engine: weight packing, layout kernels, problem set-up and dispatch to the
// per-MODE translation units (conv_umma_m0.cu .. m3.cu).  The kernel itself lives in conv_umma_kernel.cuh.
#include "conv_umma_kernel.cuh"

namespace aanet {

int conv_umma_launch_mode0(const ConvBatch &batch, int BN, cudaStream_t stream);
int conv_umma_launch_mode1(const ConvBatch &batch, int BN, cudaStream_t stream);
int conv_umma_launch_mode2(const ConvBatch &batch, int BN, cudaStream_t stream);
int conv_umma_launch_mode3(const ConvBatch &batch, int BN, cudaStream_t stream);

// weight [Cout, Cg, kh, kw] -> wpack[grp][nt][kb][hi|lo][BN x 32 swizzled], K index = tap*Cg + c.
__global__ void conv_pack_weights_kernel(const float *__restrict__ w, float *__restrict__ wpack, int Cout,
                                         int Cg, int T, int groups, int BN, int n_tiles_n, int K, int KB) {
    const int Og = Cout / groups;
    const long total = (long)groups * n_tiles_n * KB * BN * kUK;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int kq = (int)(i % kUK);
        long r = i / kUK;
        const int n = (int)(r % BN); r /= BN;
        const int kb = (int)(r % KB); r /= KB;
        const int nt = (int)(r % n_tiles_n);
        const int grp = (int)(r / n_tiles_n);
        const int kk = kb * kUK + kq, ol = nt * BN + n;
        float v = 0.f;
        if (kk < K && ol < Og) {
            const int tap = kk / Cg, c = kk % Cg;
            v = w[((long)(grp * Og + ol) * Cg + c) * T + tap];
        }
        float hi, lo;
        umma::split_tf32(v, hi, lo);
        float *blk = wpack + ((long)(grp * n_tiles_n + nt) * KB + kb) * (2 * BN * kUK);
        blk[sw128_index(n, kq)] = hi;
        blk[BN * kUK + sw128_index(n, kq)] = lo;
    }
}

// [B][R][Cc] -> [B][Cc][R] through a 32x33 shared tile (both sides coalesced).  NCHW -> NHWC is
// (R = C, Cc = HW) and NHWC -> NCHW is (R = HW, Cc = C).
__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int R, long Cc) {
    __shared__ float tile[32][33];
    pdl_wait();
    pdl_trigger();
    const long c0 = (long)blockIdx.x * 32;
    const int r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const float *sb = src + (long)blockIdx.z * R * Cc;
    float *db = dst + (long)blockIdx.z * R * Cc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8;
        const long c = c0 + tx;
        tile[ty + i * 8][tx] = (r < R && c < Cc) ? sb[(long)r * Cc + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long c = c0 + ty + i * 8;
        const int r = r0 + tx;
        if (r < R && c < Cc) db[c * R + r] = tile[tx][ty + i * 8];
    }
}

// ------------------------------------------------------------------------------------------ host side
// N tile: a multiple of 16, at most 64; wider outputs are split evenly (96 -> 2 x 48, 128 -> 2 x 64).
int conv_umma_pick_bn(int Og) {
    const int n_tiles = (Og + 63) / 64;
    const int per = (Og + n_tiles - 1) / n_tiles;
    return (per + 15) / 16 * 16;
}

static bool bn_valid(int bn) { return bn == 16 || bn == 32 || bn == 48 || bn == 64; }

bool conv_umma_supported(const MdcnDims &d, bool deform) {
    if (d.Cg % 4 || d.Cin % 4) return false;
    if (deform && (d.Cd % 4)) return false;
    if (d.P > 0x3fffffffLL || d.HW > 0x3fffffffLL) return false;
    if ((long)d.B * ceil_div_ll(d.P, kUM) > 0x3fffffffLL) return false;
    if (ceil_div(d.K * d.Cg, kUK) > kMaxKB || d.Cg > 0xffff || d.kh > 15 || d.kw > 15) return false;
    if (d.Ho > 0x7fff || d.Wo > 0xffff) return false;      // producers pack (oh, ow) into one register
    return true;
}

// bn == 0: the natural N tile of this layer; otherwise the (wider) N tile of the batch it will run in.
size_t conv_umma_wpack_bytes(const MdcnDims &d, int bn) {
    const int BN = bn ? bn : conv_umma_pick_bn(d.Og);
    const int n_tiles_n = ceil_div(d.Og, BN);
    const int KB = ceil_div(d.K * d.Cg, kUK);
    return (size_t)d.groups * n_tiles_n * KB * 2 * BN * kUK * sizeof(float);
}

int conv_umma_pack(const float *weight, void *wpack, const MdcnDims &d, int bn, cudaStream_t stream) {
    const int BN = bn ? bn : conv_umma_pick_bn(d.Og);
    if (!bn_valid(BN)) return AANET_ERR_UNSUPPORTED;
    const int n_tiles_n = ceil_div(d.Og, BN);
    const int K = d.K * d.Cg, KB = ceil_div(K, kUK);
    const long total = (long)d.groups * n_tiles_n * KB * BN * kUK;
    const int grid = (int)(ceil_div_ll(total, 256) < 1184 ? ceil_div_ll(total, 256) : 1184);
    conv_pack_weights_kernel<<<grid, 256, 0, stream>>>(weight, static_cast<float *>(wpack), d.Cout, d.Cg, d.K,
                                                       d.groups, BN, n_tiles_n, K, KB);
    return check_launch();
}

// [B][R][Cc] -> [B][Cc][R]
int conv_umma_transpose(const float *src, float *dst, int B, int R, long Cc, cudaStream_t stream) {
    const dim3 grid((unsigned)ceil_div_ll(Cc, 32), ceil_div(R, 32), B);
    return launch_pdl(transpose_kernel, grid, dim3(256), 0, stream, src, dst, R, Cc);
}

// Launch up to kMaxProblems problems of the same kind as one persistent kernel.  bn == 0: widest natural N
// tile among the problems (their weights must have been packed with that same width).
int conv_umma_launch_batch(const ConvParams *probs, int n, bool deform, int bn, cudaStream_t stream) {
    if (n < 1 || n > kMaxProblems) return AANET_ERR_SHAPE;
    int BN = bn;
    if (!BN)
        for (int i = 0; i < n; ++i) BN = BN > conv_umma_pick_bn(probs[i].d.Og) ? BN : conv_umma_pick_bn(probs[i].d.Og);
    if (!bn_valid(BN)) return AANET_ERR_UNSUPPORTED;
    if (n == 1 && !deform && probs[0].act != ACT_SOFTARGMIN) {   // dense layers with 32-channel blocks: operands from a TMA halo
        int rc = dense_tmem_launch(probs[0], BN, stream);
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
        rc = conv_halo_launch(probs[0], BN, stream);
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
    }
    if (n == 1 && deform) {         // ISA layers with 32-channel deformable groups: gather from a TMA-staged halo
        int rc = deform_tmem_launch(probs[0], BN, stream);
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
        rc = deform_halo_launch(probs[0], BN, stream);
        if (rc != AANET_ERR_UNSUPPORTED) return rc;
    }
    ConvBatch batch{};
    batch.n = n;
    long tiles = 0;
    int tbl = 0;
    for (int i = 0; i < n; ++i) {
        ConvParams p = probs[i];
        const MdcnDims &d = p.d;
        p.n_tiles_n = ceil_div(d.Og, BN);
        p.K = d.K * d.Cg;
        p.KB = ceil_div(p.K, kUK);
        // 1-D tiles (128 consecutive pixels) only where input pixel == output pixel
        p.tile2d = (d.K > 1 || d.stride != 1 || d.pad != 0) ? 1 : 0;
        p.tiles_x = ceil_div(d.Wo, kTileW);
        p.tiles_per_img = p.tile2d ? p.tiles_x * ceil_div(d.Ho, kTileH) : (int)ceil_div_ll(d.P, kUM);
        p.n_ptiles = d.B * p.tiles_per_img;
        const long total = (long)d.groups * p.n_tiles_n * p.n_ptiles;
        if (tiles + total > 0x3fffffffLL) return AANET_ERR_UNSUPPORTED;
        p.total_tiles = (int)total;
        p.tile_start = (int)tiles;
        p.tbl_off = tbl;
        tiles += total;
        tbl += p.KB;
        batch.pr[i] = p;
    }
    if (tbl > kMaxKB) return AANET_ERR_UNSUPPORTED;
    batch.total_tiles = (int)tiles;
    bool pointwise = !deform;                  // only 1-D tiles in the launch
    for (int i = 0; i < n; ++i) pointwise &= batch.pr[i].tile2d == 0;
    bool single_run = deform;                  // every K block = 32 channels of one tap inside one deformable group
    for (int i = 0; i < n; ++i) single_run &= batch.pr[i].d.Cg % kUK == 0 && batch.pr[i].d.Cd % kUK == 0;
    if (!deform) return pointwise ? conv_umma_launch_mode3(batch, BN, stream) : conv_umma_launch_mode0(batch, BN, stream);
    return single_run ? conv_umma_launch_mode2(batch, BN, stream) : conv_umma_launch_mode1(batch, BN, stream);
}

int conv_umma_launch(ConvParams p, bool deform, cudaStream_t stream) {
    return conv_umma_launch_batch(&p, 1, deform, 0, stream);
}

}  // namespace aanet
