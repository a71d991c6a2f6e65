// tcgen05 implicit-GEMM convolution engine (3xTF32), dense or deformable A-operand producer.
//
// One CTA computes a [128 output pixels] x [BN output channels] tile of
//      out[p, o] = sum_{tap, c} A[p, (tap,c)] * W[o, c, tap]
// where A is never materialised in HBM:
//   * DEFORM: A[p,(tap,c)] = mask * bilinear(x[c], p*stride - pad + tap*dil + offset)   -- the ISA
//     operator, replacing modulated_deformable_im2col + cuBLAS SGEMM of the reference
//     (deform_conv_cuda_kernel.cu:570-633, deform_conv_cuda.cpp:539-561);
//   * DENSE : A[p,(tap,c)] = x[c, p*stride - pad + tap*dil] (zero padded)                -- the 1x1 / 3x3 /
//     strided / dilated / grouped convolutions around it (nets/deform.py:216-236, nets/aggregation.py:
//     346-371, :443-450), with folded-BN scale/shift, bias, residual and activation in the epilogue.
//
// Blackwell mapping.  M = 128 pixels is the UMMA M (one TMEM lane per pixel), N = BN channels, K walks
// (tap, channel) in blocks of 32 tf32 = one 128-byte swizzle row.  Warps 0-3 (one thread per pixel)
// produce the A tile: they gather/interpolate, split every value into tf32 hi + lo (cvt.rna, exact
// remainder) and store both as SWIZZLE_128B K-major tiles in shared memory; warp 5 streams the matching
// pre-split, pre-swizzled weight block with one cp.async.bulk per stage; one thread of warp 4 issues
// tcgen05.mma.kind::tf32 three times per K step (lo*hi + hi*lo + hi*hi -> fp32-grade accuracy, the
// parity bar is 1e-4) into a TMEM accumulator and frees the stage with tcgen05.commit.  When the last
// commit lands, warps 0-3 read the accumulator back with tcgen05.ld (lane = pixel -> coalesced NCHW
// stores) and apply the epilogue.  Stages are handed over with mbarriers only; there is no
// __syncthreads in the main loop.  Two CTAs fit per SM (96 KB each at BN<=64) so one tile's epilogue
// overlaps the other's main loop.
#include "mdcn_common.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kUM = 128;            // pixels per tile (UMMA M)
constexpr int kUK = 32;             // K per stage (one 128-byte swizzle row of tf32)
constexpr int kUStages = 2;
constexpr int kUThreads = 192;      // 4 producer/epilogue warps + MMA warp + weight-loader warp
constexpr int kATileBytes = kUM * kUK * 4;   // 16 KB (hi) ; same for lo

enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_OFFSET_MASK = 3 };

struct ConvParams {
    const float *x, *offset, *mask;     // offset/mask only for DEFORM (mask may be NULL: DCNv1)
    const float *wpack;                 // packed weights, see conv_pack_weights_kernel
    float *out;
    const float *bias, *scale, *shift, *residual;
    int act; float slope; int n_offset_ch;       // ACT_OFFSET_MASK: channels >= n_offset_ch get 2*sigmoid
    float mask_scale;
    MdcnDims d;
    int K, KB;                          // K = kh*kw*Cg, KB = ceil(K / 32)
    int n_tiles_n;                      // ceil(Og / BN)
    int tiles_per_img;                  // ceil(P / 128)
};

template <int BN> struct TmemCols { static constexpr uint32_t value = BN <= 32 ? 32 : BN <= 64 ? 64 : 128; };

template <int BN>
constexpr size_t conv_umma_smem_bytes() {
    return (size_t)kUStages * (2 * kATileBytes + 2 * BN * kUK * 4) + 1024;   // + alignment slack
}

// Swizzled position (in floats) of element (row, k) inside a [rows x 32] SWIZZLE_128B K-major tile.
__host__ __device__ inline int sw128_index(int row, int k) {
    return row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

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

template <int BN, bool DEFORM>
__global__ void __launch_bounds__(kUThreads, BN <= 64 ? 2 : 1)
conv_umma_kernel(const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full_a[kUStages], bar_full_b[kUStages], bar_empty[kUStages], bar_accum;
    __shared__ uint32_t s_tmem;

    const MdcnDims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kBTileBytes = BN * kUK * 4;
    constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);   // 1024-byte aligned

    const int b = blockIdx.x / p.tiles_per_img;
    const long p0 = (long)(blockIdx.x % p.tiles_per_img) * kUM;
    const int grp = blockIdx.y / p.n_tiles_n, nt = blockIdx.y % p.n_tiles_n;

    if (tid == 0) {
        for (int s = 0; s < kUStages; ++s) {
            umma::mbar_init(&bar_full_a[s], 4);     // one arrival per producer warp
            umma::mbar_init(&bar_full_b[s], 1);     // expect_tx arrival + bulk-copy bytes
            umma::mbar_init(&bar_empty[s], 1);      // tcgen05.commit
        }
        umma::mbar_init(&bar_accum, 1);
        umma::fence_mbar_init();
    }
    if (warp == 4) umma::tmem_alloc<TmemCols<BN>::value>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < 4) {
        // ================================ A producer: one thread per output pixel ================
        const long pix = p0 + tid;
        const bool p_ok = pix < d.P;
        const long pc = p_ok ? pix : 0;
        const int ho = (int)(pc / d.Wo), wo = (int)(pc % d.Wo);
        const float *x_b = p.x + (long)b * d.Cin * d.HW;
        const float *off_b = DEFORM ? p.offset + (long)b * d.dg * 2 * d.K * d.P : nullptr;
        const float *mask_b = (DEFORM && p.mask) ? p.mask + (long)b * d.dg * d.K * d.P : nullptr;
        const int row_sw = tid & 7;
        int cur_tap = -1, cur_g = -1;
        int si[4] = {0, 0, 0, 0};
        float sw[4] = {0.f, 0.f, 0.f, 0.f};
        bool dense_ok = false; int dense_idx = 0;

        for (int kb = 0; kb < p.KB; ++kb) {
            const int s = kb % kUStages;
            const uint32_t ph = (kb / kUStages) & 1;
            umma::mbar_wait(&bar_empty[s], ph ^ 1);
            float *a_hi = reinterpret_cast<float *>(smem + (size_t)s * kStageBytes) + tid * kUK;
            float *a_lo = a_hi + kATileBytes / 4;
#pragma unroll 2
            for (int j = 0; j < 8; ++j) {
                const int kk = kb * kUK + j * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (kk < p.K && p_ok) {
                    const int tap = kk / d.Cg, c = kk - tap * d.Cg;
                    const int c_abs = grp * d.Cg + c;
                    if (DEFORM) {
                        const int g = c_abs / d.Cd;
                        if (tap != cur_tap || g != cur_g) {
                            cur_tap = tap; cur_g = g;
                            const Sample sm = sample_at(d, off_b, g, tap, ho, wo, pc);
                            const float m = mask_b ? mask_b[(long)(g * d.K + tap) * d.P + pc] : 1.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) { si[q] = sm.i[q]; sw[q] = sm.w[q] * m; }
                        }
                        const float *im = x_b + (long)c_abs * d.HW;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float *ime = im + (long)e * d.HW;
                            v[e] = sw[0] * __ldg(ime + si[0]) + sw[1] * __ldg(ime + si[1]) +
                                   sw[2] * __ldg(ime + si[2]) + sw[3] * __ldg(ime + si[3]);
                        }
                    } else {
                        if (tap != cur_tap) {
                            cur_tap = tap;
                            const int hi_ = ho * d.stride - d.pad + (tap / d.kw) * d.dil;
                            const int wi_ = wo * d.stride - d.pad + (tap % d.kw) * d.dil;
                            dense_ok = hi_ >= 0 && hi_ < d.H && wi_ >= 0 && wi_ < d.W;
                            dense_idx = dense_ok ? hi_ * d.W + wi_ : 0;
                        }
                        if (dense_ok) {
                            const float *im = x_b + (long)c_abs * d.HW + dense_idx;
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = __ldg(im + (long)e * d.HW);
                        }
                    }
                }
                float4 h4, l4;
                umma::split_tf32(v[0], h4.x, l4.x); umma::split_tf32(v[1], h4.y, l4.y);
                umma::split_tf32(v[2], h4.z, l4.z); umma::split_tf32(v[3], h4.w, l4.w);
                const int chunk = (j ^ row_sw) << 2;
                *reinterpret_cast<float4 *>(a_hi + chunk) = h4;
                *reinterpret_cast<float4 *>(a_lo + chunk) = l4;
            }
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_full_a[s]);
        }

        // ================================ epilogue: TMEM -> registers -> NCHW =====================
        umma::mbar_wait(&bar_accum, 0);
        umma::tc_fence_after();
        const long out_b = (long)b * d.Cout * d.P;
#pragma unroll 1
        for (int n0 = 0; n0 < BN; n0 += 16) {
            float acc[16];
            umma::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + n0, acc);
            if (!p_ok) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int ol = nt * BN + n0 + i;
                if (ol >= d.Og) break;
                const int o = grp * d.Og + ol;
                float t = acc[i];
                if (p.bias) t += __ldg(p.bias + o);
                if (p.scale) t = fmaf(t, __ldg(p.scale + o), __ldg(p.shift + o));
                const long oi = out_b + (long)o * d.P + pix;
                if (p.residual) t += __ldg(p.residual + oi);
                if (p.act == ACT_RELU) t = fmaxf(t, 0.f);
                else if (p.act == ACT_LEAKY) t = t > 0.f ? t : t * p.slope;
                else if (p.act == ACT_OFFSET_MASK && o >= p.n_offset_ch) t = p.mask_scale / (1.f + __expf(-t));
                p.out[oi] = t;
            }
        }
        umma::tc_fence_before();
    } else if (warp == 5) {
        // ================================ weight loader (bulk async copy) ========================
        if (lane == 0) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack) +
                                 (size_t)(grp * p.n_tiles_n + nt) * p.KB * (2 * kBTileBytes);
            for (int kb = 0; kb < p.KB; ++kb) {
                const int s = kb % kUStages;
                const uint32_t ph = (kb / kUStages) & 1;
                umma::mbar_wait(&bar_empty[s], ph ^ 1);
                umma::mbar_expect_tx(&bar_full_b[s], 2 * kBTileBytes);
                umma::bulk_g2s(smem + (size_t)s * kStageBytes + 2 * kATileBytes, src + (size_t)kb * 2 * kBTileBytes,
                               2 * kBTileBytes, &bar_full_b[s]);
            }
        }
    } else {
        // ================================ MMA issuer (one thread) ================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma::make_idesc_tf32(kUM, BN);
            for (int kb = 0; kb < p.KB; ++kb) {
                const int s = kb % kUStages;
                const uint32_t ph = (kb / kUStages) & 1;
                umma::mbar_wait(&bar_full_a[s], ph);
                umma::mbar_wait(&bar_full_b[s], ph);
                umma::tc_fence_after();
                const uint32_t a0 = umma::smem_u32(smem + (size_t)s * kStageBytes);
                const uint64_t a_hi = umma::make_desc_sw128(a0), a_lo = umma::make_desc_sw128(a0 + kATileBytes);
                const uint64_t b_hi = umma::make_desc_sw128(a0 + 2 * kATileBytes);
                const uint64_t b_lo = umma::make_desc_sw128(a0 + 2 * kATileBytes + kBTileBytes);
#pragma unroll
                for (int k = 0; k < kUK / 8; ++k) {
                    const uint32_t adv = k * 32;     // 8 tf32 = 32 bytes along K inside the swizzle row
                    umma::mma_tf32(tmem_base, umma::desc_advance(a_lo, adv), umma::desc_advance(b_hi, adv), idesc,
                                   (kb | k) != 0);
                    umma::mma_tf32(tmem_base, umma::desc_advance(a_hi, adv), umma::desc_advance(b_lo, adv), idesc, 1);
                    umma::mma_tf32(tmem_base, umma::desc_advance(a_hi, adv), umma::desc_advance(b_hi, adv), idesc, 1);
                }
                umma::tc_commit(&bar_empty[s]);      // frees this stage when the MMAs above retire
            }
            umma::tc_commit(&bar_accum);             // accumulator complete
        }
    }
    __syncthreads();
    if (warp == 4) {
        umma::tc_fence_after();
        umma::tmem_dealloc<TmemCols<BN>::value>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ host side
int conv_umma_pick_bn(int Og) {
    const int r = (Og + 15) / 16 * 16;
    if (r <= 16) return 16;
    if (r <= 32) return 32;
    if (r <= 48) return 48;
    if (r <= 64) return 64;
    if (r <= 96) return 96;
    return 128;
}

bool conv_umma_supported(const MdcnDims &d, bool deform) {
    if (d.Cg % 4) return false;
    if (deform && (d.Cd % 4)) return false;
    if (d.P > 0x7fffffffLL || d.HW > 0x7fffffffLL) return false;
    return true;
}

size_t conv_umma_wpack_bytes(const MdcnDims &d) {
    const int BN = conv_umma_pick_bn(d.Og);
    const int n_tiles_n = ceil_div(d.Og, BN);
    const int KB = ceil_div(d.K * d.Cg, kUK);
    return (size_t)d.groups * n_tiles_n * KB * 2 * BN * kUK * sizeof(float);
}

int conv_umma_pack(const float *weight, void *wpack, const MdcnDims &d, cudaStream_t stream) {
    const int BN = conv_umma_pick_bn(d.Og);
    const int n_tiles_n = ceil_div(d.Og, BN);
    const int K = d.K * d.Cg, KB = ceil_div(K, kUK);
    const long total = (long)d.groups * n_tiles_n * KB * BN * kUK;
    const int grid = (int)(ceil_div_ll(total, 256) < 1184 ? ceil_div_ll(total, 256) : 1184);
    conv_pack_weights_kernel<<<grid, 256, 0, stream>>>(weight, static_cast<float *>(wpack), d.Cout, d.Cg, d.K,
                                                       d.groups, BN, n_tiles_n, K, KB);
    return check_launch();
}

template <int BN, bool DEFORM>
static int launch_one(const ConvParams &p, dim3 grid, cudaStream_t stream) {
    constexpr size_t smem = conv_umma_smem_bytes<BN>();
    cudaFuncSetAttribute(conv_umma_kernel<BN, DEFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_umma_kernel<BN, DEFORM><<<grid, kUThreads, smem, stream>>>(p);
    return check_launch();
}

int conv_umma_launch(ConvParams p, bool deform, cudaStream_t stream) {
    const MdcnDims &d = p.d;
    const int BN = conv_umma_pick_bn(d.Og);
    p.n_tiles_n = ceil_div(d.Og, BN);
    p.K = d.K * d.Cg;
    p.KB = ceil_div(p.K, kUK);
    p.tiles_per_img = (int)ceil_div_ll(d.P, kUM);
    const long gx = (long)d.B * p.tiles_per_img;
    const int gy = d.groups * p.n_tiles_n;
    if (gx > 0x7fffffffLL || gy > 65535) return AANET_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)gx, gy);
#define AANET_CONV_CASE(bn)                                                              \
    case bn:                                                                             \
        return deform ? launch_one<bn, true>(p, grid, stream) : launch_one<bn, false>(p, grid, stream);
    switch (BN) {
        AANET_CONV_CASE(16)
        AANET_CONV_CASE(32)
        AANET_CONV_CASE(48)
        AANET_CONV_CASE(64)
        AANET_CONV_CASE(96)
        AANET_CONV_CASE(128)
    }
#undef AANET_CONV_CASE
    return AANET_ERR_UNSUPPORTED;
}

}  // namespace aanet
