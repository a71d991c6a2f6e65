// tcgen05 implicit-GEMM convolution engine (3xTF32) -- persistent, warp-specialised, channels-last.
//
// One launch computes  out[p, o] = epilogue( sum_{tap, c} A[p, (tap,c)] * W[o, c, tap] )  for all pixels,
// where A is never materialised in HBM:
//   * DEFORM: A[p,(tap,c)] = mask * bilinear(x[c], p*stride - pad + tap*dil + offset)   -- the ISA
//     operator, replacing modulated_deformable_im2col + cuBLAS SGEMM of the reference
//     (deform_conv_cuda_kernel.cu:570-633, deform_conv_cuda.cpp:539-561);
//   * DENSE : A[p,(tap,c)] = x[c, p*stride - pad + tap*dil] (zero padded)                -- the 1x1 / 3x3 /
//     strided / dilated / grouped convolutions around it (nets/deform.py:216-236, nets/aggregation.py:
//     346-371, :443-450), with folded-BN scale/shift, bias, residual and activation in the epilogue.
//
// Layout.  The engine reads x channels-last, [B][H*W][Cin]: the eight 16-byte chunks of one K row
// (32 consecutive channels of one tap) are one contiguous 128-byte line per source pixel, so a gather
// costs one L1 wavefront per (pixel, corner) whatever the learned offsets are, and a tile's input and
// output are contiguous in HBM.  (Measured on the first, NCHW / one-thread-per-pixel version: 81 M
// sectors per call for uncorrelated offsets, 1.7 TB/s on the dense 1x1, 15 k cycles to the first MMA.)
// The epilogue writes channels-last or NCHW; the NCHW <-> NHWC conversions needed at the reference-
// shaped boundary are done by transpose_kernel.
//
// Blackwell mapping.  A CTA is persistent (grid = #SMs) and walks 128-pixel tiles.  M = 128 pixels is
// the UMMA M (one TMEM lane per pixel), N = BN output channels, K walks (tap, channel) in blocks of
// 32 tf32 = one 128-byte swizzle row.  Warp roles:
//   warps 0-7   A producers: lane = (pixel row, 16-byte chunk); LDG.128 (4 per item for DEFORM, 1 for
//               DENSE), bilinear combine, split into tf32 hi + lo (cvt.rna + exact remainder), two
//               STS.128 into SWIZZLE_128B K-major tiles; fence.proxy.async + mbarrier arrive.
//   warp 13     streams the pre-split, pre-swizzled weight block of the stage with one cp.async.bulk.
//   warp 12     one thread issues tcgen05.mma.kind::tf32 three times per K step (lo*hi + hi*lo + hi*hi:
//               fp32-grade accuracy, the parity bar is 1e-4) into one of two TMEM accumulators and
//               releases the stage with tcgen05.commit.
//   warps 8-11  epilogue: tcgen05.ld (lane = pixel), bias / folded-BN affine / residual / activation,
//               128-bit channels-last stores (or coalesced NCHW stores); runs one tile behind the MMA.
// All hand-offs are mbarriers; the ring (4 stages at BN = 64) runs across tile boundaries.
#include "mdcn_common.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kUM = 128;                 // pixels per tile (UMMA M)
constexpr int kUK = 32;                  // K per stage (one 128-byte swizzle row of tf32)
constexpr int kProdWarps = 8;
constexpr int kMmaWarp = 12, kLoadWarp = 13;   // warps 8..11 are the epilogue (warp % 4 = TMEM lane quarter)
constexpr int kUThreads = 14 * 32;
constexpr int kATileBytes = kUM * kUK * 4;   // 16 KB (hi); same for lo
constexpr int kSmemBudget = 200 * 1024;

enum ConvAct { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_OFFSET_MASK = 3 };

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
    MdcnDims d;
    int K, KB;                          // K = kh*kw*Cg, KB = ceil(K / 32)
    int n_tiles_n;                      // ceil(Og / BN)
    int tiles_per_img;                  // ceil(P / 128)
    int n_ptiles;                       // B * tiles_per_img
    int total_tiles;                    // groups * n_tiles_n * n_ptiles
};

template <int BN> struct EngineCfg {
    static constexpr int kBTileBytes = BN * kUK * 4;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
    static constexpr int kStages = (kSmemBudget / kStageBytes) > 6 ? 6 : (kSmemBudget / kStageBytes);
    static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : 128;   // TMEM columns per accumulator
    static constexpr uint32_t kTmemCols = 2 * kAccStride;                     // two accumulators
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;
};

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

// [B][R][Cc] -> [B][Cc][R] through a 32x33 shared tile (both sides coalesced).  NCHW -> NHWC is
// (R = C, Cc = HW) and NHWC -> NCHW is (R = HW, Cc = C).
__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int R, long Cc) {
    __shared__ float tile[32][33];
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

struct TileCoord { int grp, nt, b; int p0; };

__device__ __forceinline__ TileCoord tile_coord(const ConvParams &p, int t) {
    TileCoord c;
    const int pt = t % p.n_ptiles, gn = t / p.n_ptiles;
    c.grp = gn / p.n_tiles_n; c.nt = gn % p.n_tiles_n;
    c.b = pt / p.tiles_per_img;
    c.p0 = (pt % p.tiles_per_img) * kUM;
    return c;
}

template <int BN, bool DEFORM>
__global__ void __launch_bounds__(kUThreads, 1)
conv_umma_kernel(const ConvParams p) {
    using Cfg = EngineCfg<BN>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full_a[S], bar_full_b[S], bar_empty[S];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;

    const MdcnDims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);   // 1024-byte aligned

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            umma::mbar_init(&bar_full_a[s], kProdWarps);   // one arrival per producer warp
            umma::mbar_init(&bar_full_b[s], 1);            // expect_tx arrival + bulk-copy bytes
            umma::mbar_init(&bar_empty[s], 1);             // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(&bar_acc_full[a], 1);          // tcgen05.commit after a tile's last MMA
            umma::mbar_init(&bar_acc_empty[a], 4);         // one arrival per epilogue warp
        }
        umma::fence_mbar_init();
    }
    if (warp == kMmaWarp) umma::tmem_alloc<Cfg::kTmemCols>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < kProdWarps) {
        // ================================ A producers ===========================================
        // item = (row, 16-byte chunk): chunk j = tid % 8 (fixed), rows r0 + 32*u, u = 0..3
        const int j = tid & 7, r0 = tid >> 3;
        const int P32 = (int)d.P;
        uint32_t it = 0;                                   // K-block counter across tiles
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(p, t);
            const float *x_b = p.x + (long)tc.b * d.HW * d.Cin;
            int oh[4], ow[4];
            bool rok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int px = tc.p0 + r0 + 32 * u;
                rok[u] = px < P32;
                const int pc = rok[u] ? px : P32 - 1;
                oh[u] = pc / d.Wo;
                ow[u] = pc - oh[u] * d.Wo;
            }
            const float *off_b = DEFORM ? p.offset + (long)tc.b * p.off_bs : nullptr;
            const float *mask_b = (DEFORM && p.mask) ? p.mask + (long)tc.b * p.mask_bs : nullptr;

            for (int kb = 0; kb < p.KB; ++kb, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                const int kk = kb * kUK + j * 4;
                const bool k_ok = kk < p.K;
                const int tap = k_ok ? kk / d.Cg : 0;
                const int c_abs = tc.grp * d.Cg + (k_ok ? kk - tap * d.Cg : 0);
                const int ki = tap / d.kw, kj = tap - ki * d.kw;
                float *a_hi = reinterpret_cast<float *>(smem + (size_t)s * Cfg::kStageBytes);
                float *a_lo = a_hi + kATileBytes / 4;
                float v[4][4];
                if (DEFORM) {
                    const int g = c_abs / d.Cd;
                    const long ch = (long)(g * d.K + tap);
                    float gh[4], gw[4], gm[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long pc = oh[u] * d.Wo + ow[u];
                        gh[u] = __ldg(off_b + pc * p.off_ps + (ch * 2) * p.off_cs);
                        gw[u] = __ldg(off_b + pc * p.off_ps + (ch * 2 + 1) * p.off_cs);
                        gm[u] = mask_b ? __ldg(mask_b + pc * p.mask_ps + ch * p.mask_cs) : 1.f;
                    }
                    float4 q[4][4];
                    float wgt[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float h = (float)(oh[u] * d.stride - d.pad + ki * d.dil) + gh[u];
                        const float w = (float)(ow[u] * d.stride - d.pad + kj * d.dil) + gw[u];
                        const Sample sm = make_sample(h, w, d.H, d.W);
                        const float m = (k_ok && rok[u]) ? gm[u] : 0.f;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            wgt[u][c4] = sm.w[c4] * m;
                            q[u][c4] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)sm.i[c4] * d.Cin + c_abs));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u][0] = wgt[u][0] * q[u][0].x + wgt[u][1] * q[u][1].x + wgt[u][2] * q[u][2].x + wgt[u][3] * q[u][3].x;
                        v[u][1] = wgt[u][0] * q[u][0].y + wgt[u][1] * q[u][1].y + wgt[u][2] * q[u][2].y + wgt[u][3] * q[u][3].y;
                        v[u][2] = wgt[u][0] * q[u][0].z + wgt[u][1] * q[u][1].z + wgt[u][2] * q[u][2].z + wgt[u][3] * q[u][3].z;
                        v[u][3] = wgt[u][0] * q[u][0].w + wgt[u][1] * q[u][1].w + wgt[u][2] * q[u][2].w + wgt[u][3] * q[u][3].w;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int hi_ = oh[u] * d.stride - d.pad + ki * d.dil;
                        const int wi_ = ow[u] * d.stride - d.pad + kj * d.dil;
                        const bool ok = k_ok && rok[u] && hi_ >= 0 && hi_ < d.H && wi_ >= 0 && wi_ < d.W;
                        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) q = __ldg(reinterpret_cast<const float4 *>(x_b + ((long)hi_ * d.W + wi_) * d.Cin + c_abs));
                        v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
                    }
                }
                umma::mbar_wait(&bar_empty[s], ph ^ 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int row = r0 + 32 * u;
                    float4 h4, l4;
                    umma::split_tf32(v[u][0], h4.x, l4.x); umma::split_tf32(v[u][1], h4.y, l4.y);
                    umma::split_tf32(v[u][2], h4.z, l4.z); umma::split_tf32(v[u][3], h4.w, l4.w);
                    const int at = row * kUK + ((j ^ (row & 7)) << 2);
                    *reinterpret_cast<float4 *>(a_hi + at) = h4;
                    *reinterpret_cast<float4 *>(a_lo + at) = l4;
                }
                umma::fence_proxy_async();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_full_a[s]);
            }
        }
    } else if (warp < kMmaWarp) {
        // ================================ epilogue: TMEM -> registers -> global ==================
        const int q = warp & 3;                                  // TMEM lane quarter
        const int row = q * 32 + lane;
        uint32_t ti = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
            const TileCoord tc = tile_coord(p, t);
            const int a = ti & 1;
            const int pix = tc.p0 + row;
            const bool p_ok = pix < (int)d.P;
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
            const int o_base = tc.grp * d.Og + tc.nt * BN;          // first global out channel of the tile
            const int n_valid = min(BN, d.Og - tc.nt * BN);
            const long pix_g = (long)tc.b * d.P + pix;
            const bool vec_ok = !p.out_nchw && ((d.Cout | o_base) & 3) == 0;
#pragma unroll 1
            for (int n0 = 0; n0 < BN; n0 += 16) {
                float acc[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + n0, acc);
                if (!p_ok || n0 >= n_valid) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int o = o_base + n0 + i;
                    if (n0 + i < n_valid) {
                        float tv = acc[i];
                        if (p.bias) tv += __ldg(p.bias + o);
                        if (p.scale) tv = fmaf(tv, __ldg(p.scale + o), __ldg(p.shift + o));
                        if (p.residual)
                            tv += p.out_nchw ? __ldg(p.residual + ((long)tc.b * d.Cout + o) * d.P + pix)
                                             : __ldg(p.residual + pix_g * d.Cout + o);
                        if (p.act == ACT_RELU) tv = fmaxf(tv, 0.f);
                        else if (p.act == ACT_LEAKY) tv = tv > 0.f ? tv : tv * p.slope;
                        else if (p.act == ACT_OFFSET_MASK && o >= p.n_offset_ch)
                            tv = p.mask_scale / (1.f + __expf(-tv));
                        acc[i] = tv;
                    }
                }
                if (p.out_nchw) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) p.out[((long)tc.b * d.Cout + o_base + n0 + i) * d.P + pix] = acc[i];
                } else {
                    float *dst = p.out + pix_g * d.Cout + o_base + n0;
                    if (vec_ok && n0 + 16 <= n_valid) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4 *>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + i < n_valid) dst[i] = acc[i];
                    }
                }
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);
        }
    } else if (warp == kLoadWarp) {
        // ================================ weight loader (bulk async copy) ========================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(p, t);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack) +
                                     (size_t)(tc.grp * p.n_tiles_n + tc.nt) * p.KB * (2 * Cfg::kBTileBytes);
                for (int kb = 0; kb < p.KB; ++kb, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    umma::mbar_wait_sleep(&bar_empty[s], ph ^ 1);
                    umma::mbar_expect_tx(&bar_full_b[s], 2 * Cfg::kBTileBytes);
                    umma::bulk_g2s(smem + (size_t)s * Cfg::kStageBytes + 2 * kATileBytes,
                                   src + (size_t)kb * 2 * Cfg::kBTileBytes, 2 * Cfg::kBTileBytes, &bar_full_b[s]);
                }
            }
        }
    } else {
        // ================================ MMA issuer (one thread) ================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma::make_idesc_tf32(kUM, BN);
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);   // epilogue drained this accumulator
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * Cfg::kAccStride;
                for (int kb = 0; kb < p.KB; ++kb, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    umma::mbar_wait_sleep(&bar_full_a[s], ph);
                    umma::mbar_wait_sleep(&bar_full_b[s], ph);
                    umma::tc_fence_after();
                    const uint32_t a0 = umma::smem_u32(smem + (size_t)s * Cfg::kStageBytes);
                    const uint64_t a_hi = umma::make_desc_sw128(a0), a_lo = umma::make_desc_sw128(a0 + kATileBytes);
                    const uint64_t b_hi = umma::make_desc_sw128(a0 + 2 * kATileBytes);
                    const uint64_t b_lo = umma::make_desc_sw128(a0 + 2 * kATileBytes + Cfg::kBTileBytes);
#pragma unroll
                    for (int k = 0; k < kUK / 8; ++k) {
                        const uint32_t adv = k * 32;     // 8 tf32 = 32 bytes along K inside the swizzle row
                        umma::mma_tf32(d_tmem, umma::desc_advance(a_lo, adv), umma::desc_advance(b_hi, adv), idesc,
                                       (kb | k) != 0);
                        umma::mma_tf32(d_tmem, umma::desc_advance(a_hi, adv), umma::desc_advance(b_lo, adv), idesc, 1);
                        umma::mma_tf32(d_tmem, umma::desc_advance(a_hi, adv), umma::desc_advance(b_hi, adv), idesc, 1);
                    }
                    umma::tc_commit(&bar_empty[s]);      // frees this stage when the MMAs above retire
                }
                umma::tc_commit(&bar_acc_full[a]);       // accumulator of this tile complete
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        umma::tc_fence_after();
        umma::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
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
    if (d.Cg % 4 || d.Cin % 4) return false;
    if (deform && (d.Cd % 4)) return false;
    if (d.P > 0x3fffffffLL || d.HW > 0x3fffffffLL) return false;
    if ((long)d.B * ceil_div_ll(d.P, kUM) > 0x3fffffffLL) return false;
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

// [B][R][Cc] -> [B][Cc][R]
int conv_umma_transpose(const float *src, float *dst, int B, int R, long Cc, cudaStream_t stream) {
    const dim3 grid((unsigned)ceil_div_ll(Cc, 32), ceil_div(R, 32), B);
    transpose_kernel<<<grid, 256, 0, stream>>>(src, dst, R, Cc);
    return check_launch();
}

template <int BN, bool DEFORM>
static int launch_one(const ConvParams &p, cudaStream_t stream) {
    constexpr size_t smem = EngineCfg<BN>::kSmemBytes;
    cudaFuncSetAttribute(conv_umma_kernel<BN, DEFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
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
    p.n_ptiles = d.B * p.tiles_per_img;
    const long total = (long)d.groups * p.n_tiles_n * p.n_ptiles;
    if (total > 0x7fffffffLL) return AANET_ERR_UNSUPPORTED;
    p.total_tiles = (int)total;
#define AANET_CONV_CASE(bn)                                                              \
    case bn:                                                                             \
        return deform ? launch_one<bn, true>(p, stream) : launch_one<bn, false>(p, stream);
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
