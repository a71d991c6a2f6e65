// Modulated deformable convolution backward, input / offset / mask gradients, tensor-core assisted.
//
// Same gradients as mdcn_bwd_input_kernel (reference deform_conv_cuda.cpp:623-646, col2im_coord cu:695-767, col2im
// cu:635-693), reorganised for B200:
//
//   1. gout and x go channels-last (transpose_kernel).
//   2. The column gradient  gcol[p, (tap, c)] = sum_o gout[p, o] * W[o, c, tap]  is a 1x1 convolution of gout with
//      the permuted weight -- one launch of the tcgen05 engine (3xTF32), written channels-last so that the 32..64
//      channels of one (pixel, tap) are contiguous.
//   3. mdcn_bwd_scatter_kernel: 8 lanes per (pixel, tap, deformable group) item; a lane handles 16-byte channel
//      chunks: one LDG.128 of gcol, four LDG.128 of the x corners, the mask / offset partial sums, and the input
//      gradient as four red.global.add.v4.f32 (a quarter of the atomic operations of the scalar kernel, which is
//      atomic-throughput bound).  The three partial sums are reduced over the 8 lanes with shuffles.
//   4. grad_input goes back to NCHW.
//
// Used when groups == 1 and the channel counts are multiples of 4 (and of 4 per deformable group); other shapes keep
// the FFMA kernel in mdcn_bwd.cu.  Workspace: see mdcn_bwd_umma_workspace_bytes.
#include "conv_engine.cuh"

namespace aanet {

// [Cout][Cin][K] -> [K*Cin][Cout]   (row n = tap*Cin + c: the 1x1 weight of the column-gradient GEMM)
__global__ void mdcn_wt_permute_kernel(const float *__restrict__ w, float *__restrict__ wt, int Cout, int Cin, int K) {
    const long n = (long)Cout * Cin * K;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int o = (int)(i % Cout);
        const long r = i / Cout;                 // r = tap*Cin + c
        const int c = (int)(r % Cin), k = (int)(r / Cin);
        wt[i] = w[((long)o * Cin + c) * K + k];
    }
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256)
mdcn_bwd_scatter_kernel(const float *__restrict__ x_nhwc, const float *__restrict__ offset,
                        const float *__restrict__ mask, const float *__restrict__ gcol, float *__restrict__ gx_nhwc,
                        float *__restrict__ goffset, float *__restrict__ gmask, float *__restrict__ col, MdcnDims d,
                        long n_items) {
    const int lane8 = threadIdx.x & 7;
    const long item0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long stride_items = ((long)gridDim.x * blockDim.x) >> 3;
    const int cv = d.Cd >> 2;                               // 16-byte chunks per deformable group
    // The loop trip count is uniform per 8-lane group only (n_items need not be a multiple of 4): the shuffles
    // below name just the 8 lanes of this group.
    const unsigned group_mask = 0xffu << (threadIdx.x & 24);
    for (long it = item0; it < n_items; it += stride_items) {   // uniform per 8-lane group
        const long p = it % d.P;
        long r = it / d.P;
        const int g = (int)(r % d.dg); r /= d.dg;
        const int k = (int)(r % d.K);
        const long b = r / d.K;
        const int ho = (int)(p / d.Wo), wo = (int)(p % d.Wo);
        const float *off_b = offset + b * d.dg * 2 * d.K * d.P;
        const Sample s = sample_at(d, off_b, g, k, ho, wo, p);
        const float m = mask ? mask[(b * d.dg * d.K + g * d.K + k) * d.P + p] : 1.f;
        const float hh = 1.f - s.lh, hw = 1.f - s.lw;
        float a_m = 0.f, a_oh = 0.f, a_ow = 0.f;
        // modulated column of this item (the forward's A operand), kept for the weight gradient: the corners are in
        // registers here anyway, and the weight kernel then needs no gather of its own
        float *col_p = col + ((b * d.P + p) * d.K + k) * d.Cin + g * d.Cd;
        if (!s.valid) {
            for (int ch = lane8; ch < cv; ch += 8)
                reinterpret_cast<float4 *>(col_p)[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (s.valid) {
            const float *gc_p = gcol + ((b * d.P + p) * d.K + k) * d.Cin + g * d.Cd;
            const float *x_b = x_nhwc + b * d.HW * d.Cin + g * d.Cd;
            float *gx_b = gx_nhwc + b * d.HW * d.Cin + g * d.Cd;
            for (int ch = lane8; ch < cv; ch += 8) {
                const float4 gc = __ldg(reinterpret_cast<const float4 *>(gc_p) + ch);
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = ((s.ok >> i) & 1) ? __ldg(reinterpret_cast<const float4 *>(x_b + (long)s.i[i] * d.Cin) + ch)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                const float gcs[4] = {gc.x, gc.y, gc.z, gc.w};
                const float vv[4][4] = {{v[0].x, v[0].y, v[0].z, v[0].w}, {v[1].x, v[1].y, v[1].z, v[1].w},
                                        {v[2].x, v[2].y, v[2].z, v[2].w}, {v[3].x, v[3].y, v[3].z, v[3].w}};
                float t[4], cl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v1 = vv[0][e], v2 = vv[1][e], v3 = vv[2][e], v4 = vv[3][e];
                    const float val = s.w[0] * v1 + s.w[1] * v2 + s.w[2] * v3 + s.w[3] * v4;
                    cl[e] = val * m;
                    a_m = fmaf(gcs[e], val, a_m);
                    t[e] = gcs[e] * m;
                    a_oh = fmaf(t[e], -hw * v1 - s.lw * v2 + hw * v3 + s.lw * v4, a_oh);
                    a_ow = fmaf(t[e], -hh * v1 + hh * v2 - s.lh * v3 + s.lh * v4, a_ow);
                }
                reinterpret_cast<float4 *>(col_p)[ch] = make_float4(cl[0], cl[1], cl[2], cl[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((s.ok >> i) & 1)
                        red_add_v4(gx_b + (long)s.i[i] * d.Cin + ch * 4, t[0] * s.w[i], t[1] * s.w[i], t[2] * s.w[i],
                                   t[3] * s.w[i]);
            }
        }
#pragma unroll
        for (int sh = 4; sh > 0; sh >>= 1) {
            a_m += __shfl_xor_sync(group_mask, a_m, sh);
            a_oh += __shfl_xor_sync(group_mask, a_oh, sh);
            a_ow += __shfl_xor_sync(group_mask, a_ow, sh);
        }
        if (lane8 == 0) {
            goffset[(b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 0) * d.P + p] = a_oh;
            goffset[(b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 1) * d.P + p] = a_ow;
            if (gmask) gmask[(b * d.dg * d.K + g * d.K + k) * d.P + p] = a_m;
        }
    }
}

static size_t align256_(size_t n) { return (n + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------ grad_weight
// grad_weight[o, c, k] = sum_{b,p} gout[b,p,o] * col[b,p,(k,c)],  col = mask * bilinear(x) -- a GEMM whose reduction
// runs over the pixels, which the convolution engine (M = pixels) cannot express.  Warp-level tensor cores instead:
// a CTA owns (tap k, 32 input channels inside one deformable group, 64 output channels) and a strided share of the
// 64-pixel tiles.  Per tile the modulated columns (written by the scatter kernel, which has the corners in registers
// anyway) and the gout tile are copied into shared memory with LDG.128, and four warps accumulate D[64 o x 32 c] with
// mma.sync.m16n8k8 tf32, each product issued three times (hi*hi + hi*lo + lo*hi, operands split by mantissa mask).
// Per-split partials are summed in a fixed order by mdcn_bwd_weight_reduce_kernel (deterministic grad_weight).
constexpr int kMP = 64, kMC = 32, kMO = 64, kMThreads = 128;
constexpr int kMColStride = kMC + 8, kMGStride = kMO + 8;      // = 8 mod 32: conflict-free fragment loads

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_bits(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

__global__ void __launch_bounds__(kMThreads)
mdcn_bwd_weight_mma_kernel(const float *__restrict__ col, const float *__restrict__ gout_nhwc,
                           float *__restrict__ partial, MdcnDims d, int n_cchunks, int n_otiles, int tiles_per_img,
                           int splits) {
    __shared__ __align__(16) float s_col[kMP][kMColStride];
    __shared__ __align__(16) float s_g[kMP][kMGStride];
    int item = blockIdx.x;
    const int ot = item % n_otiles; item /= n_otiles;
    const int ci = item % n_cchunks;
    const int k = item / n_cchunks;
    const int split = blockIdx.y;
    // channel chunks of 32 never straddle a deformable group: chunks per group = ceil(Cd / 32)
    const int cpg = (d.Cd + kMC - 1) / kMC;
    const int g = ci / cpg, c0 = g * d.Cd + (ci % cpg) * kMC;
    const int nc = min(kMC, (g + 1) * d.Cd - c0);
    const int o0 = ot * kMO;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fg = lane >> 2, ft = lane & 3;              // fragment row group / thread-in-group
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int chunk = tid & 7;                            // gather role: 16-byte channel chunk, pixel (tid >> 3) + 16 i
    const long T = (long)d.B * tiles_per_img;
    for (long t = split; t < T; t += splits) {
        const long b = t / tiles_per_img;
        const long p0 = (t % tiles_per_img) * kMP;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                 // column tile: 64 px x 8 chunks of 16 bytes
            const int px = (tid >> 3) + 16 * i;
            const long p = p0 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < d.P && chunk * 4 < nc)
                v = __ldg(reinterpret_cast<const float4 *>(col + ((b * d.P + p) * d.K + k) * d.Cin + c0) + chunk);
            *reinterpret_cast<float4 *>(&s_col[px][chunk * 4]) = v;
        }
        for (int i = tid; i < kMP * (kMO / 4); i += kMThreads) {     // gout tile: 64 px x 16 chunks
            const int px = i / (kMO / 4), oc = (i % (kMO / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p0 + px < d.P && o0 + oc < d.Cout)
                v = __ldg(reinterpret_cast<const float4 *>(gout_nhwc + (b * d.P + p0 + px) * d.Cout + o0 + oc));
            *reinterpret_cast<float4 *>(&s_g[px][oc]) = v;
        }
        __syncthreads();
        const int ow = warp * 16;                                     // this warp's 16 output channels
#pragma unroll 2
        for (int kk = 0; kk < kMP; kk += 8) {
            uint32_t ah[4], al[4];
            split_bits(s_g[kk + ft][ow + fg], ah[0], al[0]);
            split_bits(s_g[kk + ft][ow + fg + 8], ah[1], al[1]);
            split_bits(s_g[kk + ft + 4][ow + fg], ah[2], al[2]);
            split_bits(s_g[kk + ft + 4][ow + fg + 8], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t bh[2], bl[2];
                split_bits(s_col[kk + ft][nt * 8 + fg], bh[0], bl[0]);
                split_bits(s_col[kk + ft + 4][nt * 8 + fg], bh[1], bl[1]);
                mma_tf32_m16n8k8(acc[nt], ah, bh);
                mma_tf32_m16n8k8(acc[nt], ah, bl);
                mma_tf32_m16n8k8(acc[nt], al, bh);
            }
        }
        __syncthreads();
    }
    // partial[split][o][c][k]; fragment: acc[nt][0..3] = D[fg][2 ft], D[fg][2 ft + 1], D[fg + 8][2 ft], D[fg + 8][2 ft + 1]
    float *dst = partial + (size_t)split * d.Cout * d.Cin * d.K;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = o0 + warp * 16 + fg + (e >> 1) * 8;
            const int cc = nt * 8 + 2 * ft + (e & 1);
            if (cc < nc && o < d.Cout) dst[((size_t)o * d.Cin + c0 + cc) * d.K + k] = acc[nt][e];
        }
}

struct WeightMmaPlan { int n_cchunks, n_otiles, n_items, tiles_per_img, splits; };

static WeightMmaPlan plan_weight_mma(const MdcnDims &d) {
    WeightMmaPlan w;
    w.n_cchunks = d.dg * ((d.Cd + kMC - 1) / kMC);
    w.n_otiles = ceil_div(d.Cout, kMO);
    w.n_items = d.K * w.n_cchunks * w.n_otiles;
    w.tiles_per_img = (int)ceil_div_ll(d.P, kMP);
    const long T = (long)d.B * w.tiles_per_img;
    long s = ceil_div_ll(8 * num_sms(), w.n_items);
    if (s < 1) s = 1;
    if (s > 128) s = 128;
    if (s > T) s = T;
    w.splits = (int)s;
    return w;
}

size_t mdcn_bwd_weight_mma_partial_bytes(const MdcnDims &d) {
    return align256_((size_t)plan_weight_mma(d).splits * d.Cout * d.Cin * d.K * sizeof(float));
}

// col / gout_nhwc: the modulated columns and the channels-last gout that mdcn_bwd_input_umma left in its workspace
int mdcn_bwd_weight_mma(const float *col, const float *gout_nhwc, float *partial, const MdcnDims &d,
                        int *splits_out, cudaStream_t stream) {
    const WeightMmaPlan w = plan_weight_mma(d);
    mdcn_bwd_weight_mma_kernel<<<dim3(w.n_items, w.splits), kMThreads, 0, stream>>>(
        col, gout_nhwc, partial, d, w.n_cchunks, w.n_otiles, w.tiles_per_img, w.splits);
    *splits_out = w.splits;
    return check_launch();
}

// dimensions of the column-gradient GEMM as a 1x1 convolution over the OUTPUT pixel grid
static int gcol_dims(const MdcnDims &d, MdcnDims &g) {
    return mdcn_make_dims(g, d.B, d.Cout, d.Ho, d.Wo, d.K * d.Cin, 1, 1, 1, 0, 1, 1, 1);
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

bool mdcn_bwd_umma_supported(const MdcnDims &d) {
    if (d.groups != 1 || d.Cin % 4 || d.Cout % 4 || d.Cd % 4) return false;
    // 8 lanes / 32-channel MMA chunks per (pixel, tap, deformable group): with fewer than 12 channels per deformable
    // group most of them idle and the FFMA kernels are faster (config-4 sweep: Cd = 8 is 10-15 % slower, Cd = 4 40 %)
    if (d.Cd < 12) return false;
    if ((long)d.K * d.Cin > 0xffff) return false;
    MdcnDims g;
    return gcol_dims(d, g) == AANET_OK && conv_umma_supported(g, false);
}

struct BwdUmmaWs {
    size_t gout_t, x_t, gx_t, gcol, col, wt, wpack, total;
};

static BwdUmmaWs plan_ws(const MdcnDims &d) {
    MdcnDims g;
    gcol_dims(d, g);
    BwdUmmaWs w{};
    size_t o = 0;
    w.gout_t = o; o += align256((size_t)d.B * d.P * d.Cout * sizeof(float));
    w.x_t = o;    o += align256((size_t)d.B * d.HW * d.Cin * sizeof(float));
    w.gx_t = o;   o += align256((size_t)d.B * d.HW * d.Cin * sizeof(float));
    w.gcol = o;   o += align256((size_t)d.B * d.P * d.K * d.Cin * sizeof(float));
    w.col = o;    o += align256((size_t)d.B * d.P * d.K * d.Cin * sizeof(float));
    w.wt = o;     o += align256((size_t)d.K * d.Cin * d.Cout * sizeof(float));
    w.wpack = o;  o += align256(conv_umma_wpack_bytes(g, 0));
    w.total = o;
    return w;
}

size_t mdcn_bwd_umma_workspace_bytes(const MdcnDims &d) { return plan_ws(d).total; }
const float *mdcn_bwd_umma_col(const MdcnDims &d, const void *ws) {
    return reinterpret_cast<const float *>(static_cast<const char *>(ws) + plan_ws(d).col);
}
const float *mdcn_bwd_umma_gout_nhwc(const MdcnDims &d, const void *ws) {
    return reinterpret_cast<const float *>(static_cast<const char *>(ws) + plan_ws(d).gout_t);
}

int mdcn_bwd_input_umma(const float *x, const float *offset, const float *mask, const float *weight,
                        const float *gout, float *gx, float *goffset, float *gmask, const MdcnDims &d, void *ws,
                        cudaStream_t stream) {
    const BwdUmmaWs w = plan_ws(d);
    char *base = static_cast<char *>(ws);
    float *gout_t = reinterpret_cast<float *>(base + w.gout_t), *x_t = reinterpret_cast<float *>(base + w.x_t);
    float *gx_t = reinterpret_cast<float *>(base + w.gx_t), *gcol = reinterpret_cast<float *>(base + w.gcol);
    float *col = reinterpret_cast<float *>(base + w.col);
    float *wt = reinterpret_cast<float *>(base + w.wt);
    void *wpack = base + w.wpack;
    MdcnDims g;
    int rc = gcol_dims(d, g);
    if (rc) return rc;
    if ((rc = conv_umma_transpose(gout, gout_t, d.B, d.Cout, d.P, stream)) != AANET_OK) return rc;
    if ((rc = conv_umma_transpose(x, x_t, d.B, d.Cin, d.HW, stream)) != AANET_OK) return rc;
    const long nw = (long)d.Cout * d.Cin * d.K;
    mdcn_wt_permute_kernel<<<(int)(ceil_div_ll(nw, 256) < 1184 ? ceil_div_ll(nw, 256) : 1184), 256, 0, stream>>>(
        weight, wt, d.Cout, d.Cin, d.K);
    if ((rc = check_launch()) != AANET_OK) return rc;
    if ((rc = conv_umma_pack(wt, wpack, g, 0, stream)) != AANET_OK) return rc;
    ConvParams p{};
    p.x = gout_t; p.wpack = static_cast<const float *>(wpack); p.out = gcol; p.d = g; p.act = ACT_NONE;
    if ((rc = conv_umma_launch(p, false, stream)) != AANET_OK) return rc;
    if (cudaMemsetAsync(gx_t, 0, sizeof(float) * (size_t)d.B * d.HW * d.Cin, stream) != cudaSuccess) return check_launch();
    const long n_items = (long)d.B * d.K * d.dg * d.P;
    const long blocks = ceil_div_ll(n_items * 8, 256);
    mdcn_bwd_scatter_kernel<<<(int)(blocks < 16L * num_sms() ? blocks : 16L * num_sms()), 256, 0, stream>>>(
        x_t, offset, mask, gcol, gx_t, goffset, gmask, col, d, n_items);
    if ((rc = check_launch()) != AANET_OK) return rc;
    return conv_umma_transpose(gx_t, gx, d.B, (int)d.HW, d.Cin, stream);
}

}  // namespace aanet
