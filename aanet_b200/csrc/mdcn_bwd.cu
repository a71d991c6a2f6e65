// Modulated deformable convolution backward (training path, BASELINE config 4).
//
// Replaces modulated_deform_conv_cuda_backward (reference deform_conv_cuda.cpp:571-685), which per
// image runs: columns = W^T.gout (cuBLAS, cpp:623-626) -> col2im_coord (cu:695-767) -> col2im with
// atomics (cu:635-693) -> im2col again (cpp:647-650) -> gW += gout.columns^T (cpp:659-664), moving the
// [Cin*9, Ho*Wo] columns buffer through HBM four times.  Here nothing column-shaped touches HBM:
//
//  * mdcn_bwd_input_kernel: one thread per output pixel; the CTA keeps its gout tile [Cout x 128]
//    in shared memory, forms the column gradient cg[c] = sum_o W[o,c,k]*gout[o,p] for 8 channels at
//    a time in registers and immediately consumes it: grad_mask (cu:753), grad_offset through the
//    bilinear coordinate weights (cu:526-568, :758) and grad_input scattered to the four corners with
//    float atomics (cu:499-524, :688).
//  * mdcn_bwd_weight_kernel: a CTA owns (tap, 32 input channels, 64 output channels) and a strided
//    share of the pixel tiles; it re-gathers the modulated columns for its slice into shared memory
//    and reduces gout.col^T in registers; per-split partials go to the workspace and are summed in a
//    fixed order (deterministic grad_weight, like the reference's cuBLAS path).
//  * mdcn_bwd_bias_kernel: grad_bias = sum_{b,p} gout (cpp:665-671).
#include "mdcn_common.cuh"

namespace aanet {

constexpr int kBP = 128;        // pixels (threads) per CTA, input-gradient kernel
constexpr int kBC = 8;          // channels per register chunk

__global__ void __launch_bounds__(kBP)
mdcn_bwd_input_kernel(const float *__restrict__ x, const float *__restrict__ offset,
                      const float *__restrict__ mask, const float *__restrict__ weight,
                      const float *__restrict__ gout, float *__restrict__ gx,
                      float *__restrict__ goffset, float *__restrict__ gmask, MdcnDims d) {
    extern __shared__ __align__(16) float smem[];
    float *s_g = smem;                          // [Cout][kBP]
    float *s_w = smem + (size_t)d.Cout * kBP;   // [Og][kBC]

    const int tid = threadIdx.x, b = blockIdx.y;
    const long p0 = (long)blockIdx.x * kBP, p = p0 + tid;
    const bool p_ok = p < d.P;
    const long pc = p_ok ? p : 0;
    const int ho = (int)(pc / d.Wo), wo = (int)(pc % d.Wo);

    for (int o = 0; o < d.Cout; ++o)
        s_g[o * kBP + tid] = p_ok ? gout[((long)b * d.Cout + o) * d.P + p] : 0.f;

    const float *off_b = offset + (long)b * d.dg * 2 * d.K * d.P;
    const float *mask_b = mask ? mask + (long)b * d.dg * d.K * d.P : nullptr;
    const float *x_b = x + (long)b * d.Cin * d.HW;
    float *gx_b = gx + (long)b * d.Cin * d.HW;

    for (int k = 0; k < d.K; ++k)
        for (int g = 0; g < d.dg; ++g) {
            const Sample s = sample_at(d, off_b, g, k, ho, wo, pc);
            const float m = mask_b ? mask_b[(long)(g * d.K + k) * d.P + pc] : 1.f;
            const float hh = 1.f - s.lh, hw = 1.f - s.lw;
            float a_oh = 0.f, a_ow = 0.f, a_m = 0.f;
            int c = g * d.Cd;
            const int g_end = (g + 1) * d.Cd;
            while (c < g_end) {
                const int grp = c / d.Cg;
                const int nc = min(kBC, min(g_end, (grp + 1) * d.Cg) - c);
                __syncthreads();            // previous chunk's s_w readers are done (also covers s_g fill)
                for (int i = tid; i < d.Og * kBC; i += kBP) {
                    const int o = i / kBC, cc = i % kBC;
                    s_w[i] = (cc < nc)
                                 ? __ldg(weight + ((long)(grp * d.Og + o) * d.Cg + (c - grp * d.Cg + cc)) * d.K + k)
                                 : 0.f;
                }
                __syncthreads();
                float cg[kBC];
#pragma unroll
                for (int cc = 0; cc < kBC; ++cc) cg[cc] = 0.f;
                for (int o = 0; o < d.Og; ++o) {
                    const float gv = s_g[(grp * d.Og + o) * kBP + tid];
                    const float4 w0 = *reinterpret_cast<const float4 *>(&s_w[o * kBC]);
                    const float4 w1 = *reinterpret_cast<const float4 *>(&s_w[o * kBC + 4]);
                    cg[0] = fmaf(gv, w0.x, cg[0]); cg[1] = fmaf(gv, w0.y, cg[1]);
                    cg[2] = fmaf(gv, w0.z, cg[2]); cg[3] = fmaf(gv, w0.w, cg[3]);
                    cg[4] = fmaf(gv, w1.x, cg[4]); cg[5] = fmaf(gv, w1.y, cg[5]);
                    cg[6] = fmaf(gv, w1.z, cg[6]); cg[7] = fmaf(gv, w1.w, cg[7]);
                }
                if (s.valid && p_ok) {
#pragma unroll
                    for (int cc = 0; cc < kBC; ++cc) {
                        if (cc >= nc) break;
                        const float *im = x_b + (long)(c + cc) * d.HW;
                        float *gi = gx_b + (long)(c + cc) * d.HW;
                        const float v1 = (s.ok & 1) ? __ldg(im + s.i[0]) : 0.f;
                        const float v2 = (s.ok & 2) ? __ldg(im + s.i[1]) : 0.f;
                        const float v3 = (s.ok & 4) ? __ldg(im + s.i[2]) : 0.f;
                        const float v4 = (s.ok & 8) ? __ldg(im + s.i[3]) : 0.f;
                        a_m = fmaf(cg[cc], s.w[0] * v1 + s.w[1] * v2 + s.w[2] * v3 + s.w[3] * v4, a_m);
                        const float t = cg[cc] * m;
                        a_oh = fmaf(t, -hw * v1 - s.lw * v2 + hw * v3 + s.lw * v4, a_oh);
                        a_ow = fmaf(t, -hh * v1 + hh * v2 - s.lh * v3 + s.lh * v4, a_ow);
                        if (s.ok & 1) atomicAdd(gi + s.i[0], t * s.w[0]);
                        if (s.ok & 2) atomicAdd(gi + s.i[1], t * s.w[1]);
                        if (s.ok & 4) atomicAdd(gi + s.i[2], t * s.w[2]);
                        if (s.ok & 8) atomicAdd(gi + s.i[3], t * s.w[3]);
                    }
                }
                c += nc;
            }
            if (p_ok) {
                goffset[((long)b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 0) * d.P + p] = a_oh;
                goffset[((long)b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 1) * d.P + p] = a_ow;
                if (gmask) gmask[((long)b * d.dg * d.K + g * d.K + k) * d.P + p] = a_m;
            }
        }
}

// ------------------------------------------------------------------------------------------------
constexpr int kWP = 64;          // pixels per tile
constexpr int kWC = 32;          // input channels per CTA
constexpr int kWO = 64;          // output channels per CTA
constexpr int kWThreads = 128;   // (kWC/4) x (kWO/4) register tiles of 4x4
constexpr int kColStride = kWC + 4;   // 36: 16B-aligned rows, 4-way store conflicts instead of 32-way
constexpr int kGStride = kWO + 4;     // 68

struct WeightPlan {
    int n_cchunks;      // channel chunks per conv group (chunks never straddle a deformable group)
    int n_otiles;       // output tiles per conv group
    int n_items;        // groups * K * n_cchunks * n_otiles
    int tiles_per_img;  // ceil(P / kWP)
    int splits;         // pixel splits S
};

inline WeightPlan make_weight_plan(const MdcnDims &d) {
    WeightPlan w;
    // chunks per conv group: cut every kWC channels and at every deformable-group edge; take the
    // maximum over groups (items beyond a group's own count resolve to nc == 0 and exit)
    int n_max = 0;
    for (int grp = 0; grp < d.groups; ++grp) {
        int n = 0, c = grp * d.Cg;
        const int end = c + d.Cg;
        while (c < end) {
            const int g = c / d.Cd;
            const int lim = (end < (g + 1) * d.Cd) ? end : (g + 1) * d.Cd;
            c += (lim - c < kWC) ? (lim - c) : kWC;
            ++n;
        }
        if (n > n_max) n_max = n;
    }
    w.n_cchunks = n_max;
    w.n_otiles = ceil_div(d.Og, kWO);
    w.n_items = d.groups * d.K * w.n_cchunks * w.n_otiles;
    w.tiles_per_img = (int)ceil_div_ll(d.P, kWP);
    const long T = (long)d.B * w.tiles_per_img;
    long s = ceil_div_ll(4 * num_sms(), w.n_items);
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    if (s > T) s = T;
    w.splits = (int)s;
    return w;
}

// Resolve chunk index `ci` (0..n_cchunks) of conv group `grp` to a channel range [c, c+nc) that
// lies inside one deformable group; nc == 0 for the padding entries of the upper bound.
__device__ __forceinline__ void chunk_range(const MdcnDims &d, int grp, int ci, int &c_out, int &nc_out) {
    int c = grp * d.Cg;
    const int end = c + d.Cg;
    int idx = 0;
    nc_out = 0; c_out = c;
    while (c < end) {
        const int g = c / d.Cd;
        const int nc = min(kWC, min(end, (g + 1) * d.Cd) - c);
        if (idx == ci) { c_out = c; nc_out = nc; return; }
        c += nc; ++idx;
    }
}

__global__ void __launch_bounds__(kWThreads)
mdcn_bwd_weight_kernel(const float *__restrict__ x, const float *__restrict__ offset,
                       const float *__restrict__ mask, const float *__restrict__ gout,
                       float *__restrict__ partial, MdcnDims d, WeightPlan wp) {
    __shared__ __align__(16) float s_col[kWP][kColStride];
    __shared__ __align__(16) float s_g[kWP][kGStride];

    int item = blockIdx.x;
    const int ot = item % wp.n_otiles; item /= wp.n_otiles;
    const int ci = item % wp.n_cchunks; item /= wp.n_cchunks;
    const int k = item % d.K;
    const int grp = item / d.K;
    const int split = blockIdx.y;

    int c0, nc;
    chunk_range(d, grp, ci, c0, nc);
    if (nc == 0) return;                      // padding item: its partial slice is never read
    const int g = c0 / d.Cd;
    const int o0 = ot * kWO;

    const int tid = threadIdx.x;
    const int tc = (tid >> 4) * 4, to = (tid & 15) * 4;   // register tile origin (channel, out)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int gp = tid % kWP, gh = tid / kWP;     // gather role: pixel, channel half (16 each)
    const long T = (long)d.B * wp.tiles_per_img;
    for (long t = split; t < T; t += wp.splits) {
        const int b = (int)(t / wp.tiles_per_img);
        const long p0 = (t % wp.tiles_per_img) * kWP;
        const long p = p0 + gp;
        const bool p_ok = p < d.P;
        const long pc = p_ok ? p : 0;
        const float *off_b = offset + (long)b * d.dg * 2 * d.K * d.P;
        const Sample s = sample_at(d, off_b, g, k, (int)(pc / d.Wo), (int)(pc % d.Wo), pc);
        const float m = mask ? mask[((long)b * d.dg * d.K + g * d.K + k) * d.P + pc] : 1.f;
        const float *x_b = x + (long)b * d.Cin * d.HW;
#pragma unroll 4
        for (int cc = gh * (kWC / 2); cc < (gh + 1) * (kWC / 2); ++cc) {
            float v = 0.f;
            if (cc < nc && p_ok) {
                const float *im = x_b + (long)(c0 + cc) * d.HW;
                v = (s.w[0] * __ldg(im + s.i[0]) + s.w[1] * __ldg(im + s.i[1]) +
                     s.w[2] * __ldg(im + s.i[2]) + s.w[3] * __ldg(im + s.i[3])) * m;
            }
            s_col[gp][cc] = v;
        }
        for (int i = tid; i < kWP * kWO; i += kWThreads) {
            const int o = i / kWP, px = i % kWP;
            const bool ok = (o0 + o < d.Og) && (p0 + px < d.P);
            s_g[px][o] = ok ? gout[((long)b * d.Cout + grp * d.Og + o0 + o) * d.P + p0 + px] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int px = 0; px < kWP; ++px) {
            const float4 cv = *reinterpret_cast<const float4 *>(&s_col[px][tc]);
            const float4 gv = *reinterpret_cast<const float4 *>(&s_g[px][to]);
            const float c4[4] = {cv.x, cv.y, cv.z, cv.w}, g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(c4[i], g4[j], acc[i][j]);
        }
        __syncthreads();
    }
    // partial[split][o][cl][k]
    float *dst = partial + (size_t)split * d.Cout * d.Cg * d.K;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = tc + i, o = o0 + to + j;
            if (cc < nc && o < d.Og)
                dst[((size_t)(grp * d.Og + o) * d.Cg + (c0 - grp * d.Cg + cc)) * d.K + k] = acc[i][j];
        }
}

__global__ void mdcn_bwd_weight_reduce_kernel(const float *__restrict__ partial, float *__restrict__ gw,
                                              int n, int splits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * n + i];
    gw[i] = a;
}

__global__ void __launch_bounds__(256)
mdcn_bwd_bias_kernel(const float *__restrict__ gout, float *__restrict__ gbias, int B, int Cout, long P) {
    __shared__ float red[256];
    const int o = blockIdx.x;
    float a = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *row = gout + ((long)b * Cout + o) * P;
        for (long p = threadIdx.x; p < P; p += 256) a += row[p];
    }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) gbias[o] = red[0];
}

// mdcn_bwd_umma.cu: input / offset / mask gradients through the tcgen05 engine + vector atomics
bool mdcn_bwd_umma_supported(const MdcnDims &d);
size_t mdcn_bwd_umma_workspace_bytes(const MdcnDims &d);
int mdcn_bwd_input_umma(const float *x, const float *offset, const float *mask, const float *weight,
                        const float *gout, float *gx, float *goffset, float *gmask, const MdcnDims &d, void *ws,
                        cudaStream_t stream);

size_t mdcn_bwd_weight_mma_partial_bytes(const MdcnDims &d);
int mdcn_bwd_weight_mma(const float *col, const float *gout_nhwc, float *partial, const MdcnDims &d,
                        int *splits_out, cudaStream_t stream);
const float *mdcn_bwd_umma_col(const MdcnDims &d, const void *ws);
const float *mdcn_bwd_umma_gout_nhwc(const MdcnDims &d, const void *ws);

// room for the per-split partial weight gradients of whichever weight kernel runs
static size_t weight_partial_bytes(const MdcnDims &d) {
    const WeightPlan wp = make_weight_plan(d);
    size_t n = (((size_t)wp.splits * d.Cout * d.Cg * d.K * sizeof(float)) + 255) & ~(size_t)255;
    if (mdcn_bwd_umma_supported(d)) {
        const size_t m = mdcn_bwd_weight_mma_partial_bytes(d);
        n = n > m ? n : m;
    }
    return n;
}

size_t mdcn_bwd_workspace_bytes(const MdcnDims &d) {
    return weight_partial_bytes(d) + (mdcn_bwd_umma_supported(d) ? mdcn_bwd_umma_workspace_bytes(d) : 0);
}

int mdcn_bwd_launch(const float *x, const float *offset, const float *mask, const float *weight,
                    const float *gout, float *gx, float *goffset, float *gmask, float *gweight,
                    float *gbias, const MdcnDims &d, void *ws, size_t ws_bytes, cudaStream_t stream) {
    const WeightPlan wp = make_weight_plan(d);
    const size_t need = (size_t)wp.splits * d.Cout * d.Cg * d.K * sizeof(float);
    if (!ws || ws_bytes < need) return AANET_ERR_WORKSPACE;
    int rc;

    // ---- grad_input / grad_offset / grad_mask
    const size_t part = weight_partial_bytes(d);
    if (mdcn_bwd_umma_supported(d) && ws_bytes >= part + mdcn_bwd_umma_workspace_bytes(d)) {
        rc = mdcn_bwd_input_umma(x, offset, mask, weight, gout, gx, goffset, gmask, d, static_cast<char *>(ws) + part,
                                 stream);
        if (rc) return rc;
    } else {
        const size_t smem = ((size_t)d.Cout * kBP + (size_t)d.Og * kBC) * sizeof(float);
        if (smem > 200 * 1024) return AANET_ERR_UNSUPPORTED;
        cudaFuncSetAttribute(mdcn_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)d.B * d.Cin * d.HW, stream) != cudaSuccess)
            return check_launch();
        const long n_ptiles = ceil_div_ll(d.P, kBP);
        mdcn_bwd_input_kernel<<<dim3((unsigned)n_ptiles, d.B), kBP, smem, stream>>>(
            x, offset, mask, weight, gout, gx, goffset, gmask, d);
        rc = check_launch();
        if (rc) return rc;
    }

    // ---- grad_weight
    float *partial = static_cast<float *>(ws);
    const int n = d.Cout * d.Cg * d.K;
    int splits = wp.splits;
    if (mdcn_bwd_umma_supported(d) && ws_bytes >= part + mdcn_bwd_umma_workspace_bytes(d)) {
        const void *uws = static_cast<char *>(ws) + part;      // the modulated columns and channels-last gout are there
        rc = mdcn_bwd_weight_mma(mdcn_bwd_umma_col(d, uws), mdcn_bwd_umma_gout_nhwc(d, uws), partial, d, &splits, stream);
    } else {
        // padding items (nc == 0) leave holes only in slices that are never read; real slices are
        // fully written because every (o, c, k) belongs to exactly one item.
        mdcn_bwd_weight_kernel<<<dim3(wp.n_items, wp.splits), kWThreads, 0, stream>>>(x, offset, mask, gout,
                                                                                     partial, d, wp);
        rc = check_launch();
    }
    if (rc) return rc;
    mdcn_bwd_weight_reduce_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(partial, gweight, n, splits);
    rc = check_launch();
    if (rc) return rc;

    if (gbias) {
        mdcn_bwd_bias_kernel<<<d.Cout, 256, 0, stream>>>(gout, gbias, d.B, d.Cout, d.P);
        rc = check_launch();
    }
    return rc;
}

}  // namespace aanet
