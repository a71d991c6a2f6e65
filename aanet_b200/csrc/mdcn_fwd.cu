// Modulated deformable convolution forward -- generic fused kernel (any Cin/Cout/groups/dg/k).
//
// Replaces modulated_deform_conv_cuda_forward (reference deform_conv_cuda.cpp:490-569): per image
// the reference writes a [Cin*9, Ho*Wo] `columns` buffer with modulated_deformable_im2col_gpu_kernel
// (cu:570-633), reads it back in a cuBLAS SGEMM (cpp:550-555) and pre-zeroes the output (cpp:530)
// -- >= 370 MB of HBM traffic for 38.9 MB algorithmic at the 1/3 scale.  Here the sampled columns
// never leave the SM: a CTA owns 128 output pixels x 64 output channels; for each (tap, 16-channel
// chunk) its threads bilinear-gather a [16 x 128] column tile straight into shared memory next to
// the matching [16 x 64] weight slice and accumulate an 8x8 register tile per thread with FFMA.
// Bias / folded-BN scale+shift / ReLU are applied in the epilogue; the whole batch is one launch.
//
// The tcgen05 (3xTF32) specialisation for the hot ISA shapes lives in mdcn_fwd_umma.cu; this
// kernel is the shape-generic path and the one the backward's column recomputation mirrors.
#include "mdcn_common.cuh"

namespace aanet {

constexpr int kFP = 128;   // pixels per CTA
constexpr int kFN = 64;    // output channels per CTA
constexpr int kFC = 16;    // input channels per smem stage
constexpr int kFThreads = 128;

__global__ void __launch_bounds__(kFThreads)
mdcn_fwd_kernel(const float *__restrict__ x, const float *__restrict__ offset,
                const float *__restrict__ mask, const float *__restrict__ weight,
                const float *__restrict__ bias, const float *__restrict__ post_scale,
                const float *__restrict__ post_shift, int relu, float *__restrict__ out, MdcnDims d) {
    __shared__ __align__(16) float s_col[kFC][kFP];
    __shared__ __align__(16) float s_w[kFC][kFN];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int n_otiles = ceil_div(d.Og, kFN);
    const int grp = blockIdx.y / n_otiles;            // conv group
    const int o0 = (blockIdx.y % n_otiles) * kFN;     // first out channel (within group) of the tile
    const long p0 = (long)blockIdx.x * kFP;
    const long p = p0 + tid;                          // this thread's gather pixel
    const bool p_ok = p < d.P;
    const int ho = p_ok ? (int)(p / d.Wo) : 0, wo = p_ok ? (int)(p % d.Wo) : 0;
    const long pc = p_ok ? p : 0;

    const float *off_b = offset + (long)b * d.dg * 2 * d.K * d.P;
    const float *mask_b = mask ? mask + (long)b * d.dg * d.K * d.P : nullptr;
    const float *x_b = x + (long)b * d.Cin * d.HW;

    const int tp = (tid & 15) * 8, tn = (tid >> 4) * 8;   // register tile origin
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int c_begin = grp * d.Cg, c_end = c_begin + d.Cg;
    for (int k = 0; k < d.K; ++k) {
        int c = c_begin;
        while (c < c_end) {
            const int g = c / d.Cd;                              // deformable group of this run
            const int run_end = min(c_end, (g + 1) * d.Cd);
            // sampling geometry for (pixel, tap, deformable group): once per run
            Sample s = sample_at(d, off_b, g, k, ho, wo, pc);
            const float m = mask_b ? mask_b[(long)(g * d.K + k) * d.P + pc] : 1.f;
            for (; c < run_end; c += kFC) {
                const int nc = min(kFC, run_end - c);
                // gather the column tile
#pragma unroll 4
                for (int cc = 0; cc < kFC; ++cc) {
                    float v = 0.f;
                    if (cc < nc && p_ok) {
                        const float *im = x_b + (long)(c + cc) * d.HW;
                        v = s.w[0] * __ldg(im + s.i[0]) + s.w[1] * __ldg(im + s.i[1]) +
                            s.w[2] * __ldg(im + s.i[2]) + s.w[3] * __ldg(im + s.i[3]);
                        v *= m;
                    }
                    s_col[cc][tid] = v;
                }
                // weight slice W[o0.., c-c_begin.., k] -> s_w[cc][o]
                for (int i = tid; i < kFC * kFN; i += kFThreads) {
                    const int cc = i / kFN, o = i % kFN;
                    float wv = 0.f;
                    if (cc < nc && o0 + o < d.Og)
                        wv = __ldg(weight + ((long)(grp * d.Og + o0 + o) * d.Cg + (c - c_begin + cc)) * d.K + k);
                    s_w[cc][o] = wv;
                }
                __syncthreads();
#pragma unroll
                for (int cc = 0; cc < kFC; ++cc) {
                    float a[8], w8[8];
                    *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&s_col[cc][tp]);
                    *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&s_col[cc][tp + 4]);
                    *reinterpret_cast<float4 *>(&w8[0]) = *reinterpret_cast<const float4 *>(&s_w[cc][tn]);
                    *reinterpret_cast<float4 *>(&w8[4]) = *reinterpret_cast<const float4 *>(&s_w[cc][tn + 4]);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w8[j], acc[i][j]);
                }
                __syncthreads();
            }
            c = run_end;
        }
    }

    // epilogue: bias, optional per-channel affine + ReLU, store [o][p] (pixels contiguous)
    const bool vec_ok = (d.P % 4 == 0) && aligned16(out);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ol = o0 + tn + j;
        if (ol >= d.Og) continue;
        const int o = grp * d.Og + ol;
        const float bv = bias ? bias[o] : 0.f;
        const float sc = post_scale ? post_scale[o] : 1.f;
        const float sh = post_shift ? post_shift[o] : 0.f;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = acc[i][j] + bv;
            if (post_scale) t = fmaf(t, sc, sh);
            v[i] = relu ? fmaxf(t, 0.f) : t;
        }
        float *orow = out + ((long)b * d.Cout + o) * d.P;
        const long pp = p0 + tp;
        if (vec_ok && pp + 8 <= d.P) {
            *reinterpret_cast<float4 *>(orow + pp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(orow + pp + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (pp + i < d.P) orow[pp + i] = v[i];
        }
    }
}

int mdcn_fwd_generic(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, const MdcnDims &d, const float *post_scale,
                     const float *post_shift, int relu, cudaStream_t stream) {
    const long n_ptiles = ceil_div_ll(d.P, kFP);
    const int n_otiles = ceil_div(d.Og, kFN) * d.groups;
    if (n_ptiles > 2147483647LL || n_otiles > 65535) return AANET_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)n_ptiles, n_otiles, d.B);
    mdcn_fwd_kernel<<<grid, kFThreads, 0, stream>>>(x, offset, mask, weight, bias, post_scale,
                                                    post_shift, relu, out, d);
    return check_launch();
}

}  // namespace aanet
