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
                        float *__restrict__ goffset, float *__restrict__ gmask, MdcnDims d, long n_items) {
    const int lane8 = threadIdx.x & 7;
    const long item0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long stride_items = ((long)gridDim.x * blockDim.x) >> 3;
    const int cv = d.Cd >> 2;                               // 16-byte chunks per deformable group
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
                float t[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v1 = vv[0][e], v2 = vv[1][e], v3 = vv[2][e], v4 = vv[3][e];
                    a_m = fmaf(gcs[e], s.w[0] * v1 + s.w[1] * v2 + s.w[2] * v3 + s.w[3] * v4, a_m);
                    t[e] = gcs[e] * m;
                    a_oh = fmaf(t[e], -hw * v1 - s.lw * v2 + hw * v3 + s.lw * v4, a_oh);
                    a_ow = fmaf(t[e], -hh * v1 + hh * v2 - s.lh * v3 + s.lh * v4, a_ow);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((s.ok >> i) & 1)
                        red_add_v4(gx_b + (long)s.i[i] * d.Cin + ch * 4, t[0] * s.w[i], t[1] * s.w[i], t[2] * s.w[i],
                                   t[3] * s.w[i]);
            }
        }
#pragma unroll
        for (int sh = 4; sh > 0; sh >>= 1) {
            a_m += __shfl_xor_sync(0xffffffffu, a_m, sh);
            a_oh += __shfl_xor_sync(0xffffffffu, a_oh, sh);
            a_ow += __shfl_xor_sync(0xffffffffu, a_ow, sh);
        }
        if (lane8 == 0) {
            goffset[(b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 0) * d.P + p] = a_oh;
            goffset[(b * d.dg * 2 * d.K + (long)(g * d.K + k) * 2 + 1) * d.P + p] = a_ow;
            if (gmask) gmask[(b * d.dg * d.K + g * d.K + k) * d.P + p] = a_m;
        }
    }
}

// dimensions of the column-gradient GEMM as a 1x1 convolution over the OUTPUT pixel grid
static int gcol_dims(const MdcnDims &d, MdcnDims &g) {
    return mdcn_make_dims(g, d.B, d.Cout, d.Ho, d.Wo, d.K * d.Cin, 1, 1, 1, 0, 1, 1, 1);
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

bool mdcn_bwd_umma_supported(const MdcnDims &d) {
    if (d.groups != 1 || d.Cin % 4 || d.Cout % 4 || d.Cd % 4) return false;
    if ((long)d.K * d.Cin > 0xffff) return false;
    MdcnDims g;
    return gcol_dims(d, g) == AANET_OK && conv_umma_supported(g, false);
}

struct BwdUmmaWs {
    size_t gout_t, x_t, gx_t, gcol, wt, wpack, total;
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
    w.wt = o;     o += align256((size_t)d.K * d.Cin * d.Cout * sizeof(float));
    w.wpack = o;  o += align256(conv_umma_wpack_bytes(g, 0));
    w.total = o;
    return w;
}

size_t mdcn_bwd_umma_workspace_bytes(const MdcnDims &d) { return plan_ws(d).total; }

int mdcn_bwd_input_umma(const float *x, const float *offset, const float *mask, const float *weight,
                        const float *gout, float *gx, float *goffset, float *gmask, const MdcnDims &d, void *ws,
                        cudaStream_t stream) {
    const BwdUmmaWs w = plan_ws(d);
    char *base = static_cast<char *>(ws);
    float *gout_t = reinterpret_cast<float *>(base + w.gout_t), *x_t = reinterpret_cast<float *>(base + w.x_t);
    float *gx_t = reinterpret_cast<float *>(base + w.gx_t), *gcol = reinterpret_cast<float *>(base + w.gcol);
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
    mdcn_bwd_scatter_kernel<<<(int)(blocks < 16L * kNumSMs ? blocks : 16L * kNumSMs), 256, 0, stream>>>(
        x_t, offset, mask, gcol, gx_t, goffset, gmask, d, n_items);
    if ((rc = check_launch()) != AANET_OK) return rc;
    return conv_umma_transpose(gx_t, gx, d.B, (int)d.HW, d.Cin, stream);
}

}  // namespace aanet
