// Cross-scale aggregation fuse: resize + ordered sum + LeakyReLU in one pass (HBM-bound).
//
// Replaces the tail of AdaptiveAggregationModule.forward (reference nets/aggregation.py:387-400),
// which per output scale runs F.interpolate (one kernel per coarser term), two adds and an in-place
// LeakyReLU -- every intermediate makes a round trip through HBM.  Here each output element is
// produced once: same-size terms are read with 128-bit loads, coarser terms are sampled with ATen's
// align_corners=False bilinear rule (aggregation.py:395-396; src = (dst+0.5)*in/out - 0.5 clamped at
// 0, i1 = min(i0+1, in-1)), and the sum keeps the reference's order j = 0, 1, 2.
#include "common.cuh"

namespace aanet {

struct CsaTerms {
    const float *ptr[AANET_CSA_MAX_TERMS];
    int th[AANET_CSA_MAX_TERMS], tw[AANET_CSA_MAX_TERMS];
    int n;
};
struct CsaGrads {
    float *ptr[AANET_CSA_MAX_TERMS];
    int th[AANET_CSA_MAX_TERMS], tw[AANET_CSA_MAX_TERMS];
    int n;
};

// ATen area_pixel_compute_source_index + guard (float arithmetic, as for float tensors).
__device__ __forceinline__ void src_index(int dst, int in, float scale, int &i0, int &i1, float &l0, float &l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in - 1);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

__device__ __forceinline__ float sample_term(const float *__restrict__ src, int th, int tw, int H, int W,
                                             int h, int w) {
    int h0, h1, w0, w1; float a0, a1, b0, b1;
    src_index(h, th, (float)th / (float)H, h0, h1, a0, a1);
    src_index(w, tw, (float)tw / (float)W, w0, w1, b0, b1);
    const float *r0 = src + (long)h0 * tw, *r1 = src + (long)h1 * tw;
    return a0 * (b0 * __ldg(r0 + w0) + b1 * __ldg(r0 + w1)) + a1 * (b0 * __ldg(r1 + w0) + b1 * __ldg(r1 + w1));
}

template <int VEC>
__global__ void __launch_bounds__(256)
csa_fuse_fwd_kernel(CsaTerms t, float *__restrict__ out, int H, int W, long n_vec, float slope) {
    const int Wv = W / VEC;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long)gridDim.x * blockDim.x) {
        const int wv = (int)(i % Wv);
        const long r = i / Wv;
        const int h = (int)(r % H);
        const long bc = r / H;
        const int w = wv * VEC;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
            if (k >= t.n) break;
            float v[VEC];
            if (t.th[k] == H && t.tw[k] == W) {
                const float *p = t.ptr[k] + (bc * H + h) * W + w;
                if (VEC == 4) {
                    const float4 q = ldg_stream4(p);
                    v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
                } else {
                    v[0] = ldg_stream(p);
                }
            } else {
                const float *src = t.ptr[k] + bc * t.th[k] * t.tw[k];
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = sample_term(src, t.th[k], t.tw[k], H, W, h, w + e);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = (k == 0) ? v[e] : acc[e] + v[e];
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = acc[e] > 0.f ? acc[e] : acc[e] * slope;
        float *o = out + (bc * H + h) * W + w;
        if (VEC == 4) *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
        else o[0] = acc[0];
    }
}

// Channels-last variant for the fused inference path: terms and out are [B][h][w][C], C % 4 == 0.
// A thread owns kCsaChunks 16-byte channel chunks of ONE pixel (chunk q, q + Cv/kCsaChunks, ...), so the
// bilinear source indices / weights of the resized terms are computed once per thread instead of once per
// chunk (ncu on the one-chunk-per-thread version: 68 % issue utilisation, i.e. bound by that index math),
// while the lanes of a pixel still read consecutive 16-byte chunks.
constexpr int kCsaChunks = 4;

__global__ void __launch_bounds__(256)
csa_fuse_nhwc_kernel(CsaTerms t, float *__restrict__ out, int H, int W, int C, long n_items, int tpp, float slope) {
    const int Cv = C / 4;                 // 16-byte chunks per pixel; tpp = threads per pixel = ceil(Cv / kCsaChunks)
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % tpp);
        long r = i / tpp;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const long b = r / H;
        // per-term geometry: 4 source pixel offsets (in chunks) and 4 weights; same-size terms use slot 0 only
        long o00[AANET_CSA_MAX_TERMS], o01[AANET_CSA_MAX_TERMS], o10[AANET_CSA_MAX_TERMS], o11[AANET_CSA_MAX_TERMS];
        float w00[AANET_CSA_MAX_TERMS], w01[AANET_CSA_MAX_TERMS], w10[AANET_CSA_MAX_TERMS], w11[AANET_CSA_MAX_TERMS];
        bool same[AANET_CSA_MAX_TERMS];
#pragma unroll
        for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
            if (k >= t.n) break;
            const int th = t.th[k], tw = t.tw[k];
            const long base = b * th * tw;
            same[k] = (th == H && tw == W);
            if (same[k]) {
                o00[k] = (base + (long)h * W + w) * Cv;
            } else {
                int h0, h1, x0, x1; float a0, a1, b0, b1;
                src_index(h, th, (float)th / (float)H, h0, h1, a0, a1);
                src_index(w, tw, (float)tw / (float)W, x0, x1, b0, b1);
                o00[k] = (base + (long)h0 * tw + x0) * Cv; o01[k] = (base + (long)h0 * tw + x1) * Cv;
                o10[k] = (base + (long)h1 * tw + x0) * Cv; o11[k] = (base + (long)h1 * tw + x1) * Cv;
                w00[k] = a0 * b0; w01[k] = a0 * b1; w10[k] = a1 * b0; w11[k] = a1 * b1;
            }
        }
        float4 *orow = reinterpret_cast<float4 *>(out) + ((b * H + h) * W + w) * Cv;
#pragma unroll
        for (int u = 0; u < kCsaChunks; ++u) {
            const int cv = q + u * tpp;
            if (cv >= Cv) break;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
                if (k >= t.n) break;
                const float4 *src = reinterpret_cast<const float4 *>(t.ptr[k]) + cv;
                float4 v;
                if (same[k]) {
                    v = __ldg(src + o00[k]);
                } else {
                    // a0*(b0*v00 + b1*v01) + a1*(b0*v10 + b1*v11), evaluated with the products of the weights
                    const float4 v00 = __ldg(src + o00[k]), v01 = __ldg(src + o01[k]);
                    const float4 v10 = __ldg(src + o10[k]), v11 = __ldg(src + o11[k]);
                    v.x = w00[k] * v00.x + w01[k] * v01.x + w10[k] * v10.x + w11[k] * v11.x;
                    v.y = w00[k] * v00.y + w01[k] * v01.y + w10[k] * v10.y + w11[k] * v11.y;
                    v.z = w00[k] * v00.z + w01[k] * v01.z + w10[k] * v10.z + w11[k] * v11.z;
                    v.w = w00[k] * v00.w + w01[k] * v01.w + w10[k] * v10.w + w11[k] * v11.w;
                }
                if (k == 0) acc = v;
                else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            }
            acc.x = acc.x > 0.f ? acc.x : acc.x * slope; acc.y = acc.y > 0.f ? acc.y : acc.y * slope;
            acc.z = acc.z > 0.f ? acc.z : acc.z * slope; acc.w = acc.w > 0.f ? acc.w : acc.w * slope;
            orow[cv] = acc;
        }
    }
}

// Tiled channels-last variant (the one the fused executor hits): a block owns a kFuseTH x kFuseTW patch of
// output pixels.  The source patch of every RESIZED term (its bilinear footprint, a few coarse pixels) is
// staged in shared memory once, and the row / column interpolation entries (2 source indices + 2 weights)
// are tabulated once per tile, so an output chunk costs one coalesced global load per same-size term, 4
// LDS.128 per resized term and one store -- against 9 global loads per chunk plus per-thread index math in
// csa_fuse_nhwc_kernel (ncu: 22 us at the 1/3 scale, 23 % warps active at 115 registers, 33 % issue active).
constexpr int kFuseTH = 4, kFuseTW = 32, kFuseThreads = 256;

struct CsaTile {
    int rows_max[AANET_CSA_MAX_TERMS], cols_max[AANET_CSA_MAX_TERMS];   // patch capacity of each resized term
    int patch_off[AANET_CSA_MAX_TERMS];                                 // float4 offset of its patch in smem
    int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(kFuseThreads)
csa_fuse_tiled_kernel(CsaTerms t, CsaTile g, float *__restrict__ out, int H, int W, int C, float slope) {
    extern __shared__ float4 s_patch[];
    // interpolation entries: (i0, i1 as patch-relative indices, l0, l1) per tile row / column and term
    __shared__ int2 s_ri[AANET_CSA_MAX_TERMS][kFuseTH], s_ci[AANET_CSA_MAX_TERMS][kFuseTW];
    __shared__ float2 s_rw[AANET_CSA_MAX_TERMS][kFuseTH], s_cw[AANET_CSA_MAX_TERMS][kFuseTW];
    __shared__ int s_pw[AANET_CSA_MAX_TERMS];          // patch width (pixels) of each resized term

    const int Cv = C >> 2;
    int bt = blockIdx.x;
    const int tx = bt % g.tiles_x; bt /= g.tiles_x;
    const int ty = bt % g.tiles_y;
    const int b = bt / g.tiles_y;
    const int h0 = ty * kFuseTH, w0 = tx * kFuseTW;
    const int nh = min(kFuseTH, H - h0), nw = min(kFuseTW, W - w0);
    const int tid = threadIdx.x;
    pdl_wait();          // launched with programmatic stream serialization: the exchange convs may still be running

    // 1. tables (one thread per row / column entry and term)
#pragma unroll
    for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
        if (k >= t.n) break;
        const int th = t.th[k], tw = t.tw[k];
        if ((th == H && tw == W) || tid >= kFuseTH + kFuseTW) continue;
        int i0, i1, lo0, lo1; float l0, l1, u0, u1;
        if (tid < kFuseTH) {
            src_index(h0, th, (float)th / (float)H, lo0, lo1, u0, u1);          // first source row of the patch
            src_index(min(h0 + tid, H - 1), th, (float)th / (float)H, i0, i1, l0, l1);
            s_ri[k][tid] = make_int2(i0 - lo0, i1 - lo0);
            s_rw[k][tid] = make_float2(l0, l1);
        } else {
            const int c = tid - kFuseTH;
            src_index(w0, tw, (float)tw / (float)W, lo0, lo1, u0, u1);
            src_index(min(w0 + c, W - 1), tw, (float)tw / (float)W, i0, i1, l0, l1);
            s_ci[k][c] = make_int2(i0 - lo0, i1 - lo0);
            s_cw[k][c] = make_float2(l0, l1);
            if (c == 0) {
                int j0, j1;
                src_index(w0 + nw - 1, tw, (float)tw / (float)W, j0, j1, u0, u1);
                s_pw[k] = j1 - lo0 + 1;
            }
        }
    }
    // 2. source patches of the resized terms -> shared memory (rows are contiguous in channels-last memory)
#pragma unroll
    for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
        if (k >= t.n) break;
        const int th = t.th[k], tw = t.tw[k];
        if (th == H && tw == W) continue;
        int r_lo, r_hi, c_lo, c_hi, tmp; float f0, f1;
        src_index(h0, th, (float)th / (float)H, r_lo, tmp, f0, f1);
        src_index(h0 + nh - 1, th, (float)th / (float)H, tmp, r_hi, f0, f1);
        src_index(w0, tw, (float)tw / (float)W, c_lo, tmp, f0, f1);
        src_index(w0 + nw - 1, tw, (float)tw / (float)W, tmp, c_hi, f0, f1);
        const int pr = r_hi - r_lo + 1, pw = c_hi - c_lo + 1;
        if (pr > g.rows_max[k] || pw > g.cols_max[k]) __trap();          // host bound violated
        const int row_v = pw * Cv;                                       // float4 per patch row
        const float4 *src = reinterpret_cast<const float4 *>(t.ptr[k]) + ((long)b * th * tw + (long)r_lo * tw + c_lo) * Cv;
        float4 *dst = s_patch + g.patch_off[k];
        for (int i = tid; i < pr * row_v; i += kFuseThreads) {
            const int r = i / row_v, c = i - r * row_v;
            dst[i] = __ldg(src + (long)r * tw * Cv + c);
        }
    }
    __syncthreads();
    // 3. outputs: item = (pixel of the tile, 16-byte chunk); consecutive threads = consecutive chunks of a pixel
    const int items = nh * kFuseTW * Cv;
    // Cv a power of two (16 / 8 / 4 for the pyramid's 64 / 32 / 16 channels): item -> (pixel, chunk) by shift and mask
    const int cshift = (Cv & (Cv - 1)) == 0 ? 31 - __clz(Cv) : -1;
    for (int i = tid; i < items; i += kFuseThreads) {
        const int pix = cshift >= 0 ? i >> cshift : i / Cv, cv = i - pix * Cv;
        const int r = pix / kFuseTW, c = pix - r * kFuseTW;
        if (c >= nw) continue;
        const long opix = ((long)b * H + (h0 + r)) * W + (w0 + c);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
            if (k >= t.n) break;
            float4 v;
            if (t.th[k] == H && t.tw[k] == W) {
                v = __ldg(reinterpret_cast<const float4 *>(t.ptr[k]) + opix * Cv + cv);
            } else {
                const int2 ri = s_ri[k][r], ci = s_ci[k][c];
                const float2 rw = s_rw[k][r], cw = s_cw[k][c];
                const int pw = s_pw[k];
                const float4 *pp = s_patch + g.patch_off[k] + cv;
                const float4 v00 = pp[(ri.x * pw + ci.x) * Cv], v01 = pp[(ri.x * pw + ci.y) * Cv];
                const float4 v10 = pp[(ri.y * pw + ci.x) * Cv], v11 = pp[(ri.y * pw + ci.y) * Cv];
                const float w00 = rw.x * cw.x, w01 = rw.x * cw.y, w10 = rw.y * cw.x, w11 = rw.y * cw.y;
                v.x = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
                v.y = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
                v.z = w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
                v.w = w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
            }
            if (k == 0) acc = v;
            else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        }
        acc.x = acc.x > 0.f ? acc.x : acc.x * slope; acc.y = acc.y > 0.f ? acc.y : acc.y * slope;
        acc.z = acc.z > 0.f ? acc.z : acc.z * slope; acc.w = acc.w > 0.f ? acc.w : acc.w * slope;
        reinterpret_cast<float4 *>(out)[opix * Cv + cv] = acc;
    }
    pdl_trigger();       // the next kernel's launch latency overlaps this kernel's drain
}

// Backward, same-size term: g * LeakyReLU'(pre); sign(pre) == sign(out) because slope > 0.
__global__ void __launch_bounds__(256)
csa_bwd_same_kernel(const float *__restrict__ out, const float *__restrict__ gout, float *__restrict__ gt,
                    long n, float slope) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        gt[i] = gout[i] * (out[i] > 0.f ? 1.f : slope);
}

// Backward, resized term: adjoint of the bilinear resize in gather form.  For a source row hs the
// destination rows that reference it satisfy floor(src(h)) in {hs-1, hs}; a conservative window is
// scanned and every candidate is re-derived with the forward's exact arithmetic.
__global__ void __launch_bounds__(256)
csa_bwd_resize_kernel(const float *__restrict__ out, const float *__restrict__ gout, float *__restrict__ gt,
                      int th, int tw, int H, int W, long n_src, float slope) {
    const float sh = (float)th / (float)H, sw = (float)tw / (float)W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += (long)gridDim.x * blockDim.x) {
        const int ws = (int)(i % tw);
        const long r = i / tw;
        const int hs = (int)(r % th);
        const long bc = r / th;
        const int h_lo = max(0, (int)floorf(((float)hs - 0.5f) / sh - 0.5f) - 1);
        const int h_hi = min(H - 1, (int)ceilf(((float)hs + 1.5f) / sh - 0.5f) + 1);
        const int w_lo = max(0, (int)floorf(((float)ws - 0.5f) / sw - 0.5f) - 1);
        const int w_hi = min(W - 1, (int)ceilf(((float)ws + 1.5f) / sw - 0.5f) + 1);
        const float *o_b = out + bc * H * W, *g_b = gout + bc * H * W;
        float acc = 0.f;
        for (int h = h_lo; h <= h_hi; ++h) {
            int h0, h1; float a0, a1;
            src_index(h, th, sh, h0, h1, a0, a1);
            const float wh = (h0 == hs ? a0 : 0.f) + (h1 == hs ? a1 : 0.f);
            if (wh == 0.f) continue;
            float row = 0.f;
            for (int w = w_lo; w <= w_hi; ++w) {
                int w0, w1; float b0, b1;
                src_index(w, tw, sw, w0, w1, b0, b1);
                const float ww = (w0 == ws ? b0 : 0.f) + (w1 == ws ? b1 : 0.f);
                if (ww == 0.f) continue;
                const long q = (long)h * W + w;
                row = fmaf(ww, g_b[q] * (o_b[q] > 0.f ? 1.f : slope), row);
            }
            acc = fmaf(wh, row, acc);
        }
        gt[i] = acc;
    }
}

inline unsigned grid_for(long n, int block) {
    long g = ceil_div_ll(n, block);
    const long cap = (long)num_sms() * 16;   // grid-stride beyond 16 CTAs per SM
    return (unsigned)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace aanet

using namespace aanet;

extern "C" int aanet_csa_fuse_fwd(const float *const *terms, const int *th, const int *tw, int n_terms,
                                  float *out, int B, int C, int H, int W, float slope, void *stream) {
    if (!terms || !th || !tw || !out) return AANET_ERR_NULL;
    if (n_terms < 1 || n_terms > AANET_CSA_MAX_TERMS || B <= 0 || C <= 0 || H <= 0 || W <= 0)
        return AANET_ERR_SHAPE;
    CsaTerms t;
    t.n = n_terms;
    bool vec = (W % 4 == 0) && aligned16(out);
    for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
        t.ptr[k] = nullptr; t.th[k] = t.tw[k] = 0;
        if (k >= n_terms) continue;
        if (!terms[k]) return AANET_ERR_NULL;
        if (th[k] <= 0 || tw[k] <= 0) return AANET_ERR_SHAPE;
        t.ptr[k] = terms[k]; t.th[k] = th[k]; t.tw[k] = tw[k];
        if (th[k] == H && tw[k] == W && !aligned16(terms[k])) vec = false;
    }
    const long n = (long)B * C * H * W;
    if (vec)
        csa_fuse_fwd_kernel<4><<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(t, out, H, W, n / 4, slope);
    else
        csa_fuse_fwd_kernel<1><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(t, out, H, W, n, slope);
    return check_launch();
}

extern "C" int aanet_csa_fuse_nhwc(const float *const *terms, const int *th, const int *tw, int n_terms,
                                   float *out, int B, int C, int H, int W, float slope, void *stream) {
    if (!terms || !th || !tw || !out) return AANET_ERR_NULL;
    if (n_terms < 1 || n_terms > AANET_CSA_MAX_TERMS || B <= 0 || C <= 0 || H <= 0 || W <= 0)
        return AANET_ERR_SHAPE;
    if (C % 4 || !aligned16(out)) return AANET_ERR_UNSUPPORTED;
    CsaTerms t;
    t.n = n_terms;
    for (int k = 0; k < AANET_CSA_MAX_TERMS; ++k) {
        t.ptr[k] = nullptr; t.th[k] = t.tw[k] = 0;
        if (k >= n_terms) continue;
        if (!terms[k]) return AANET_ERR_NULL;
        if (th[k] <= 0 || tw[k] <= 0) return AANET_ERR_SHAPE;
        if (!aligned16(terms[k])) return AANET_ERR_UNSUPPORTED;
        t.ptr[k] = terms[k]; t.th[k] = th[k]; t.tw[k] = tw[k];
    }
    // Tiled kernel when every resized term is an upsampling whose per-tile source patches fit in shared memory.
    CsaTile g{};
    g.tiles_x = ceil_div(W, kFuseTW); g.tiles_y = ceil_div(H, kFuseTH);
    long smem_v = 0;                                  // float4 units
    bool tiled = (long)B * g.tiles_x * g.tiles_y <= 0x7fffffffL;
    for (int k = 0; k < n_terms && tiled; ++k) {
        if (th[k] == H && tw[k] == W) continue;
        if (th[k] > H || tw[k] > W) { tiled = false; break; }
        // bilinear footprint of kFuseTH (kFuseTW) consecutive destination pixels: span * scale + both neighbours
        g.rows_max[k] = (int)((long)(kFuseTH - 1) * th[k] / H) + 3;
        g.cols_max[k] = (int)((long)(kFuseTW - 1) * tw[k] / W) + 3;
        g.patch_off[k] = (int)smem_v;
        smem_v += (long)g.rows_max[k] * g.cols_max[k] * (C / 4);
    }
    if (tiled && smem_v * 16 <= 96 * 1024) {
        const size_t smem = (size_t)smem_v * 16;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(csa_fuse_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return launch_pdl(csa_fuse_tiled_kernel, dim3(B * g.tiles_x * g.tiles_y), dim3(kFuseThreads), smem,
                          as_stream(stream), t, g, out, H, W, C, slope);
    }
    const int tpp = ceil_div(C / 4, kCsaChunks);
    const long n = (long)B * H * W * tpp;
    csa_fuse_nhwc_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(t, out, H, W, C, n, tpp, slope);
    return check_launch();
}

extern "C" int aanet_csa_fuse_bwd(const float *out, const float *gout, float *const *gterms, const int *th,
                                  const int *tw, int n_terms, int B, int C, int H, int W, float slope,
                                  void *stream) {
    if (!out || !gout || !gterms || !th || !tw) return AANET_ERR_NULL;
    if (n_terms < 1 || n_terms > AANET_CSA_MAX_TERMS || B <= 0 || C <= 0 || H <= 0 || W <= 0)
        return AANET_ERR_SHAPE;
    if (!(slope > 0.f)) return AANET_ERR_UNSUPPORTED;
    for (int k = 0; k < n_terms; ++k) {
        if (!gterms[k]) continue;           // gradient not requested for this term
        if (th[k] <= 0 || tw[k] <= 0) return AANET_ERR_SHAPE;
        if (th[k] == H && tw[k] == W) {
            const long n = (long)B * C * H * W;
            csa_bwd_same_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(out, gout, gterms[k], n, slope);
        } else {
            const long n = (long)B * C * th[k] * tw[k];
            csa_bwd_resize_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(out, gout, gterms[k], th[k],
                                                                                  tw[k], H, W, n, slope);
        }
        const int rc = check_launch();
        if (rc) return rc;
    }
    return AANET_OK;
}
