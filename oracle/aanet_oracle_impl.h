/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the AANet hot path.
 *
 * This header is the type-generic body of oracle/aanet_oracle.c.  It is
 * included twice, once with REAL=float (suffix _f32) and once with
 * REAL=double (suffix _f64).  Storage, coordinate arithmetic and the bilinear
 * weights are done in REAL exactly as the reference does them in scalar_t;
 * long sums are accumulated in double so that the oracle sits closer to the
 * exact value than either implementation under test.
 *
 * Nothing under aanet_b200/ may include, link or call this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it (through oracle/oracle.py).
 *
 * Reference citations are relative to /root/reference/.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ------------------------------------------------------------------ */
/* Correlation cost volume.  nets/cost.py:40-48                        */
/*   cost[b,d,h,w] = (1/C) sum_c L[b,c,h,w] * R[b,c,h,w-d]   (w >= d)   */
/*   cost[b,d,h,w] = 0                                       (w <  d)   */
/* (the volume is created with new_zeros, cost.py:41, and only the      */
/*  slice [:, i, :, i:] is written, cost.py:45)                         */
/* ------------------------------------------------------------------ */
void FN(orc_corr_fwd)(const REAL *L, const REAL *R, REAL *cost,
                      int B, int C, int H, int W, int D)
{
    const long HW = (long)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h) {
            for (int d = 0; d < D; ++d) {
                REAL *o = cost + (((long)b * D + d) * H + h) * W;
                for (int w = 0; w < W; ++w) {
                    if (w < d) { o[w] = 0; continue; }
                    double acc = 0.0;
                    const REAL *l = L + ((long)b * C) * HW + (long)h * W + w;
                    const REAL *r = R + ((long)b * C) * HW + (long)h * W + (w - d);
                    for (int c = 0; c < C; ++c)
                        acc += (double)l[c * HW] * (double)r[c * HW];
                    o[w] = (REAL)(acc / C);
                }
            }
        }
}

/* Autograd of cost.py:45-48:
 *   gL[b,c,h,w]  = (1/C) sum_{d<=w, d<D}   g[b,d,h,w]    * R[b,c,h,w-d]
 *   gR[b,c,h,w'] = (1/C) sum_{d<D, w'+d<W} g[b,d,h,w'+d] * L[b,c,h,w'+d]   */
void FN(orc_corr_bwd)(const REAL *L, const REAL *R, const REAL *g,
                      REAL *gL, REAL *gR, int B, int C, int H, int W, int D)
{
    const long HW = (long)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const REAL *l = L + ((long)b * C + c) * HW;
            const REAL *r = R + ((long)b * C + c) * HW;
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    double al = 0.0, ar = 0.0;
                    for (int d = 0; d < D; ++d) {
                        const REAL *gd = g + (((long)b * D + d) * H + h) * W;
                        if (d <= w) al += (double)gd[w] * (double)r[h * W + w - d];
                        if (w + d < W) ar += (double)gd[w + d] * (double)l[h * W + w + d];
                    }
                    gL[((long)b * C + c) * HW + h * W + w] = (REAL)(al / C);
                    gR[((long)b * C + c) * HW + h * W + w] = (REAL)(ar / C);
                }
        }
}

/* ------------------------------------------------------------------ */
/* Soft-argmin.  nets/estimation.py:13-30                               */
/*   p = softmax_d(sign * cost);  disp = sum_d d * p_d                  */
/* ------------------------------------------------------------------ */
void FN(orc_softargmin_fwd)(const REAL *cost, REAL *disp,
                            int B, int D, int H, int W, int similarity)
{
    const long HW = (long)H * W;
    const double sgn = similarity ? 1.0 : -1.0;
#pragma omp parallel for schedule(static)
    for (long bp = 0; bp < (long)B * HW; ++bp) {
        const long b = bp / HW, p = bp % HW;
        const REAL *c = cost + b * D * HW + p;
        double m = -1e300;
        for (int d = 0; d < D; ++d) { double v = sgn * c[d * HW]; if (v > m) m = v; }
        double s = 0.0, ws = 0.0;
        for (int d = 0; d < D; ++d) {
            double e = exp(sgn * c[d * HW] - m);
            s += e; ws += e * d;
        }
        disp[bp] = (REAL)(ws / s);
    }
}

/* d cost = sign * g * p_d * (d - disp)    */
void FN(orc_softargmin_bwd)(const REAL *cost, const REAL *gdisp, REAL *gcost,
                            int B, int D, int H, int W, int similarity)
{
    const long HW = (long)H * W;
    const double sgn = similarity ? 1.0 : -1.0;
#pragma omp parallel for schedule(static)
    for (long bp = 0; bp < (long)B * HW; ++bp) {
        const long b = bp / HW, p = bp % HW;
        const REAL *c = cost + b * D * HW + p;
        REAL *gc = gcost + b * D * HW + p;
        double m = -1e300;
        for (int d = 0; d < D; ++d) { double v = sgn * c[d * HW]; if (v > m) m = v; }
        double s = 0.0, ws = 0.0;
        for (int d = 0; d < D; ++d) {
            double e = exp(sgn * c[d * HW] - m);
            s += e; ws += e * d;
        }
        const double dsp = ws / s, g = gdisp[bp];
        for (int d = 0; d < D; ++d) {
            double pd = exp(sgn * c[d * HW] - m) / s;
            gc[d * HW] = (REAL)(sgn * g * pd * (d - dsp));
        }
    }
}

/* ------------------------------------------------------------------ */
/* Modulated deformable convolution (DCNv2).                            */
/*   sampling + validity rule: deform_conv_cuda_kernel.cu:467-497,      */
/*   :570-633;  GEMM with weight[g].flatten(1): deform_conv_cuda.cpp    */
/*   :539-561;  bias: cpp:565-567.                                      */
/* Layouts (all contiguous NCHW):                                       */
/*   x [B,Cin,H,W]; offset [B, dg*2*kh*kw, Ho, Wo] with channel         */
/*   (g*kh*kw + k)*2 + {0:dh, 1:dw}; mask [B, dg*kh*kw, Ho, Wo];        */
/*   weight [Cout, Cin/groups, kh, kw]; out [B,Cout,Ho,Wo].             */
/* mask == NULL means DCNv1 (mask = 1), cu:190-243.                     */
/* ------------------------------------------------------------------ */
typedef struct {
    int B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, Ho, Wo;
} FN(mdcn_dims);

/* The four corner indices/weights of one sampling point; valid==0 when the
 * point fails the (-1,H)x(-1,W) test of cu:618.  Corner i is dropped (weight
 * kept, value treated as 0) when it lies outside the image, cu:481-492.   */
typedef struct {
    int valid;
    int h0, w0;         /* h_low, w_low */
    REAL lh, lw;        /* fractional parts */
    int ok[4];          /* corner in bounds: (h0,w0) (h0,w1) (h1,w0) (h1,w1) */
} FN(sample_pt);

static inline void FN(make_pt)(FN(sample_pt) *s, REAL h, REAL w, int H, int W)
{
    s->valid = (h > -1 && w > -1 && h < H && w < W);
    if (!s->valid) return;
    s->h0 = (int)floor((double)h);
    s->w0 = (int)floor((double)w);
    s->lh = h - (REAL)s->h0;
    s->lw = w - (REAL)s->w0;
    const int h1 = s->h0 + 1, w1 = s->w0 + 1;
    s->ok[0] = (s->h0 >= 0 && s->w0 >= 0);
    s->ok[1] = (s->h0 >= 0 && w1 <= W - 1);
    s->ok[2] = (h1 <= H - 1 && s->w0 >= 0);
    s->ok[3] = (h1 <= H - 1 && w1 <= W - 1);
}

static inline REAL FN(bilinear)(const FN(sample_pt) *s, const REAL *im, int W)
{
    if (!s->valid) return 0;
    const REAL hh = 1 - s->lh, hw = 1 - s->lw;
    const REAL v1 = s->ok[0] ? im[(long)s->h0 * W + s->w0] : 0;
    const REAL v2 = s->ok[1] ? im[(long)s->h0 * W + s->w0 + 1] : 0;
    const REAL v3 = s->ok[2] ? im[(long)(s->h0 + 1) * W + s->w0] : 0;
    const REAL v4 = s->ok[3] ? im[(long)(s->h0 + 1) * W + s->w0 + 1] : 0;
    return hh * hw * v1 + hh * s->lw * v2 + s->lh * hw * v3 + s->lh * s->lw * v4;
}

static inline void FN(pt_at)(FN(sample_pt) *s, const FN(mdcn_dims) *d,
                             const REAL *off_b, int g, int k, int ho, int wo)
{
    const int K = d->kh * d->kw;
    const long P = (long)d->Ho * d->Wo, p = (long)ho * d->Wo + wo;
    const int i = k / d->kw, j = k % d->kw;
    const REAL oh = off_b[((long)(g * K + k) * 2 + 0) * P + p];
    const REAL ow = off_b[((long)(g * K + k) * 2 + 1) * P + p];
    /* cu:615-616: int + int + scalar_t, evaluated left to right */
    const REAL h = (REAL)(ho * d->stride - d->pad + i * d->dil) + oh;
    const REAL w = (REAL)(wo * d->stride - d->pad + j * d->dil) + ow;
    FN(make_pt)(s, h, w, d->H, d->W);
}

int FN(orc_mdcn_fwd)(const REAL *x, const REAL *offset, const REAL *mask,
                     const REAL *weight, const REAL *bias, REAL *out,
                     int B, int Cin, int H, int W, int Cout, int kh, int kw,
                     int stride, int pad, int dil, int groups, int dg)
{
    FN(mdcn_dims) d = {B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, 0, 0};
    d.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    d.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (Cin % groups || Cout % groups || Cin % dg || d.Ho <= 0 || d.Wo <= 0) return 1;
    const int K = kh * kw, Cg = Cin / groups, Og = Cout / groups, Cd = Cin / dg;
    const long P = (long)d.Ho * d.Wo, HW = (long)H * W;
#pragma omp parallel
    {
        REAL *col = (REAL *)malloc(sizeof(REAL) * (size_t)Cin * K);
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
            for (int ho = 0; ho < d.Ho; ++ho)
                for (int wo = 0; wo < d.Wo; ++wo) {
                    const long p = (long)ho * d.Wo + wo;
                    const REAL *off_b = offset + (long)b * dg * 2 * K * P;
                    for (int g = 0; g < dg; ++g)
                        for (int k = 0; k < K; ++k) {
                            FN(sample_pt) s;
                            FN(pt_at)(&s, &d, off_b, g, k, ho, wo);
                            const REAL m = mask ? mask[((long)b * dg * K + g * K + k) * P + p] : (REAL)1;
                            for (int c = g * Cd; c < (g + 1) * Cd; ++c)
                                col[c * K + k] = FN(bilinear)(&s, x + ((long)b * Cin + c) * HW, W) * m;
                        }
                    for (int o = 0; o < Cout; ++o) {
                        const int gi = o / Og;
                        const REAL *wr = weight + (long)o * Cg * K;
                        const REAL *cr = col + (long)gi * Cg * K;
                        double acc = 0.0;
                        for (int q = 0; q < Cg * K; ++q) acc += (double)wr[q] * (double)cr[q];
                        if (bias) acc += bias[o];
                        out[((long)b * Cout + o) * P + p] = (REAL)acc;
                    }
                }
        free(col);
    }
    return 0;
}

/* Backward.  cpp:571-685 orchestration; kernels cu:635-693 (grad_input),
 * cu:695-767 (grad_offset, grad_mask), cpp:659-671 (grad_weight, grad_bias).
 * Every output is overwritten (the reference accumulates into zeros).
 * gmask may be NULL when mask is NULL (DCNv1).                            */
int FN(orc_mdcn_bwd)(const REAL *x, const REAL *offset, const REAL *mask,
                     const REAL *weight, const REAL *gout,
                     REAL *gx, REAL *goffset, REAL *gmask, REAL *gweight, REAL *gbias,
                     int B, int Cin, int H, int W, int Cout, int kh, int kw,
                     int stride, int pad, int dil, int groups, int dg)
{
    FN(mdcn_dims) d = {B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg, 0, 0};
    d.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    d.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (Cin % groups || Cout % groups || Cin % dg || d.Ho <= 0 || d.Wo <= 0) return 1;
    const int K = kh * kw, Cg = Cin / groups, Og = Cout / groups, Cd = Cin / dg;
    const long P = (long)d.Ho * d.Wo, HW = (long)H * W;

    /* double accumulators for the scattered / reduced outputs */
    double *agx = (double *)calloc((size_t)B * Cin * HW, sizeof(double));
    double *agw = (double *)calloc((size_t)Cout * Cg * K, sizeof(double));
    if (!agx || !agw) { free(agx); free(agw); return 2; }

    /* pass 1: gx, goffset, gmask -- one (b, deformable group) per task */
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < dg; ++g) {
            const REAL *off_b = offset + (long)b * dg * 2 * K * P;
            for (int ho = 0; ho < d.Ho; ++ho)
                for (int wo = 0; wo < d.Wo; ++wo) {
                    const long p = (long)ho * d.Wo + wo;
                    for (int k = 0; k < K; ++k) {
                        FN(sample_pt) s;
                        FN(pt_at)(&s, &d, off_b, g, k, ho, wo);
                        const REAL m = mask ? mask[((long)b * dg * K + g * K + k) * P + p] : (REAL)1;
                        double a_oh = 0.0, a_ow = 0.0, a_m = 0.0;
                        for (int c = g * Cd; c < (g + 1) * Cd; ++c) {
                            /* columns = W^T . grad_out  (cpp:623-626) */
                            const int gi = c / Cg, cl = c % Cg;
                            double cg = 0.0;
                            for (int o = gi * Og; o < (gi + 1) * Og; ++o)
                                cg += (double)weight[((long)o * Cg + cl) * K + k] *
                                      (double)gout[((long)b * Cout + o) * P + p];
                            if (!s.valid) continue;   /* cu:747-750, cu:503-507 */
                            const REAL *im = x + ((long)b * Cin + c) * HW;
                            double *gi_ = agx + ((long)b * Cin + c) * HW;
                            const REAL hh = 1 - s.lh, hw = 1 - s.lw;
                            const long i00 = (long)s.h0 * W + s.w0;
                            const REAL v1 = s.ok[0] ? im[i00] : 0, v2 = s.ok[1] ? im[i00 + 1] : 0;
                            const REAL v3 = s.ok[2] ? im[i00 + W] : 0, v4 = s.ok[3] ? im[i00 + W + 1] : 0;
                            /* grad_mask: cu:753 */
                            a_m += cg * (double)(hh * hw * v1 + hh * s.lw * v2 + s.lh * hw * v3 + s.lh * s.lw * v4);
                            /* grad_offset: coordinate weights cu:526-568, times mask cu:758 */
                            a_oh += cg * (double)m * (double)(-hw * v1 - s.lw * v2 + hw * v3 + s.lw * v4);
                            a_ow += cg * (double)m * (double)(-hh * v1 + hh * v2 - s.lh * v3 + s.lh * v4);
                            /* grad_input: cu:662-691, corner weights cu:499-524 */
                            const double t = cg * (double)m;
                            if (s.ok[0]) gi_[i00] += t * (double)(hh * hw);
                            if (s.ok[1]) gi_[i00 + 1] += t * (double)(hh * s.lw);
                            if (s.ok[2]) gi_[i00 + W] += t * (double)(s.lh * hw);
                            if (s.ok[3]) gi_[i00 + W + 1] += t * (double)(s.lh * s.lw);
                        }
                        goffset[((long)b * dg * 2 * K + (long)(g * K + k) * 2 + 0) * P + p] = (REAL)a_oh;
                        goffset[((long)b * dg * 2 * K + (long)(g * K + k) * 2 + 1) * P + p] = (REAL)a_ow;
                        if (gmask) gmask[((long)b * dg * K + g * K + k) * P + p] = (REAL)a_m;
                    }
                }
        }
    for (long i = 0; i < (long)B * Cin * HW; ++i) gx[i] = (REAL)agx[i];
    free(agx);

    /* pass 2: grad_weight[o,cl,k] = sum_{b,p} gout[b,o,p] * col[b,c,k,p]  (cpp:647-664) */
#pragma omp parallel for schedule(dynamic)
    for (int c = 0; c < Cin; ++c) {
        const int g = c / Cd, gi = c / Cg, cl = c % Cg;
        for (int b = 0; b < B; ++b) {
            const REAL *off_b = offset + (long)b * dg * 2 * K * P;
            const REAL *im = x + ((long)b * Cin + c) * HW;
            for (int ho = 0; ho < d.Ho; ++ho)
                for (int wo = 0; wo < d.Wo; ++wo) {
                    const long p = (long)ho * d.Wo + wo;
                    for (int k = 0; k < K; ++k) {
                        FN(sample_pt) s;
                        FN(pt_at)(&s, &d, off_b, g, k, ho, wo);
                        const REAL m = mask ? mask[((long)b * dg * K + g * K + k) * P + p] : (REAL)1;
                        const double col = (double)(FN(bilinear)(&s, im, W) * m);
                        if (col == 0.0) continue;
                        for (int o = gi * Og; o < (gi + 1) * Og; ++o)
                            agw[((long)o * Cg + cl) * K + k] += col * (double)gout[((long)b * Cout + o) * P + p];
                    }
                }
        }
    }
    for (long i = 0; i < (long)Cout * Cg * K; ++i) gweight[i] = (REAL)agw[i];
    free(agw);

    if (gbias)  /* cpp:665-671 */
        for (int o = 0; o < Cout; ++o) {
            double a = 0.0;
            for (int b = 0; b < B; ++b)
                for (long p = 0; p < P; ++p) a += gout[((long)b * Cout + o) * P + p];
            gbias[o] = (REAL)a;
        }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Cross-scale aggregation fuse.  nets/aggregation.py:387-400           */
/*   out = LeakyReLU_slope( ((t0 + r(t1)) + r(t2)) ... )                 */
/* where r() is F.interpolate(size=(H,W), mode='bilinear',              */
/* align_corners=False) when a term's spatial size differs (:394-396).  */
/* Resize arithmetic follows ATen's upsample_bilinear2d:                */
/*   scale = in/out; src = scale*(dst+0.5)-0.5, clamped at 0;           */
/*   i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1.  */
/* ------------------------------------------------------------------ */
static inline void FN(src_index)(int dst, int in, int out, int *i0, int *i1, REAL *l0, REAL *l1)
{
    const REAL scale = (REAL)in / (REAL)out;
    REAL src = scale * ((REAL)dst + (REAL)0.5) - (REAL)0.5;
    if (src < 0) src = 0;
    *i0 = (int)src;
    if (*i0 > in - 1) *i0 = in - 1;
    *i1 = *i0 + ((*i0 < in - 1) ? 1 : 0);
    *l1 = src - (REAL)*i0;
    *l0 = (REAL)1 - *l1;
}

void FN(orc_csa_fuse_fwd)(const REAL *const *terms, const int *th, const int *tw, int n_terms,
                          REAL *out, int B, int C, int H, int W, REAL slope)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int bc = 0; bc < B * C; ++bc)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                REAL acc = 0;
                for (int t = 0; t < n_terms; ++t) {
                    const REAL *src = terms[t] + (long)bc * th[t] * tw[t];
                    REAL v;
                    if (th[t] == H && tw[t] == W) {
                        v = src[(long)h * W + w];
                    } else {
                        int h0, h1, w0, w1; REAL a0, a1, b0, b1;
                        FN(src_index)(h, th[t], H, &h0, &h1, &a0, &a1);
                        FN(src_index)(w, tw[t], W, &w0, &w1, &b0, &b1);
                        v = a0 * (b0 * src[(long)h0 * tw[t] + w0] + b1 * src[(long)h0 * tw[t] + w1]) +
                            a1 * (b0 * src[(long)h1 * tw[t] + w0] + b1 * src[(long)h1 * tw[t] + w1]);
                    }
                    acc = (t == 0) ? v : acc + v;
                }
                out[((long)bc * H + h) * W + w] = acc > 0 ? acc : acc * slope;
            }
}

/* Adjoint of the above: given `out` (post-activation) and gout, produce one
 * gradient per term (same shape as the term).  LeakyReLU'(pre) is decided on
 * the sign of `out` (slope > 0 keeps the sign), as torch's in-place
 * leaky_relu backward does.                                              */
void FN(orc_csa_fuse_bwd)(const REAL *out, const REAL *gout, REAL *const *gterms,
                          const int *th, const int *tw, int n_terms,
                          int B, int C, int H, int W, REAL slope)
{
    for (int t = 0; t < n_terms; ++t) {
        const long n = (long)B * C * th[t] * tw[t];
        double *acc = (double *)calloc((size_t)n, sizeof(double));
#pragma omp parallel for schedule(static)
        for (int bc = 0; bc < B * C; ++bc) {
            double *a = acc + (long)bc * th[t] * tw[t];
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const long i = ((long)bc * H + h) * W + w;
                    const double gp = (double)gout[i] * (out[i] > 0 ? 1.0 : (double)slope);
                    if (th[t] == H && tw[t] == W) { a[(long)h * W + w] += gp; continue; }
                    int h0, h1, w0, w1; REAL a0, a1, b0, b1;
                    FN(src_index)(h, th[t], H, &h0, &h1, &a0, &a1);
                    FN(src_index)(w, tw[t], W, &w0, &w1, &b0, &b1);
                    a[(long)h0 * tw[t] + w0] += gp * (double)(a0 * b0);
                    a[(long)h0 * tw[t] + w1] += gp * (double)(a0 * b1);
                    a[(long)h1 * tw[t] + w0] += gp * (double)(a1 * b0);
                    a[(long)h1 * tw[t] + w1] += gp * (double)(a1 * b1);
                }
        }
        for (long i = 0; i < n; ++i) gterms[t][i] = (REAL)acc[i];
        free(acc);
    }
}

/* ------------------------------------------------------------------ */
/* Refinement front end.  nets/refinement.py:80-95 (= :144-160) and     */
/* nets/warp.py:41-64:                                                   */
/*   disp  = bilinear(low_disp, size=(H,W), align_corners=False) * (W/w) */
/*           (low_disp itself when W == w, refinement.py:84-88)          */
/*   grid  = (x - disp, y), normalised to [-1,1] as warp.py:12-13 does,  */
/*           2*(g/(size-1)) - 1, and un-normalised again by grid_sample  */
/*           (align_corners=True): ((g+1)/2)*(size-1) -- both in the     */
/*           working precision, so the sampling position carries the     */
/*           same rounding as the reference's                            */
/*   warped = bilinear sample of right at grid, padding_mode='border'    */
/*           (coordinates clipped to [0,size-1]; ATen grid_sampler_2d:   */
/*           weights nw=(xse-x)(yse-y) ..., out-of-range corners skipped) */
/*   concat = cat(warped - left, left)   [B,2C,H,W]                      */
/* ------------------------------------------------------------------ */
void FN(orc_refine_frontend_fwd)(const REAL *low, const REAL *left, const REAL *right, REAL *concat,
                                 REAL *disp, int B, int C, int h, int w, int H, int W)
{
    const REAL scale = (REAL)((double)W / (double)w);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                REAL d;
                if (W == w) {
                    d = low[((long)b * h + y) * w + x];
                } else {
                    const REAL *src = low + (long)b * h * w;
                    int h0, h1, w0, w1; REAL a0, a1, b0, b1;
                    FN(src_index)(y, h, H, &h0, &h1, &a0, &a1);
                    FN(src_index)(x, w, W, &w0, &w1, &b0, &b1);
                    d = a0 * (b0 * src[(long)h0 * w + w0] + b1 * src[(long)h0 * w + w1]) +
                        a1 * (b0 * src[(long)h1 * w + w0] + b1 * src[(long)h1 * w + w1]);
                    d = d * scale;
                }
                disp[((long)b * H + y) * W + x] = d;
                REAL gx = (REAL)x - d, gy = (REAL)y;
                gx = (REAL)2 * (gx / (REAL)(W - 1)) - (REAL)1;
                gy = (REAL)2 * (gy / (REAL)(H - 1)) - (REAL)1;
                REAL ix = ((gx + (REAL)1) / (REAL)2) * (REAL)(W - 1);
                REAL iy = ((gy + (REAL)1) / (REAL)2) * (REAL)(H - 1);
                ix = ix < 0 ? 0 : (ix > (REAL)(W - 1) ? (REAL)(W - 1) : ix);
                iy = iy < 0 ? 0 : (iy > (REAL)(H - 1) ? (REAL)(H - 1) : iy);
                const int xw = (int)floor((double)ix), yn = (int)floor((double)iy);
                const REAL fxw = (REAL)xw, fyn = (REAL)yn;
                const REAL nw = (fxw + 1 - ix) * (fyn + 1 - iy), ne = (ix - fxw) * (fyn + 1 - iy);
                const REAL sw = (fxw + 1 - ix) * (iy - fyn), se = (ix - fxw) * (iy - fyn);
                for (int c = 0; c < C; ++c) {
                    const REAL *img = right + ((long)b * C + c) * H * W;
                    REAL v = 0;
                    if (yn >= 0 && yn < H && xw >= 0 && xw < W) v += img[(long)yn * W + xw] * nw;
                    if (yn >= 0 && yn < H && xw + 1 < W) v += img[(long)yn * W + xw + 1] * ne;
                    if (yn + 1 < H && xw >= 0 && xw < W) v += img[(long)(yn + 1) * W + xw] * sw;
                    if (yn + 1 < H && xw + 1 < W) v += img[(long)(yn + 1) * W + xw + 1] * se;
                    const REAL l = left[(((long)b * C + c) * H + y) * W + x];
                    concat[(((long)b * 2 * C + c) * H + y) * W + x] = v - l;
                    concat[(((long)b * 2 * C + C + c) * H + y) * W + x] = l;
                }
            }
}

#undef FN
#undef CAT
#undef CAT_
