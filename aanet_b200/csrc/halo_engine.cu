// Halo-staged tcgen05 convolution kernels (round 2): the operand tiles come from a TMA-loaded input halo.
//
//   halo_conv_kernel   DENSE stride-1 convolutions (1x1, 3x3, dilated, grouped; nets/deform.py:70-72 offset head,
//                      :164-184 / :216-236 bottleneck convolutions, nets/aggregation.py:353-358, :443-450).
//                      Per 8 x 16 output tile and 32-channel block, ONE cp.async.bulk.tensor (4-D tensor map over the
//                      channels-last activation, SWIZZLE_128B, out-of-image pixels zero-filled = the convolution's
//                      zero padding) brings the (16 + (kh-1) dil) x (8 + (kw-1) dil) pixel halo into shared memory as
//                      128-byte lines [pixel][32 channels].  The A operand of tap (ki, kj) is then just a WINDOW into
//                      that buffer: a K-major SWIZZLE_128B descriptor whose start address is shifted by
//                      (ki dil pitch + kj dil) lines and whose stride-byte-offset between 8-row groups is the halo
//                      pitch -- the hardware XORs absolute address bits, so shifted windows read the swizzled lines
//                      correctly (profiles/probes/tma_umma_probe.cu).  No im2col, no producer warps, no per-tap
//                      copies: the only register pass is lo = x - trunc_tf32(x) over the halo (once per halo instead
//                      of once per tap), because kind::tf32 reads the raw fp32 lines as the hi part.
//   Roles (384 threads): warps 0-3 epilogue (TMEM lane quarters), 4-7 hi/lo converters, 8 halo TMA, 9 weight loader
//                      (cp.async.bulk of the pre-packed [B_hi | B_lo] block of a (tap, channel block)), 10 MMA issuer.
//   K order            (channel block, tap): a halo slot is released as soon as its taps are done, so the next
//                      block's / tile's halo streams in underneath the MMAs (2-4 slots).
//   Precision          3xTF32 exactly as the gather engine (conv_umma_kernel.cuh): A_raw x [B_hi | B_lo] + A_lo x B_hi.
#include <stdio.h>
#include <stdlib.h>
#include "conv_engine.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kHM = 128;                   // pixels per tile (UMMA M)
constexpr int kHTW = 8, kHTH = 16;         // output tile: 8 wide (one 8-row descriptor group) x 16 tall
constexpr int kHThreads = 384;
constexpr int kHConvWarp0 = 4, kHTmaWarp = 8, kHLoadWarp = 9, kHMmaWarp = 10;
constexpr int kHMaxSlots = 4, kHMaxBStages = 8;
constexpr int kHSmemBudget = 216 * 1024;

struct HaloParams {
    ConvParams p;          // the problem (tiles_x, tiles_per_img, n_ptiles, n_tiles_n, total_tiles re-derived for 8x16 tiles)
    int HH, HWd, lines;    // halo box (rows, pixels per row, lines = HH * HWd)
    int slot_bytes;        // lines * 128 rounded up to 1024 (raw; the lo copy follows at + slot_bytes)
    int n_cb;              // 32-channel blocks per convolution group
    int n_slots, n_bst;
    int rot;               // per-CTA tap rotation (AANET_HALO_ROT=0 turns it off)
    int prof;              // debug: block 0 prints per-role cycle counters (AANET_HALO_PROF=1)
};

struct HaloItem { int grp, nt, b, ty, tx; };

__device__ __forceinline__ HaloItem halo_item(const ConvParams &p, int t) {
    HaloItem it;
    const int pt = t % p.n_ptiles, gn = t / p.n_ptiles;
    it.grp = gn / p.n_tiles_n; it.nt = gn - it.grp * p.n_tiles_n;
    it.b = pt / p.tiles_per_img;
    const int r = pt - it.b * p.tiles_per_img;
    it.ty = r / p.tiles_x; it.tx = r - it.ty * p.tiles_x;
    return it;
}

template <int BN, bool RES, bool LEAN>
__global__ void __launch_bounds__(kHThreads, 1)
halo_conv_kernel(const __grid_constant__ HaloParams hp, const __grid_constant__ CUtensorMap tm) {
    using Cfg = EngineCfgLite<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_halo_full[kHMaxSlots], bar_halo_lo[kHMaxSlots], bar_halo_empty[kHMaxSlots];
    __shared__ __align__(8) uint64_t bar_b_full[kHMaxBStages], bar_b_empty[kHMaxBStages];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_aff[2][BN];

    const ConvParams &p = hp.p;
    const MdcnDims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);
    const int S = hp.n_slots, SB = hp.n_bst;
    uint8_t *bstage0 = smem + (size_t)S * 2 * hp.slot_bytes;
    constexpr int kBStageBytes = 2 * BN * 32 * 4;
    const int T = d.K, n_cb = hp.n_cb, total = p.total_tiles;
    // Every CTA streams the SAME weight blocks; walking the taps from a per-CTA start spreads the CTAs over the
    // whole packed weight array instead of having all of them ask the same L2 lines at the same time.
    const int rot = hp.rot ? (int)((blockIdx.x * 4u) % (unsigned)T) : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            umma::mbar_init(&bar_halo_full[s], 1);       // expect_tx arrival + TMA bytes
            umma::mbar_init(&bar_halo_lo[s], 4);         // one arrival per converter warp
            umma::mbar_init(&bar_halo_empty[s], 1);      // tcgen05.commit after the slot's last tap
        }
        for (int s = 0; s < SB; ++s) {
            umma::mbar_init(&bar_b_full[s], 1);
            umma::mbar_init(&bar_b_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(&bar_acc_full[a], 1);
            umma::mbar_init(&bar_acc_empty[a], 4);
        }
        umma::fence_mbar_init();
    }
    if (warp == kHMmaWarp) umma::tmem_alloc<Cfg::kTmemCols>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    pdl_wait();                       // everything above reads kernel parameters only
    bool triggered = false;

    if (warp < 4) {
        // ================================ epilogue: TMEM -> registers -> global ==================
        const int q = warp, row = q * 32 + lane;
        uint32_t ti = 0;
        int cur_gn = -1;
        long long e_wait = 0, e_work = 0, et0 = clock64();
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
            if (t + (int)gridDim.x >= total) { pdl_trigger(); triggered = true; }
            const HaloItem it = halo_item(p, t);
            const int a = ti & 1;
            int e_oh = it.ty * kHTH + (row >> 3), e_ow = it.tx * kHTW + (row & 7);
            const bool p_ok = e_oh < d.Ho && e_ow < d.Wo;
            e_oh = min(e_oh, d.Ho - 1); e_ow = min(e_ow, d.Wo - 1);
            const int pix = e_oh * d.Wo + e_ow;
            const int o_base = it.grp * d.Og + it.nt * BN;
            const int n_valid = min(BN, d.Og - it.nt * BN);
            if (it.grp * p.n_tiles_n + it.nt != cur_gn) {
                cur_gn = it.grp * p.n_tiles_n + it.nt;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid < BN) {
                    float sc = 1.f, sh = 0.f;
                    if (tid < n_valid) {
                        const int o = o_base + tid;
                        if (p.scale) { sc = __ldg(p.scale + o); sh = __ldg(p.shift + o); }
                        if (p.bias) sh = fmaf(__ldg(p.bias + o), sc, sh);
                    }
                    s_aff[0][tid] = sc; s_aff[1][tid] = sh;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const long pix_g = (long)it.b * d.P + pix;
            const bool vec_ok = !p.out_nchw && ((d.Cout | o_base) & 3) == 0;
            { const long long t1 = clock64(); e_work += t1 - et0; et0 = t1; }
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
            { const long long t1 = clock64(); e_wait += t1 - et0; et0 = t1; }
#pragma unroll 1
            for (int n0 = 0; n0 < BN; n0 += 16) {
                const bool live = p_ok && n0 < n_valid;
                const bool full = LEAN || (vec_ok && n0 + 16 <= n_valid);
                float res[16];
                if (RES && p.residual && live) {
                    if (full) {
                        const float4 *rp = reinterpret_cast<const float4 *>(p.residual + pix_g * d.Cout + o_base + n0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 r4 = __ldg(rp + i);
                            res[4 * i] = r4.x; res[4 * i + 1] = r4.y; res[4 * i + 2] = r4.z; res[4 * i + 3] = r4.w;
                        }
                    } else if (!LEAN) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int o = o_base + n0 + i;
                            res[i] = (n0 + i < n_valid)
                                         ? (p.out_nchw ? __ldg(p.residual + ((long)it.b * d.Cout + o) * d.P + pix)
                                                       : __ldg(p.residual + pix_g * d.Cout + o))
                                         : 0.f;
                        }
                    }
                }
                float acc[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + n0, acc);
                {
                    float acc2[16];
                    umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + BN + n0, acc2);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += acc2[i];
                }
                if (!live) continue;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(&s_aff[0][n0 + i]);
                    const float4 sh = *reinterpret_cast<const float4 *>(&s_aff[1][n0 + i]);
                    acc[i] = fmaf(acc[i], sc.x, sh.x); acc[i + 1] = fmaf(acc[i + 1], sc.y, sh.y);
                    acc[i + 2] = fmaf(acc[i + 2], sc.z, sh.z); acc[i + 3] = fmaf(acc[i + 3], sc.w, sh.w);
                }
                if (RES && p.residual) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += res[i];
                }
                if (p.act == ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = fmaxf(acc[i], 0.f);
                } else if (p.act == ACT_LEAKY) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = acc[i] > 0.f ? acc[i] : acc[i] * p.slope;
                } else if (!LEAN && p.act == ACT_OFFSET_MASK) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (o_base + n0 + i >= p.n_offset_ch) acc[i] = __fdividef(p.mask_scale, 1.f + __expf(-acc[i]));
                }
                if (full) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out + pix_g * d.Cout + o_base + n0);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                } else if (!LEAN && p.out_nchw) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) p.out[((long)it.b * d.Cout + o_base + n0 + i) * d.P + pix] = acc[i];
                } else if (!LEAN) {
                    float *dst = p.out + pix_g * d.Cout + o_base + n0;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) dst[i] = acc[i];
                }
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);
        }
        if (hp.prof && blockIdx.x == 0 && tid == 0)
            printf("halo epilogue warp 0: wait acc_full %lld, work %lld cycles\n", e_wait, e_work + (clock64() - et0));
    } else if (warp < kHTmaWarp) {
        // ================================ converters: lo = x - trunc_tf32(x) over the halo ========
        const int ct = tid - kHConvWarp0 * 32;            // 0..127
        const int n16 = hp.lines * 8;                     // 16-byte chunks of the halo
        uint32_t hs = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            for (int cb = 0; cb < n_cb; ++cb, ++hs) {
                const int s = hs % S;
                const uint32_t ph = (hs / S) & 1;
                const float4 *raw = reinterpret_cast<const float4 *>(smem + (size_t)s * 2 * hp.slot_bytes);
                float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * 2 * hp.slot_bytes + hp.slot_bytes);
                umma::mbar_wait_sleep(&bar_halo_full[s], ph);
#pragma unroll 4
                for (int i = ct; i < n16; i += 128) {
                    const float4 v = raw[i];
                    float4 l;
                    float h;
                    umma::split_tf32(v.x, h, l.x); umma::split_tf32(v.y, h, l.y);
                    umma::split_tf32(v.z, h, l.z); umma::split_tf32(v.w, h, l.w);
                    lo[i] = l;
                }
                umma::fence_proxy_async();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_halo_lo[s]);
            }
        }
    } else if (warp == kHTmaWarp) {
        if (lane == 0) {
            // ================================ halo loader (tensor-map TMA) ========================
            uint32_t hs = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const HaloItem it = halo_item(p, t);
                for (int cb = 0; cb < n_cb; ++cb, ++hs) {
                    const int s = hs % S;
                    const uint32_t ph = (hs / S) & 1;
                    umma::mbar_wait_sleep(&bar_halo_empty[s], ph ^ 1);
                    umma::mbar_expect_tx(&bar_halo_full[s], hp.lines * 128);
                    umma::tma_load_4d(smem + (size_t)s * 2 * hp.slot_bytes, &tm, it.grp * d.Cg + cb * 32,
                                      it.tx * kHTW - d.pad, it.ty * kHTH - d.pad, it.b, &bar_halo_full[s]);
                }
            }
        }
    } else if (warp == kHLoadWarp) {
        if (lane == 0) {
            // ================================ weight loader (bulk async copy) =====================
            uint32_t itc = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const HaloItem it = halo_item(p, t);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack) +
                                     (size_t)(it.grp * p.n_tiles_n + it.nt) * p.KB * kBStageBytes;
                for (int cb = 0; cb < n_cb; ++cb)
                    for (int tq = 0; tq < T; ++tq, ++itc) {
                        const int tap = tq + rot < T ? tq + rot : tq + rot - T;
                        const int s = itc % SB;
                        const uint32_t ph = (itc / SB) & 1;
                        umma::mbar_wait_sleep(&bar_b_empty[s], ph ^ 1);
                        umma::mbar_expect_tx(&bar_b_full[s], kBStageBytes);
                        // packed K order is (tap, channel block); this kernel walks (channel block, tap)
                        umma::bulk_g2s(bstage0 + (size_t)s * kBStageBytes, src + (size_t)(tap * n_cb + cb) * kBStageBytes,
                                       kBStageBytes, &bar_b_full[s]);
                    }
            }
        }
    } else if (warp == kHMmaWarp) {
        if (lane == 0) {
            // ================================ MMA issuer ==========================================
            constexpr uint32_t idesc = umma::make_idesc_tf32(kHM, BN);
            constexpr uint32_t idesc2 = umma::make_idesc_tf32(kHM, 2 * BN);
            const uint32_t sbo = (uint32_t)hp.HWd * 128;
            uint32_t itc = 0, hs = 0, ti = 0;
            long long c_acc = 0, c_halo = 0, c_b = 0, c_issue = 0, t0 = clock64();
            const long long t_start = t0;
#define HPROF(acc) do { const long long t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                HPROF(c_acc);
                const uint32_t d_tmem = tmem_base + a * Cfg::kAccStride;
                bool first = true;
                for (int cb = 0; cb < n_cb; ++cb, ++hs) {
                    const int s = hs % S;
                    umma::mbar_wait_sleep(&bar_halo_lo[s], (hs / S) & 1);      // raw landed and lo written
                    umma::tc_fence_after();
                    HPROF(c_halo);
                    const uint32_t slot = umma::smem_u32(smem + (size_t)s * 2 * hp.slot_bytes);
                    for (int tq = 0; tq < T; ++tq, ++itc) {
                        const int tap = tq + rot < T ? tq + rot : tq + rot - T;
                        const int sb = itc % SB;
                        umma::mbar_wait_sleep(&bar_b_full[sb], (itc / SB) & 1);
                        umma::tc_fence_after();
                        HPROF(c_b);
                        const int ki = tap / d.kw, kj = tap - ki * d.kw;
                        const uint32_t win = slot + (uint32_t)(ki * d.dil * hp.HWd + kj * d.dil) * 128;
                        const uint64_t a_hi = umma::make_desc_sw128_sbo(win, sbo);
                        const uint64_t a_lo = umma::make_desc_sw128_sbo(win + hp.slot_bytes, sbo);
                        const uint64_t b_hi = umma::make_desc_sw128(umma::smem_u32(bstage0 + (size_t)sb * kBStageBytes));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t adv = k * 32;
                            umma::mma_tf32(d_tmem, umma::desc_advance(a_hi, adv), umma::desc_advance(b_hi, adv), idesc2,
                                           (first && k == 0) ? 0u : 1u);
                            umma::mma_tf32(d_tmem, umma::desc_advance(a_lo, adv), umma::desc_advance(b_hi, adv), idesc, 1);
                        }
                        first = false;
                        umma::tc_commit(&bar_b_empty[sb]);
                        HPROF(c_issue);
                    }
                    umma::tc_commit(&bar_halo_empty[s]);
                }
                umma::tc_commit(&bar_acc_full[a]);
            }
            if (hp.prof && blockIdx.x == 0)
                printf("halo MMA thread: %u tiles, total %lld cycles; wait acc %lld, wait halo %lld, wait B %lld, issue %lld\n",
                       ti, clock64() - t_start, c_acc, c_halo, c_b, c_issue);
        }
    }
    if (!triggered) pdl_trigger();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kHMmaWarp) {
        umma::tc_fence_after();
        umma::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// --------------------------------------------------------------------------------------------- host side
static bool halo_enabled() {
    const char *e = getenv("AANET_HALO");       // read per call: tests and A/B runs flip it at run time
    return e && e[0] == '1';                   // opt-in until validated
}

// Fills the halo geometry; false when the problem is outside what the kernel covers.
static bool halo_plan(const ConvParams &src, int BN, HaloParams &hp) {
    const MdcnDims &d = src.d;
    if (d.stride != 1 || d.Cg % 32 || d.kh > 7 || d.kw > 7) return false;
    if (!aligned16(src.x)) return false;
    hp.p = src;
    hp.HH = kHTH + (d.kh - 1) * d.dil;
    hp.HWd = kHTW + (d.kw - 1) * d.dil;
    if (hp.HH > 256 || hp.HWd > 256) return false;
    hp.lines = hp.HH * hp.HWd;
    hp.slot_bytes = (hp.lines * 128 + 1023) & ~1023;
    hp.n_cb = d.Cg / 32;
    const int bstage = 2 * BN * 32 * 4;
    // shared-memory plan: halo slots first (2 hide a halo load behind the other slot's taps; 3 when one slot is a
    // whole tile), the rest of the budget goes to the weight ring, whose depth hides the L2 latency of the
    // per-K-block weight blocks
    const char *es = getenv("AANET_HALO_SLOTS"), *eb = getenv("AANET_HALO_BST");
    int slots = es ? atoi(es) : (hp.n_cb == 1 ? 3 : 2);
    if (slots < 2) slots = 2;
    if (slots > kHMaxSlots) slots = kHMaxSlots;
    while (slots > 2 && kHSmemBudget - slots * 2 * hp.slot_bytes < 2 * bstage) --slots;
    int bst = (kHSmemBudget - slots * 2 * hp.slot_bytes) / bstage;
    if (bst < 2) return false;
    if (eb && atoi(eb) >= 2 && atoi(eb) < bst) bst = atoi(eb);
    hp.n_bst = bst > kHMaxBStages ? kHMaxBStages : bst;
    hp.n_slots = slots;
    { const char *ep = getenv("AANET_HALO_PROF"); hp.prof = ep && ep[0] == '1'; }
    { const char *er = getenv("AANET_HALO_ROT"); hp.rot = !(er && er[0] == '0'); }
    ConvParams &p = hp.p;
    p.n_tiles_n = ceil_div(d.Og, BN);
    p.K = d.K * d.Cg;
    p.KB = p.K / 32;
    p.tiles_x = ceil_div(d.Wo, kHTW);
    p.tiles_per_img = p.tiles_x * ceil_div(d.Ho, kHTH);
    p.n_ptiles = d.B * p.tiles_per_img;
    const long total = (long)d.groups * p.n_tiles_n * p.n_ptiles;
    if (total > 0x3fffffffL) return false;
    p.total_tiles = (int)total;
    return true;
}

template <int BN, bool RES, bool LEAN>
static int halo_launch_inst(const HaloParams &hp, const CUtensorMap &tm, cudaStream_t stream) {
    const size_t smem = (size_t)hp.n_slots * 2 * hp.slot_bytes + (size_t)hp.n_bst * 2 * BN * 32 * 4 + 1024;
    cudaFuncSetAttribute(halo_conv_kernel<BN, RES, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rounds = ceil_div(hp.p.total_tiles, num_sms());
    const int grid = ceil_div(hp.p.total_tiles, rounds);
    return launch_pdl(halo_conv_kernel<BN, RES, LEAN>, dim3(grid), dim3(kHThreads), smem, stream, hp, tm);
}

template <int BN>
static int halo_launch_bn(const HaloParams &hp, const CUtensorMap &tm, cudaStream_t stream) {
    const ConvParams &p = hp.p;
    const bool res = p.residual != nullptr;
    const bool lean = !p.out_nchw && p.act != ACT_OFFSET_MASK && p.d.Og % 16 == 0 && (p.d.Cout & 3) == 0;
    if (res) return lean ? halo_launch_inst<BN, true, true>(hp, tm, stream) : halo_launch_inst<BN, true, false>(hp, tm, stream);
    return lean ? halo_launch_inst<BN, false, true>(hp, tm, stream) : halo_launch_inst<BN, false, false>(hp, tm, stream);
}

// Returns AANET_ERR_UNSUPPORTED when the problem should take the gather engine instead.
int conv_halo_launch(const ConvParams &src, int BN, cudaStream_t stream) {
    if (!halo_enabled()) return AANET_ERR_UNSUPPORTED;
    HaloParams hp;
    if (!halo_plan(src, BN, hp)) return AANET_ERR_UNSUPPORTED;
    const MdcnDims &d = src.d;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)d.Cin, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    const uint64_t strides[3] = {(uint64_t)d.Cin * 4, (uint64_t)d.W * d.Cin * 4, (uint64_t)d.HW * d.Cin * 4};
    const uint32_t box[4] = {32, (uint32_t)hp.HWd, (uint32_t)hp.HH, 1};
    const int rc = make_tensor_map_f32(&tm, src.x, 4, dims, strides, box, 1);
    if (rc) return rc;
    switch (BN) {
        case 16: return halo_launch_bn<16>(hp, tm, stream);
        case 32: return halo_launch_bn<32>(hp, tm, stream);
        case 48: return halo_launch_bn<48>(hp, tm, stream);
        case 64: return halo_launch_bn<64>(hp, tm, stream);
    }
    return AANET_ERR_UNSUPPORTED;
}

}  // namespace aanet
