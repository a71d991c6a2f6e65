// Modulated deformable convolution (DCNv2, the ISA operator; reference deform_conv_cuda_kernel.cu:570-633 +
// deform_conv_cuda.cpp:539-561) with the bilinear gather served from a TMA-staged input halo in shared memory.
//
// Round-1 kernel (conv_umma_kernel.cuh, MODE 1/2): every producer lane gathers its four corners with LDG.128 from
// L1/L2; 12 % of those lines miss L1 (compulsory: the tile's footprint is new), nearly every warp-level gather waits
// on at least one L2 round trip, and the producers are latency-bound (ncu r01: 31 % warps active, 40 % issue active,
// tensor pipe 16 %).  Here the raw input patch a tile can reach -- the regular dilated grid plus a margin of kMargin
// pixels for the learned offsets and one for the bilinear corner -- is brought into shared memory by ONE
// cp.async.bulk.tensor per (tile, 32-channel block), a whole tile ahead of its use.  Out-of-image pixels are
// zero-filled by the TMA unit, which IS the operator's zero padding rule (each corner outside the image
// contributes 0; a point with h <= -1 or h >= H has both rows outside): no per-corner validity logic is left in the
// fast path.  Producers then read corners with LDS.128 at immediate offsets from one base address (32-bit address
// math, ~30-cycle latency, no misses), combine, split into tf32 hi + lo and store the K-major SWIZZLE_128B operand
// tile exactly as before.  A sample whose 2 x 2 footprint leaves the staged patch (|offset| > kMargin) falls back
// to the global-memory gather with the reference's validity rules, per row, so correctness does not depend on the
// margin.
//
// Per-K-block geometry is computed ONCE per (pixel, tap, deformable group) by the lane that owns the row and
// published through a small per-warp table in shared memory (two LDS.128 per row instead of eight shuffles).
//
// K order is (32-channel block, tap): a halo slot holds one channel block and is released after its taps, so the
// next block's (or tile's) halo streams in underneath the current one (two slots).  The packed weights keep their
// (tap, channel) order; the loader maps the block index.
//
// Roles (512 threads): warps 0-3 epilogue, 4-11 producers (2 groups x 4 warps, a group fills a whole stage),
// 12 halo TMA, 13 weight loader, 14 MMA issuer, 15 idle.  Two A/B stages, two TMEM accumulators.
// Requirements: stride 1, channels per conv group and per deformable group multiples of 32, channels-last output
// with a multiple of 16 channels per group, no residual (the ISA layer of nets/deform.py:216-236 as the fused
// executor calls it).  Everything else takes the round-1 engine.
#include <stdio.h>
#include <stdlib.h>
#include "conv_engine.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kDM = 128, kDTW = 16, kDTH = 8;          // 16 x 8 output pixels per tile
constexpr int kDProdWarp0 = 4, kDGroups = 2;
constexpr int kDStages = kDGroups;                     // one stage per producer group (see conv_umma_kernel.cuh)
constexpr int kDATile = kDM * 32 * 4;                  // 16 KB (hi); same for lo
constexpr int kDSmemBudget = 216 * 1024;

// R = rows of the operand tile one producer thread fills per K block (8 lanes share a row, one 16-byte chunk each):
// R = 8 -> 2 groups x 4 warps (512 threads), R = 4 -> 2 groups x 8 warps (768 threads): more warps to hide the
// shared-memory latency of the gather at fewer registers each.
template <int R> struct DCfg {
    static constexpr int kGroupWarps = 4 * (8 / R);
    static constexpr int kProdWarps = kDGroups * kGroupWarps;
    static constexpr int kTmaWarp = kDProdWarp0 + kProdWarps, kLoadWarp = kTmaWarp + 1, kMmaWarp = kTmaWarp + 2;
    static constexpr int kThreads = ((kMmaWarp + 1 + 3) / 4) * 4 * 32;
    static constexpr int kRowsPerWarp = 4 * R;
    // per-warp geometry table: bilinear weights (16 B per row, +16 B per row group: conflict-free broadcast reads),
    // halo byte offsets (4 B per row) and the global-fallback indices (16 B per row)
    static constexpr int kWOff = 0, kBaseOff = kRowsPerWarp * 16 + 64, kIdxOff = kBaseOff + kRowsPerWarp * 4;
    static constexpr int kTableBytes = kIdxOff + kRowsPerWarp * 16;
};

struct DeformHaloParams {
    ConvParams p;
    int HH, HWd, lines, slot_bytes;    // halo box (rows, pixels per row), lines = HH * HWd, bytes rounded to 1024
    int margin;                        // pixels of offset the halo covers on every side
    int n_cb;                          // 32-channel blocks per convolution group
    int prof;
};

struct DItem { int grp, nt, b, ty, tx; };

__device__ __forceinline__ DItem d_item(const ConvParams &p, int t) {
    DItem it;
    const int pt = t % p.n_ptiles, gn = t / p.n_ptiles;
    it.grp = gn / p.n_tiles_n; it.nt = gn - it.grp * p.n_tiles_n;
    it.b = pt / p.tiles_per_img;
    const int r = pt - it.b * p.tiles_per_img;
    it.ty = r / p.tiles_x; it.tx = r - it.ty * p.tiles_x;
    return it;
}

template <int BN, int R>
__global__ void __launch_bounds__(DCfg<R>::kThreads, 1)
deform_halo_kernel(const __grid_constant__ DeformHaloParams hp, const __grid_constant__ CUtensorMap tm) {
    using Cfg = EngineCfgLite<BN>;
    using DC = DCfg<R>;
    constexpr int kDProdWarps = DC::kProdWarps, kDTmaWarp = DC::kTmaWarp, kDLoadWarp = DC::kLoadWarp,
                  kDMmaWarp = DC::kMmaWarp;
    constexpr int S = kDStages;
    constexpr int kBTile = 2 * BN * 32 * 4;                   // [B_hi | B_lo] of one K block
    constexpr int kStageBytes = 2 * kDATile + kBTile;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_halo_full[2], bar_halo_empty[2];
    __shared__ __align__(8) uint64_t bar_full_a[S], bar_full_b[S], bar_empty[S];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_aff[2][BN];
    __shared__ __align__(16) uint8_t s_table[kDProdWarps][DC::kTableBytes];

    const ConvParams &p = hp.p;
    const MdcnDims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t *halo0 = smem + (size_t)S * kStageBytes;          // two halo slots behind the A/B ring
    const int T = d.K, n_cb = hp.n_cb, total = p.total_tiles;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&bar_halo_full[s], 1);            // expect_tx + TMA bytes
            umma::mbar_init(&bar_halo_empty[s], kDProdWarps); // every producer warp, after its last tap of the slot
        }
        for (int s = 0; s < S; ++s) {
            umma::mbar_init(&bar_full_a[s], DC::kGroupWarps); // the warps of the filling group
            umma::mbar_init(&bar_full_b[s], 1);
            umma::mbar_init(&bar_empty[s], 1);                // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(&bar_acc_full[a], 1);
            umma::mbar_init(&bar_acc_empty[a], 4);
        }
        umma::fence_mbar_init();
    }
    if (warp == kDMmaWarp) umma::tmem_alloc<Cfg::kTmemCols>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    pdl_wait();
    bool triggered = false;

    if (warp < 4) {
        // ================================ epilogue (channels-last, affine + activation) ===========
        const int q = warp, row = q * 32 + lane;
        uint32_t ti = 0;
        int cur_gn = -1;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
            if (t + (int)gridDim.x >= total) { pdl_trigger(); triggered = true; }
            const DItem it = d_item(p, t);
            const int a = ti & 1;
            int e_oh = it.ty * kDTH + (row >> 4), e_ow = it.tx * kDTW + (row & 15);
            const bool p_ok = e_oh < d.Ho && e_ow < d.Wo;
            e_oh = min(e_oh, d.Ho - 1); e_ow = min(e_ow, d.Wo - 1);
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
            float *dst_px = p.out + ((long)it.b * d.P + (long)e_oh * d.Wo + e_ow) * d.Cout + o_base;
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
#pragma unroll 1
            for (int n0 = 0; n0 < BN; n0 += 16) {
                float acc[16], acc2[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + n0, acc);
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + BN + n0, acc2);
                if (!p_ok || n0 >= n_valid) continue;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(&s_aff[0][n0 + i]);
                    const float4 sh = *reinterpret_cast<const float4 *>(&s_aff[1][n0 + i]);
                    acc[i] = fmaf(acc[i] + acc2[i], sc.x, sh.x); acc[i + 1] = fmaf(acc[i + 1] + acc2[i + 1], sc.y, sh.y);
                    acc[i + 2] = fmaf(acc[i + 2] + acc2[i + 2], sc.z, sh.z); acc[i + 3] = fmaf(acc[i + 3] + acc2[i + 3], sc.w, sh.w);
                }
                if (p.act == ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = fmaxf(acc[i], 0.f);
                } else if (p.act == ACT_LEAKY) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = acc[i] > 0.f ? acc[i] : acc[i] * p.slope;
                }
                float4 *dst = reinterpret_cast<float4 *>(dst_px + n0);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);
        }
    } else if (warp < kDProdWarp0 + kDProdWarps) {
        // ================================ A producers: gather from the staged halo ================
        const int pw = warp - kDProdWarp0;                    // producer warp
        const int grpi = pw / DC::kGroupWarps;                // producer group
        const int tg = (tid - kDProdWarp0 * 32) % (DC::kGroupWarps * 32);    // thread within the group
        const int j = tg & 7, rg = lane >> 3;                 // 16-byte chunk; row group inside the warp
        const int row0 = (tg >> 3) * R;                       // first of this thread's R rows
        uint8_t *table = s_table[pw];
        // lanes j < R own the geometry of row (row0 + j); local row index inside the warp's 4 * R rows
        const bool owner = j < R;
        const int lrow = rg * R + (owner ? j : 0);
        uint8_t *my_w = table + DC::kWOff + lrow * 16 + rg * 16;
        int *my_base = reinterpret_cast<int *>(table + DC::kBaseOff) + lrow;
        uint8_t *my_idx = table + DC::kIdxOff + lrow * 16;
        const uint8_t *grp_w = table + DC::kWOff + rg * R * 16 + rg * 16;            // weights of rows row0 .. row0+R-1
        const int *grp_base = reinterpret_cast<const int *>(table + DC::kBaseOff) + rg * R;
        const uint8_t *grp_idx = table + DC::kIdxOff + rg * R * 16;
        long long c_wait_halo = 0, c_wait_stage = 0, c_total = 0;
        const long long pt0 = clock64();

        // Flat walk over this group's K blocks (every kDGroups-th of the CTA's (tile, channel block, tap) sequence).
        // The offsets / mask of the NEXT K block are requested before the current one is processed: each (tap,
        // deformable group) plane is touched once per tile, so those loads are DRAM misses (~1000+ cycles) and
        // must not sit at the head of a K block.
        struct Cur { int t, cb, tap; uint32_t it, hs; };
        auto advance = [&](Cur &c, int n) {
            c.it += n; c.tap += n;
            while (c.tap >= T) {
                c.tap -= T; ++c.cb; ++c.hs;
                if (c.cb == n_cb) { c.cb = 0; c.t += gridDim.x; }
            }
        };
        struct Pix { int oh, ow; bool ok; long pc; };
        auto my_pixel = [&](int t) {
            const DItem item = d_item(p, t);
            Pix px;
            const int my_row = row0 + (owner ? j : 0);
            px.oh = item.ty * kDTH + (my_row >> 4); px.ow = item.tx * kDTW + (my_row & 15);
            px.ok = px.oh < d.Ho && px.ow < d.Wo;
            px.oh = min(px.oh, d.Ho - 1); px.ow = min(px.ow, d.Wo - 1);
            px.pc = (long)px.oh * d.Wo + px.ow;
            return px;
        };
        auto load_geom = [&](const Cur &c, const Pix &px, float &gh, float &gw, float &gm) {
            const DItem item = d_item(p, c.t);
            const int dgi = (item.grp * d.Cg + c.cb * 32) / d.Cd;
            const long ch = (long)dgi * d.K + c.tap;
            const float *off_b = p.offset + (long)item.b * p.off_bs;
            if (!owner) return;
            gh = __ldg(off_b + px.pc * p.off_ps + (ch * 2) * p.off_cs);
            gw = __ldg(off_b + px.pc * p.off_ps + (ch * 2 + 1) * p.off_cs);
            gm = p.mask ? __ldg(p.mask + (long)item.b * p.mask_bs + px.pc * p.mask_ps + ch * p.mask_cs) : 1.f;
        };

        Cur cur = {(int)blockIdx.x, 0, 0, 0u, 0u};
        advance(cur, grpi);
        Pix pix = {0, 0, false, 0};
        float gh = 0.f, gw = 0.f, gm = 0.f;
        if (cur.t < total) { pix = my_pixel(cur.t); load_geom(cur, pix, gh, gw, gm); }
        uint32_t ready_hs = 0xffffffffu;                      // halo slot this thread has already waited for
        while (cur.t < total) {
            Cur nxt = cur;
            advance(nxt, kDGroups);
            Pix npix = pix;
            float ngh = 0.f, ngw = 0.f, ngm = 0.f;
            if (nxt.t < total) {
                if (nxt.t != cur.t) npix = my_pixel(nxt.t);
                load_geom(nxt, npix, ngh, ngw, ngm);          // in flight while this K block is produced
            }
            const DItem item = d_item(p, cur.t);
            const int s = cur.it % S;
            const uint32_t ph = (cur.it / S) & 1;
            const int hslot = cur.hs & 1;
            const uint32_t halo = umma::smem_u32(halo0 + (size_t)hslot * hp.slot_bytes);
            const int c_abs = item.grp * d.Cg + cur.cb * 32;
            const float *x_b = p.x + (long)item.b * d.HW * d.Cin + c_abs + j * 4;        // global fallback
            float *a_hi = reinterpret_cast<float *>(smem + (size_t)s * kStageBytes);
            float *a_lo = a_hi + kDATile / 4;
            // ---- geometry of my row for this (tap, deformable group): one sample per lane
            {
                const int hy0 = item.ty * kDTH - d.pad - hp.margin, hx0 = item.tx * kDTW - d.pad - hp.margin;
                const int ki = cur.tap / d.kw, kj = cur.tap - ki * d.kw;
                const float py = (float)(pix.oh - d.pad + ki * d.dil) + gh;
                const float px = (float)(pix.ow - d.pad + kj * d.dil) + gw;
                const float fy = floorf(py), fx = floorf(px);
                const float lh = py - fy, lw = px - fx;
                const float m = pix.ok ? gm : 0.f;
                const float ry = fy - (float)hy0, rx = fx - (float)hx0;          // top-left corner inside the halo?
                const bool inside = ry >= 0.f && rx >= 0.f && ry <= (float)(hp.HH - 2) && rx <= (float)(hp.HWd - 2);
                float4 w4 = make_float4((1.f - lh) * (1.f - lw) * m, (1.f - lh) * lw * m, lh * (1.f - lw) * m, lh * lw * m);
                int base = 0;
                int4 i4 = make_int4(0, 0, 0, 0);
                if (inside) {
                    base = ((int)ry * hp.HWd + (int)rx) * 128;                   // byte offset of the top-left line
                } else {
                    // footprint leaves the staged patch: global gather with the reference's validity rules
                    const Sample sm = make_sample(py, px, d.H, d.W);
                    w4 = make_float4(sm.w[0] * m, sm.w[1] * m, sm.w[2] * m, sm.w[3] * m);
                    base = sm.i[0] | (int)0x80000000;
                    i4 = make_int4(sm.i[0], sm.i[1], sm.i[2], sm.i[3]);
                }
                __syncwarp();                                   // previous K block's readers are done
                if (owner) {
                    *reinterpret_cast<float4 *>(my_w) = w4;
                    *my_base = base;
                    if (!inside) *reinterpret_cast<int4 *>(my_idx) = i4;
                }
                __syncwarp();
            }
            if (ready_hs != cur.hs) {
                const long long w0 = clock64();
                umma::mbar_wait(&bar_halo_full[hslot], (cur.hs >> 1) & 1);
                c_wait_halo += clock64() - w0;
                ready_hs = cur.hs;
            }
            {
                const long long w0 = clock64();
                umma::mbar_wait(&bar_empty[s], ph ^ 1);
                c_wait_stage += clock64() - w0;
            }
            // All R halo offsets of my rows at once, so that no corner load waits for a table load; rows are then
            // double-buffered in registers: the corners of row u+1 are requested before row u is combined.
            int bases[R];
#pragma unroll
            for (int u = 0; u < R; u += 4) {
                const int4 b4 = *reinterpret_cast<const int4 *>(grp_base + u);
                bases[u] = b4.x; bases[u + 1] = b4.y; bases[u + 2] = b4.z; bases[u + 3] = b4.w;
            }
            float4 wq[2], q[2][4];
            auto fetch = [&](int u, int buf) {
                wq[buf] = *reinterpret_cast<const float4 *>(grp_w + u * 16);
                if (bases[u] >= 0) {
                    const uint32_t a0 = halo + (uint32_t)bases[u] + j * 16, a1 = a0 + hp.HWd * 128;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[buf][0].x), "=f"(q[buf][0].y), "=f"(q[buf][0].z), "=f"(q[buf][0].w) : "r"(a0));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+128];" : "=f"(q[buf][1].x), "=f"(q[buf][1].y), "=f"(q[buf][1].z), "=f"(q[buf][1].w) : "r"(a0));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[buf][2].x), "=f"(q[buf][2].y), "=f"(q[buf][2].z), "=f"(q[buf][2].w) : "r"(a1));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+128];" : "=f"(q[buf][3].x), "=f"(q[buf][3].y), "=f"(q[buf][3].z), "=f"(q[buf][3].w) : "r"(a1));
                } else {
                    const int4 i4 = *reinterpret_cast<const int4 *>(grp_idx + u * 16);
                    q[buf][0] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)i4.x * d.Cin));
                    q[buf][1] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)i4.y * d.Cin));
                    q[buf][2] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)i4.z * d.Cin));
                    q[buf][3] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)i4.w * d.Cin));
                }
            };
            fetch(0, 0);
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int b = u & 1;
                if (u + 1 < R) fetch(u + 1, b ^ 1);
                const float4 w4 = wq[b];
                float v[4];
                v[0] = w4.x * q[b][0].x + w4.y * q[b][1].x + w4.z * q[b][2].x + w4.w * q[b][3].x;
                v[1] = w4.x * q[b][0].y + w4.y * q[b][1].y + w4.z * q[b][2].y + w4.w * q[b][3].y;
                v[2] = w4.x * q[b][0].z + w4.y * q[b][1].z + w4.z * q[b][2].z + w4.w * q[b][3].z;
                v[3] = w4.x * q[b][0].w + w4.y * q[b][1].w + w4.z * q[b][2].w + w4.w * q[b][3].w;
                float4 h4, l4;
                umma::split_tf32(v[0], h4.x, l4.x); umma::split_tf32(v[1], h4.y, l4.y);
                umma::split_tf32(v[2], h4.z, l4.z); umma::split_tf32(v[3], h4.w, l4.w);
                const int arow = row0 + u;
                const int at = arow * 32 + ((j ^ (arow & 7)) << 2);
                *reinterpret_cast<float4 *>(a_hi + at) = h4;
                *reinterpret_cast<float4 *>(a_lo + at) = l4;
            }
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                umma::mbar_arrive(&bar_full_a[s]);
                // last K block of this group inside the halo slot: this warp has read everything it needs from it
                if (nxt.hs != cur.hs) umma::mbar_arrive(&bar_halo_empty[hslot]);
            }
            cur = nxt; pix = npix; gh = ngh; gw = ngw; gm = ngm;
        }
        c_total = clock64() - pt0;
        if (hp.prof && blockIdx.x == 0 && lane == 0 && pw % DC::kGroupWarps == 0)
            printf("deform halo producer group %d: total %lld cycles, wait halo %lld, wait stage %lld\n", grpi, c_total,
                   c_wait_halo, c_wait_stage);
    } else if (warp == kDTmaWarp) {
        if (lane == 0) {
            // ================================ halo loader (tensor-map TMA) ========================
            uint32_t hs = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const DItem it = d_item(p, t);
                for (int cb = 0; cb < n_cb; ++cb, ++hs) {
                    const int s = hs & 1;
                    umma::mbar_wait_sleep(&bar_halo_empty[s], ((hs >> 1) & 1) ^ 1);
                    umma::mbar_expect_tx(&bar_halo_full[s], hp.lines * 128);
                    umma::tma_load_4d(halo0 + (size_t)s * hp.slot_bytes, &tm, it.grp * d.Cg + cb * 32,
                                      it.tx * kDTW - d.pad - hp.margin, it.ty * kDTH - d.pad - hp.margin, it.b,
                                      &bar_halo_full[s]);
                }
            }
        }
    } else if (warp == kDLoadWarp) {
        if (lane == 0) {
            // ================================ weight loader ========================================
            uint32_t itc = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const DItem it = d_item(p, t);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack) +
                                     (size_t)(it.grp * p.n_tiles_n + it.nt) * p.KB * kBTile;
                for (int cb = 0; cb < n_cb; ++cb)
                    for (int tap = 0; tap < T; ++tap, ++itc) {
                        const int s = itc % S;
                        umma::mbar_wait_sleep(&bar_empty[s], ((itc / S) & 1) ^ 1);
                        umma::mbar_expect_tx(&bar_full_b[s], kBTile);
                        umma::bulk_g2s(smem + (size_t)s * kStageBytes + 2 * kDATile,
                                       src + (size_t)(tap * n_cb + cb) * kBTile, kBTile, &bar_full_b[s]);
                    }
            }
        }
    } else if (warp == kDMmaWarp) {
        if (lane == 0) {
            // ================================ MMA issuer ==========================================
            constexpr uint32_t idesc = umma::make_idesc_tf32(kDM, BN);
            constexpr uint32_t idesc2 = umma::make_idesc_tf32(kDM, 2 * BN);
            uint32_t itc = 0, ti = 0;
            long long c_acc = 0, c_a = 0, c_b = 0, c_issue = 0, t0 = clock64();
            const long long t_start = t0;
#define DPROF(acc_) do { const long long t1 = clock64(); acc_ += t1 - t0; t0 = t1; } while (0)
            const int nkb = n_cb * T;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                DPROF(c_acc);
                const uint32_t d_tmem = tmem_base + a * Cfg::kAccStride;
                for (int kb = 0; kb < nkb; ++kb, ++itc) {
                    const int s = itc % S;
                    const uint32_t ph = (itc / S) & 1;
                    umma::mbar_wait_sleep(&bar_full_a[s], ph);
                    DPROF(c_a);
                    umma::mbar_wait_sleep(&bar_full_b[s], ph);
                    umma::tc_fence_after();
                    DPROF(c_b);
                    const uint32_t a0 = umma::smem_u32(smem + (size_t)s * kStageBytes);
                    const uint64_t a_hi = umma::make_desc_sw128(a0), a_lo = umma::make_desc_sw128(a0 + kDATile);
                    const uint64_t b_hi = umma::make_desc_sw128(a0 + 2 * kDATile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t adv = k * 32;
                        umma::mma_tf32(d_tmem, umma::desc_advance(a_hi, adv), umma::desc_advance(b_hi, adv), idesc2,
                                       (kb | k) != 0);
                        umma::mma_tf32(d_tmem, umma::desc_advance(a_lo, adv), umma::desc_advance(b_hi, adv), idesc, 1);
                    }
                    umma::tc_commit(&bar_empty[s]);
                    DPROF(c_issue);
                }
                umma::tc_commit(&bar_acc_full[a]);
            }
            if (hp.prof && blockIdx.x == 0)
                printf("deform halo MMA thread: %u tiles, total %lld cycles; wait acc %lld, wait A %lld, wait B %lld, issue %lld\n",
                       ti, clock64() - t_start, c_acc, c_a, c_b, c_issue);
        }
    }
    if (!triggered) pdl_trigger();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kDMmaWarp) {
        umma::tc_fence_after();
        umma::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// --------------------------------------------------------------------------------------------- host side
static bool deform_halo_enabled() {
    const char *e = getenv("AANET_DEFORM_HALO");
    return e && e[0] == '1';                   // opt-in until validated
}

template <int BN, int R>
static int deform_halo_launch_inst(const DeformHaloParams &hp, const CUtensorMap &tm, cudaStream_t stream) {
    constexpr size_t stage = 2 * kDATile + 2 * BN * 32 * 4;
    const size_t smem = kDStages * stage + 2 * (size_t)hp.slot_bytes + 1024;
    cudaFuncSetAttribute(deform_halo_kernel<BN, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rounds = ceil_div(hp.p.total_tiles, num_sms());
    const int grid = ceil_div(hp.p.total_tiles, rounds);
    return launch_pdl(deform_halo_kernel<BN, R>, dim3(grid), dim3(DCfg<R>::kThreads), smem, stream, hp, tm);
}

template <int BN>
static int deform_halo_launch_bn(const DeformHaloParams &hp, const CUtensorMap &tm, cudaStream_t stream) {
    const char *er = getenv("AANET_DEFORM_ROWS");          // rows per producer thread: 4 (768 threads) or 8 (512)
    if (er && atoi(er) == 8) return deform_halo_launch_inst<BN, 8>(hp, tm, stream);
    return deform_halo_launch_inst<BN, 4>(hp, tm, stream);
}

// Returns AANET_ERR_UNSUPPORTED when the problem should take the round-1 gather engine instead.
int deform_halo_launch(const ConvParams &src, int BN, cudaStream_t stream) {
    if (!deform_halo_enabled()) return AANET_ERR_UNSUPPORTED;
    const MdcnDims &d = src.d;
    if (d.stride != 1 || d.Cg % 32 || d.Cd % 32 || !aligned16(src.x)) return AANET_ERR_UNSUPPORTED;
    if (src.out_nchw || src.residual || d.Og % 16 || (d.Cout & 3) || src.act == ACT_OFFSET_MASK) return AANET_ERR_UNSUPPORTED;
    if (BN != 32 && BN != 64) return AANET_ERR_UNSUPPORTED;
    if (d.K < kDGroups) return AANET_ERR_UNSUPPORTED;          // every producer group must own a tap in every halo slot
    DeformHaloParams hp;
    hp.p = src;
    const size_t ring = (size_t)kDStages * (2 * kDATile + 2 * BN * 32 * 4);
    const char *em = getenv("AANET_DEFORM_MARGIN");
    int margin = em ? atoi(em) : 3;
    for (; margin >= 0; --margin) {
        hp.HH = kDTH + (d.kh - 1) * d.dil + 2 * margin + 1;
        hp.HWd = kDTW + (d.kw - 1) * d.dil + 2 * margin + 1;
        hp.lines = hp.HH * hp.HWd;
        hp.slot_bytes = (hp.lines * 128 + 1023) & ~1023;
        if (hp.HH <= 256 && hp.HWd <= 256 && ring + 2 * (size_t)hp.slot_bytes <= (size_t)kDSmemBudget) break;
    }
    if (margin < 0) return AANET_ERR_UNSUPPORTED;
    hp.margin = margin;
    hp.n_cb = d.Cg / 32;
    { const char *ep = getenv("AANET_HALO_PROF"); hp.prof = ep && ep[0] == '1'; }
    ConvParams &p = hp.p;
    p.n_tiles_n = ceil_div(d.Og, BN);
    p.K = d.K * d.Cg;
    p.KB = p.K / 32;
    p.tiles_x = ceil_div(d.Wo, kDTW);
    p.tiles_per_img = p.tiles_x * ceil_div(d.Ho, kDTH);
    p.n_ptiles = d.B * p.tiles_per_img;
    const long total = (long)d.groups * p.n_tiles_n * p.n_ptiles;
    if (total > 0x3fffffffL) return AANET_ERR_UNSUPPORTED;
    p.total_tiles = (int)total;
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)d.Cin, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    const uint64_t strides[3] = {(uint64_t)d.Cin * 4, (uint64_t)d.W * d.Cin * 4, (uint64_t)d.HW * d.Cin * 4};
    const uint32_t box[4] = {32, (uint32_t)hp.HWd, (uint32_t)hp.HH, 1};
    const int rc = make_tensor_map_f32(&tm, src.x, 4, dims, strides, box, 0);
    if (rc) return rc;
    return BN == 64 ? deform_halo_launch_bn<64>(hp, tm, stream) : deform_halo_launch_bn<32>(hp, tm, stream);
}

}  // namespace aanet
