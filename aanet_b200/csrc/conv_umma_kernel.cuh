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
// Blackwell mapping.  A CTA is persistent (grid <= #SMs) and walks 128-pixel tiles.  M = 128 pixels is the UMMA M
// (one TMEM lane per pixel), N = BN <= 64 output channels (wider layers = several N tiles), K walks (tap, channel)
// in blocks of 32 tf32 = one 128-byte swizzle row.  Warp roles (20 warps = 5 warpgroups, see kUThreads):
//   warps 0-11  A producers, 3 groups of 4 warps; a group fills a whole stage, so three K blocks are in
//               production at once.  lane = (pixel row, 16-byte chunk): LDG.128 (4 per item for DEFORM, 1
//               for DENSE), bilinear combine, split into tf32 hi + lo (mantissa mask + exact remainder), two
//               STS.128 into SWIZZLE_128B K-major tiles; fence.proxy.async + mbarrier arrive.  DEFORM: each
//               lane owns the bilinear sample of ONE of the 8 rows its row group covers and broadcasts it with
//               __shfl_sync, so the sampling math is done once per (pixel, tap, deformable group) instead of
//               once per 16-byte chunk.  DENSE: a thread's 8 rows are 8 consecutive pixels, one address
//               computation per K block.
//   warp 17     streams the pre-split, pre-swizzled weight block of the stage with one cp.async.bulk.
//   warp 16     one thread issues the 3xTF32 products per K step (hi*hi + hi*lo + lo*hi: fp32-grade accuracy,
//               the parity bar is 1e-4) as tcgen05.mma.kind::tf32 into one of two TMEM accumulators -- two
//               MMAs with the stacked [B_hi|B_lo] operand at BN <= 64, three otherwise -- and releases the
//               stage with tcgen05.commit.
//   warps 12-15 epilogue: tcgen05.ld (lane = pixel), bias / folded-BN affine / residual / activation,
//               128-bit channels-last stores (or coalesced NCHW stores); runs one tile behind the MMA.
//   warps 18-19 pad the control warpgroup (setmaxnreg is warpgroup-collective).
// Launches are chained with programmatic dependent launch (pdl_wait after the prologue, pdl_trigger when a CTA
// starts its last epilogue).  All hand-offs are mbarriers; the 3-stage ring runs across tile boundaries.  3 stages x 48 KB keep the smem
// carve-out at 164 KB, i.e. ~90 KB of L1 for the gathers; multi-tap convolutions use 16 x 8 pixel tiles so
// that a tile's footprint over all taps fits it (1-D 128-pixel tiles + 206 KB of smem gave a 13 % L1 hit rate
// and 579 MB of L2->L1 traffic per 1/3-scale deformable conv).
#pragma once
#include "conv_engine.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kUM = 128;                 // pixels per tile (UMMA M)
constexpr int kUK = 32;                  // K per stage (one 128-byte swizzle row of tf32)
constexpr int kGroups = 3;               // producer groups; each fills one whole stage (3 K blocks in production at once)
constexpr int kProdWarps = 4 * kGroups;  // A producers: kGroups groups of 4 warps
constexpr int kTileW = 16, kTileH = 8;   // 2-D pixel tile (multi-tap convolutions): 128 = 16 x 8 output pixels
constexpr int kMmaWarp = kProdWarps + 4, kLoadWarp = kProdWarps + 5;   // warps kProdWarps..+3: epilogue
// 20 warps = 5 complete warpgroups: 3 producer groups, the epilogue, and {MMA issuer, weight loader, 2 idle warps}.
// Every SM sub-partition hosts 5 warps, so a uniform allocation stops at 96 registers per thread (16384 / 160) and
// the DEFORM producers spilled 80-96 bytes.  setmaxnreg (warpgroup-collective, hence the two padding warps) moves
// registers from the control warpgroup (96 -> 40) to the producers (96 -> 112); the epilogue keeps its 96.
// What the B200 accepted: 3 x 112 + 96 + 40.  120 for the producers, or 72 for the epilogue / 32 for the control
// warps, ended in "unspecified launch failure" (allocation is per sub-partition and, it seems, in units of 16
// registers per thread; code that needs more than its decreased budget faults instead of spilling).
constexpr int kUThreads = (kProdWarps + 8) * 32;                        // 640 threads
constexpr int kRegsProducer = 112, kRegsControl = 40;
constexpr int kATileBytes = kUM * kUK * 4;   // 16 KB (hi); same for lo
constexpr int kSmemBudget = 150 * 1024;   // dynamic (3 stages x <= 48 KB); ~9 KB of static tables on top.  Staying
                                          // under the 164 KB carve-out leaves ~90 KB of L1 for the gathers.
constexpr int kMaxKB = 256;               // K <= 8192

#ifdef AANET_PROFILE
// Profile build only (-DAANET_PROFILE): per-CTA cycle counters of where each role waits.
static __device__ long long g_prof[148 * 16];   // one copy per translation unit (= per MODE)
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(slot) do { const long long prof_t1 = clock64(); prof_acc[slot] += prof_t1 - prof_t0; prof_t0 = prof_t1; } while (0)
#define PROF_DECL() long long prof_acc[16] = {0}
#define PROF_FLUSH(slot) g_prof[blockIdx.x * 16 + (slot)] = prof_acc[slot]
#else
#define PROF_T0()
#define PROF_ADD(slot)
#define PROF_DECL()
#define PROF_FLUSH(slot)
#endif

template <int BN> struct EngineCfg {
    static constexpr int kBTileBytes = BN * kUK * 4;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
    // One stage per producer group: group g always refills stage g, so a group can never run two mbarrier
    // phases ahead of the stage it waits on (with stages != groups the parity wait aliases).
    static constexpr int kStages = kGroups;
    static_assert(BN <= 64 && kStages * kStageBytes <= kSmemBudget, "N tile is capped at 64 (4 stages must fit)");
    // The hi and lo weight tiles (adjacent in smem) are fed as ONE N = 2*BN operand, so a K step is
    // A_hi x [B_hi | B_lo] (columns [0,BN) and [BN,2BN)) + A_lo x B_hi (columns [0,BN)): 2 MMAs and 14 KB of
    // smem operand reads instead of 3 MMAs and 18 KB; the epilogue adds the two column halves.
    static constexpr int kAccCols = 2 * BN;
    static constexpr int kAccStride = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : 128;   // TMEM columns per accumulator
    static constexpr uint32_t kTmemCols = 2 * kAccStride;                     // two accumulators
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;
};

// Swizzled position (in floats) of element (row, k) inside a [rows x 32] SWIZZLE_128B K-major tile.
__host__ __device__ inline int sw128_index(int row, int k) {
    return row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

struct TileCoord { int pi, grp, nt, b, p0; };

// Problem descriptors are staged in shared memory (dynamic indexing of kernel parameters would force a
// local-memory copy; three-way code specialisation blew the register budget).
__device__ __forceinline__ int tile_problem(const ConvParams *pr, int n, int t) {
    int pi = 0;
    if (n > 1 && t >= pr[1].tile_start) pi = 1;
    if (n > 2 && t >= pr[2].tile_start) pi = 2;
    return pi;
}

__device__ __forceinline__ TileCoord tile_coord(const ConvParams *pr, int n, int t) {
    TileCoord c;
    c.pi = tile_problem(pr, n, t);
    const ConvParams &p = pr[c.pi];
    const int lt = t - p.tile_start;
    const int pt = lt % p.n_ptiles, gn = lt / p.n_ptiles;
    c.grp = gn / p.n_tiles_n; c.nt = gn % p.n_tiles_n;
    c.b = pt / p.tiles_per_img;
    c.p0 = pt % p.tiles_per_img;          // tile index inside the image
    return c;
}

// Output pixel of row r (0..127) of tile `ti` of an image: 1-D tiles are 128 consecutive pixels (contiguous
// channels-last input/output: best for 1x1 convolutions), 2-D tiles are 16 x 8 patches whose gather footprint
// over all taps fits L1.  Out-of-range rows are clamped to a valid pixel and flagged.
template <bool POINTWISE = false>
__device__ __forceinline__ void tile_row(const ConvParams &p, int ti, int r, int &oh, int &ow, bool &ok) {
    if (POINTWISE) {              // 1-D tiles only: the image is one row of P pixels, no divisions
        const int px = ti * kUM + r, P32 = (int)p.d.P;
        ok = px < P32;
        oh = 0;
        ow = ok ? px : P32 - 1;
        return;
    }
    if (p.tile2d) {
        const int ty = ti / p.tiles_x, tx = ti - ty * p.tiles_x;
        oh = ty * kTileH + (r >> 4);
        ow = tx * kTileW + (r & 15);
        ok = oh < p.d.Ho && ow < p.d.Wo;
        oh = min(oh, p.d.Ho - 1); ow = min(ow, p.d.Wo - 1);
    } else {
        const int px = ti * kUM + r, P32 = (int)p.d.P;
        ok = px < P32;
        const int pc = ok ? px : P32 - 1;
        oh = pc / p.d.Wo;
        ow = pc - oh * p.d.Wo;
    }
}

// MULTI = false: one problem, its descriptor stays in the constant bank (operands come straight from c[][]);
// MULTI = true: up to kMaxProblems descriptors staged in shared memory and selected per tile.
// RES = false: no problem of the launch has a residual input; the epilogue then does not carry the 16 residual
// registers per 16-channel step (with them in every instantiation all launches were 4-5 % slower: 916 vs 960 pairs/s
// in an experiment that compiled the residual path out).
// LEAN = true: every problem of the launch writes channels-last, has a multiple of 16 output channels per group and
// no offset/mask head, so the epilogue only needs its 128-bit store path (the NCHW / ragged / sigmoid branches are
// compiled out: 923 vs 942 pairs/s in an experiment that removed them everywhere).
// MODE: 0 = DENSE, 1 = DEFORM, 2 = DEFORM where every K block is one (tap, deformable group) run (channels per
// conv group and per deformable group both multiples of 32: the ISA layers of the 1/3 scale) -- the producer then
// carries one bilinear sample per lane instead of two and has no general path; 3 = DENSE with 1-D tiles only
// (1x1 / stride 1 / pad 0 convolutions: a third of the launches), no 2-D tile arithmetic anywhere.
template <int BN, int MODE, bool MULTI, bool RES, bool LEAN>
__global__ void __launch_bounds__(kUThreads, 1)
conv_umma_kernel(const __grid_constant__ ConvBatch B) {
    constexpr bool DEFORM = MODE == 1 || MODE == 2, SINGLE_RUN = MODE == 2, POINTWISE = MODE == 3;
    using Cfg = EngineCfg<BN>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[S], bar_empty[S];   // full: the filling group's warps + the weight block's bytes
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    // (tap, first channel) of every 16-byte chunk of every K block: c | ki << 16 | kj << 20 | tap << 24 | ok << 31
    __shared__ uint32_t s_chunk[kMaxKB * 8];
    // DEFORM: per K block, index of the first chunk whose (tap, deformable group) differs from chunk 0
    // (8 = none) | 16 if the block has at most two such runs (the shuffle-shared geometry path applies)
    __shared__ uint8_t s_kbinfo[kMaxKB];
    // epilogue affine of the current (group, n-tile): out = acc * s_aff[0][n] + s_aff[1][n]
    __shared__ __align__(16) float s_aff[2][BN];
    __shared__ __align__(16) ConvParams s_pr[kMaxProblems];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);   // 1024-byte aligned

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            // one arrival per warp of the filling group + the weight loader's arrive.expect_tx (and its bytes): ONE wait
            // per K block for the MMA warp, whose loop is the CTA's critical path (round 2: every barrier operation
            // costs it ~100 cycles; the tensor pipe drains meanwhile)
            umma::mbar_init(&bar_full[s], kProdWarps / kGroups + 1);
            umma::mbar_init(&bar_empty[s], 1);             // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(&bar_acc_full[a], 1);          // tcgen05.commit after a tile's last MMA
            umma::mbar_init(&bar_acc_empty[a], 4);         // one arrival per epilogue warp
        }
        umma::fence_mbar_init();
    }
    if (warp == kMmaWarp) umma::tmem_alloc<Cfg::kTmemCols>(&s_tmem);
    if (MULTI) {
        static_assert(sizeof(ConvParams) % 4 == 0, "descriptor is copied word-wise");
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&B.pr[0]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_pr[0]);
        for (int i = tid; i < (int)(sizeof(ConvParams) / 4) * B.n; i += kUThreads) dst[i] = src[i];
    }
    const int n_prob = MULTI ? B.n : 1, total_tiles = B.total_tiles;
    // descriptor table the roles index: shared copy (MULTI) or the kernel parameter itself
    const ConvParams *const prob = MULTI ? s_pr : B.pr;
    auto build_tables = [&](const ConvParams &p) {
        const MdcnDims &d = p.d;
        for (int i = tid; i < p.KB * 8; i += kUThreads) {
            const int kk = (i >> 3) * kUK + (i & 7) * 4;
            uint32_t e = 0;
            if (kk < p.K) {
                const int tap = kk / d.Cg, c = kk - tap * d.Cg;
                const int ki = tap / d.kw, kj = tap - ki * d.kw;
                e = (uint32_t)c | ((uint32_t)ki << 16) | ((uint32_t)kj << 20) | ((uint32_t)tap << 24) | 0x80000000u;
            }
            s_chunk[p.tbl_off * 8 + i] = e;
        }
        if (DEFORM) {
            for (int kb = tid; kb < p.KB; kb += kUThreads) {
                int key[8];
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int kk = kb * kUK + c8 * 4;
                    if (kk < p.K) {
                        const int tap = kk / d.Cg, c = kk - tap * d.Cg;
                        key[c8] = tap * 4096 + c / d.Cd;          // (tap, deformable group within the conv group)
                    } else {
                        key[c8] = -1 - c8;                        // padding chunks: weight 0, geometry irrelevant
                    }
                }
                int split = 8;
                for (int c8 = 1; c8 < 8; ++c8)
                    if (key[c8] >= 0 && key[c8] != key[0]) { split = c8; break; }
                bool two = true;
                for (int c8 = split; c8 < 8; ++c8)
                    if (key[c8] >= 0 && key[c8] != key[split]) two = false;
                for (int c8 = 1; c8 < split && c8 < 8; ++c8)
                    if (key[c8] >= 0 && key[c8] != key[0]) two = false;
                s_kbinfo[p.tbl_off + kb] = (uint8_t)(split | (two ? 16 : 0));
            }
        }
    };
    build_tables(B.pr[0]);
    if (B.n > 1) build_tables(B.pr[1]);
    if (B.n > 2) build_tables(B.pr[2]);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    // Everything above (barriers, TMEM, chunk tables) only reads kernel parameters: with programmatic dependent
    // launch it overlaps the tail of the previous kernel of the stream.  Global memory is touched from here on.
    pdl_wait();
    bool triggered = false;
    if (warp >= kMmaWarp) umma::setmaxnreg_dec<kRegsControl>();
    else if (warp < kProdWarps) umma::setmaxnreg_inc<kRegsProducer>();

    if (warp < kProdWarps) {
        // ================================ A producers ===========================================
        // Group g (4 warps) fills every kGroups-th K block of the CTA's (tile, K block) sequence on its own,
        // so kGroups stages are in production concurrently and the wait / fence / arrive chain is paid once per
        // 8 rows per thread.  Inside a group, the 8 lanes t..t+7 of a "row group" cover the eight 16-byte
        // chunks of rows row0..row0+7: lane j handles chunk j of every row and OWNS the geometry of row
        // row0 + j (output coordinates; for DEFORM the bilinear sample), broadcast with __shfl_sync.
        const int grpi = warp >> 2;
        const int tg = tid & 127, j = tg & 7;
        const int row0 = (tg >> 3) * 8;
        const int lane_base = lane & ~7;
        PROF_DECL();
        PROF_T0();

        int t = blockIdx.x, kb = grpi;
        uint32_t it = grpi;
        // (tile, K block) cursor: skip whole tiles while kb runs past the K blocks of the tile's problem
        auto normalize = [&]() {
            while (t < total_tiles) {
                const int nkb = (MULTI ? prob[tile_problem(prob, n_prob, t)].KB : B.pr[0].KB);
                if (kb < nkb) break;
                kb -= nkb; t += gridDim.x;
            }
        };
        normalize();
        int cur_t = -1;
        TileCoord tc = {0, 0, 0, 0, 0};
        int my_oh = 0, my_ow = 0;
        bool my_ok = false;

        int g_oh = 0, g_ow0 = 0;                       // DENSE: first of the 8 consecutive pixels this thread copies
        auto enter_tile = [&](const ConvParams &p) {
            if (DEFORM) {                              // coordinates of the row this lane owns
                tile_row(p, tc.p0, row0 + j, my_oh, my_ow, my_ok);
            } else if (!POINTWISE && p.tile2d) {       // rows row0..row0+7 = 8 consecutive pixels of one tile row
                const int ty = tc.p0 / p.tiles_x, tx = tc.p0 - ty * p.tiles_x;
                g_oh = ty * kTileH + (row0 >> 4);
                g_ow0 = tx * kTileW + (row0 & 15);
            } else {                                   // 1-D tile of a 1x1 / stride 1 / pad 0 conv: pixel index
                g_ow0 = tc.p0 * kUM + row0;
            }
        };
        auto produce = [&](const ConvParams &p) {
            const MdcnDims &d = p.d;
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            float *a_hi = reinterpret_cast<float *>(smem + (size_t)s * Cfg::kStageBytes);
            float *a_lo = a_hi + kATileBytes / 4;
            const uint32_t e = s_chunk[(p.tbl_off + kb) * 8 + j];   // this lane's chunk of the K block
            const bool k_ok = (e >> 31) != 0;
            const int c_abs = tc.grp * d.Cg + (int)(e & 0xffffu);
            const float *x_b = p.x + (long)tc.b * d.HW * d.Cin + c_abs;

            auto store_row = [&](int u, const float (&v)[4]) {
                const int row = row0 + u;
                float4 h4, l4;
                umma::split_tf32(v[0], h4.x, l4.x); umma::split_tf32(v[1], h4.y, l4.y);
                umma::split_tf32(v[2], h4.z, l4.z); umma::split_tf32(v[3], h4.w, l4.w);
                const int at = row * kUK + ((j ^ (row & 7)) << 2);
                *reinterpret_cast<float4 *>(a_hi + at) = h4;
                *reinterpret_cast<float4 *>(a_lo + at) = l4;
            };

            if (!DEFORM) {
                // The 8 rows of this thread are 8 consecutive output pixels of one image row (2-D tiles) or 8
                // consecutive pixels of a 1x1 convolution (1-D tiles), so their input addresses are one base
                // plus a constant step: one address computation per K block instead of eight, no shuffles.
                const int ki = (e >> 16) & 15, kj = (e >> 20) & 15;
                const float *base;
                int step, w0, wlim;
                bool row_ok;
                if (!POINTWISE && p.tile2d) {
                    const int hi_ = g_oh * d.stride - d.pad + ki * d.dil;
                    w0 = g_ow0 * d.stride - d.pad + kj * d.dil;
                    row_ok = k_ok && (unsigned)hi_ < (unsigned)d.H;
                    base = x_b + ((long)hi_ * d.W + w0) * d.Cin;
                    step = d.stride; wlim = d.W;
                } else {
                    w0 = g_ow0; row_ok = k_ok; step = 1; wlim = (int)d.P;
                    base = x_b + (long)g_ow0 * d.Cin;
                }
                const int estep = step * d.Cin;
                float4 q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = row_ok && (unsigned)(w0 + u * step) < (unsigned)wlim;
                    q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) q[u] = __ldg(reinterpret_cast<const float4 *>(base + u * estep));
                }
                PROF_ADD(1);
                umma::mbar_wait_sleep(&bar_empty[s], ph ^ 1);
                PROF_ADD(2);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
                    store_row(u, v);
                }
            } else {
                const uint8_t info = SINGLE_RUN ? (uint8_t)(16 | 8) : s_kbinfo[p.tbl_off + kb];
                const int split = info & 15;
                const float *off_b = p.offset + (long)tc.b * p.off_bs;
                const float *mask_b = p.mask ? p.mask + (long)tc.b * p.mask_bs : nullptr;
                // bilinear sample of (pixel of `pk`, chunk entry `ce`): 4 clamped indices + 4 weights (mask folded in)
                auto geometry = [&](int oh, int ow, bool ok, uint32_t ce, int (&gi)[4], float (&gwt)[4]) {
                    const bool ce_ok = (ce >> 31) != 0;
                    const int tap = (ce >> 24) & 127, ki = (ce >> 16) & 15, kj = (ce >> 20) & 15;
                    const long ch = (long)(((tc.grp * d.Cg + (int)(ce & 0xffffu)) / d.Cd) * d.K + tap);
                    const long pc = (long)oh * d.Wo + ow;
                    const float gh = __ldg(off_b + pc * p.off_ps + (ch * 2) * p.off_cs);
                    const float gw = __ldg(off_b + pc * p.off_ps + (ch * 2 + 1) * p.off_cs);
                    const float gm = mask_b ? __ldg(mask_b + pc * p.mask_ps + ch * p.mask_cs) : 1.f;
                    const Sample sm = make_sample((float)(oh * d.stride - d.pad + ki * d.dil) + gh,
                                                  (float)(ow * d.stride - d.pad + kj * d.dil) + gw, d.H, d.W);
                    const float m = (ok && ce_ok) ? gm : 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) { gi[c4] = sm.i[c4]; gwt[c4] = sm.w[c4] * m; }
                };
                auto gather = [&](const int (&gi)[4], float4 (&q)[4]) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
                        q[c4] = __ldg(reinterpret_cast<const float4 *>(x_b + (long)gi[c4] * d.Cin));
                };
                auto combine = [&](const float (&w4)[4], const float4 (&q)[4], float (&v)[4]) {
                    v[0] = w4[0] * q[0].x + w4[1] * q[1].x + w4[2] * q[2].x + w4[3] * q[3].x;
                    v[1] = w4[0] * q[0].y + w4[1] * q[1].y + w4[2] * q[2].y + w4[3] * q[3].y;
                    v[2] = w4[0] * q[0].z + w4[1] * q[1].z + w4[2] * q[2].z + w4[3] * q[3].z;
                    v[3] = w4[0] * q[0].w + w4[1] * q[1].w + w4[2] * q[2].w + w4[3] * q[3].w;
                };
                constexpr int kPf = 3;             // rows of gathers in flight per lane (4: 891 vs 913 pairs/s)
                int ri[kPf][4];
                float rw[kPf][4];
                float4 q[kPf][4];
                if (info & 16) {
                    // shared path: this lane samples ITS row once per run (at most 2 runs per K block) ...
                    int ia[4], ib[4];
                    float wa[4], wb[4];
                    geometry(my_oh, my_ow, my_ok, s_chunk[(p.tbl_off + kb) * 8], ia, wa);
                    if (split < 8) geometry(my_oh, my_ow, my_ok, s_chunk[(p.tbl_off + kb) * 8 + split], ib, wb);
                    const bool second = j >= split;
                    // ... and every lane receives the sample of row u from its owner (lane_base + u)
                    auto fetch = [&](int u, int (&gi)[4], float (&gwt)[4]) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            gi[c4] = __shfl_sync(0xffffffffu, ia[c4], lane_base + u);
                            gwt[c4] = __shfl_sync(0xffffffffu, wa[c4], lane_base + u);
                        }
                        if (split < 8) {           // block-uniform
#pragma unroll
                            for (int c4 = 0; c4 < 4; ++c4) {
                                const int i2 = __shfl_sync(0xffffffffu, ib[c4], lane_base + u);
                                const float w2 = __shfl_sync(0xffffffffu, wb[c4], lane_base + u);
                                if (second) { gi[c4] = i2; gwt[c4] = w2; }
                            }
                        }
                    };
#pragma unroll
                    for (int u = 0; u < kPf - 1; ++u) { fetch(u, ri[u], rw[u]); gather(ri[u], q[u]); }
                    PROF_ADD(1);
                    umma::mbar_wait_sleep(&bar_empty[s], ph ^ 1);      // the first gathers are already in flight
                    PROF_ADD(2);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int cur = u % kPf, nxt = (u + kPf - 1) % kPf;
                        if (u + kPf - 1 < 8) { fetch(u + kPf - 1, ri[nxt], rw[nxt]); gather(ri[nxt], q[nxt]); }
                        float v[4];
                        combine(rw[cur], q[cur], v);       // K-padding chunks need no zeroing: their weights are 0
                        store_row(u, v);
                    }
                } else {
                    // general path (more than two (tap, group) runs in the block: tiny channel counts): every
                    // lane samples every row for its own chunk
                    const int my_pack = (my_oh << 16) | my_ow | (my_ok ? (int)0x80000000 : 0);
                    PROF_ADD(1);
                    umma::mbar_wait_sleep(&bar_empty[s], ph ^ 1);
                    PROF_ADD(2);
#pragma unroll 1
                    for (int u = 0; u < 8; ++u) {
                        const int pk = __shfl_sync(0xffffffffu, my_pack, lane_base + u);
                        geometry((pk >> 16) & 0x7fff, pk & 0xffff, pk < 0, e, ri[0], rw[0]);
                        gather(ri[0], q[0]);
                        float v[4];
                        combine(rw[0], q[0], v);
                        store_row(u, v);
                    }
                }
            }
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_full[s]);
        };

        while (t < total_tiles) {
            if (t != cur_t) {
                cur_t = t;
                tc = tile_coord(prob, n_prob, t);
                enter_tile(MULTI ? prob[tc.pi] : B.pr[0]);
            }
            produce(MULTI ? prob[tc.pi] : B.pr[0]);
            kb += kGroups; it += kGroups;
            normalize();
        }
        PROF_ADD(1);
        if (tid == 0) { PROF_FLUSH(1); PROF_FLUSH(2); }
    } else if (warp < kMmaWarp) {
        // ================================ epilogue: TMEM -> registers -> global ==================
        const int q = warp & 3;                                  // TMEM lane quarter
        const int row = q * 32 + lane;
        const int et = tid - kProdWarps * 32;                    // 0..127 within the epilogue group
        uint32_t ti = 0;
        int cur_gn = -1;
        PROF_DECL();
        PROF_T0();
        TileCoord tc = {0, 0, 0, 0, 0};
        auto epilogue = [&](const ConvParams &p) {
            const MdcnDims &d = p.d;
            const int a = ti & 1;
            int e_oh, e_ow;
            bool p_ok;
            tile_row<POINTWISE>(p, tc.p0, row, e_oh, e_ow, p_ok);
            const int pix = POINTWISE ? e_ow : e_oh * d.Wo + e_ow;
            const int o_base = tc.grp * d.Og + tc.nt * BN;          // first global out channel of the tile
            const int n_valid = min(BN, d.Og - tc.nt * BN);
            if ((tc.pi << 20) + tc.grp * p.n_tiles_n + tc.nt != cur_gn) {
                // (re)build the per-channel affine of this (problem, group, n-tile):
                //   (acc + bias) * scale + shift  ==  acc * scale + (bias * scale + shift)
                cur_gn = (tc.pi << 20) + tc.grp * p.n_tiles_n + tc.nt;
                asm volatile("bar.sync 1, 128;" ::: "memory");       // previous tile's readers are done
                if (et < BN) {
                    float sc = 1.f, sh = 0.f;
                    if (et < n_valid) {
                        const int o = o_base + et;
                        if (p.scale) { sc = __ldg(p.scale + o); sh = __ldg(p.shift + o); }
                        if (p.bias) sh = fmaf(__ldg(p.bias + o), sc, sh);
                    }
                    s_aff[0][et] = sc; s_aff[1][et] = sh;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const long pix_g = (long)tc.b * d.P + pix;
            const bool vec_ok = !p.out_nchw && ((d.Cout | o_base) & 3) == 0;
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
            PROF_ADD(3);                                   // slot 3: epilogue waiting for an accumulator
            float sm_m = -INFINITY, sm_s = 0.f, sm_w = 0.f;   // ACT_SOFTARGMIN: running max / sum / disparity-weighted sum
#pragma unroll 1
            for (int n0 = 0; n0 < BN; n0 += 16) {
                const bool live = p_ok && n0 < n_valid;
                const bool full = LEAN || (vec_ok && n0 + 16 <= n_valid);
                float res[16];
                if (RES && p.residual && live) {           // issue the residual loads before the TMEM read
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
                                         ? (p.out_nchw ? __ldg(p.residual + ((long)tc.b * d.Cout + o) * d.P + pix)
                                                       : __ldg(p.residual + pix_g * d.Cout + o))
                                         : 0.f;
                        }
                    }
                }
                float acc[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * Cfg::kAccStride + n0, acc);
                {                                          // + A_hi x B_lo half
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
                } else if (!LEAN && p.act == ACT_SOFTARGMIN) {
                    // online softmax over this thread's pixel (same arithmetic as softargmin_fwd_kernel: __expf of the
                    // max-shifted value); channel o = disparity candidate o
                    float cm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) cm = fmaxf(cm, acc[i]);
                    const float nm = fmaxf(sm_m, cm), r = __expf(sm_m - nm);
                    sm_s *= r; sm_w *= r; sm_m = nm;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) {
                            const float e = __expf(acc[i] - nm);
                            sm_s += e; sm_w = fmaf(e, (float)(o_base + n0 + i), sm_w);
                        }
                    continue;
                }
                if (full) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out + pix_g * d.Cout + o_base + n0);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                } else if (!LEAN && p.out_nchw) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) p.out[((long)tc.b * d.Cout + o_base + n0 + i) * d.P + pix] = acc[i];
                } else if (!LEAN) {
                    float *dst = p.out + pix_g * d.Cout + o_base + n0;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + i < n_valid) dst[i] = acc[i];
                }
            }
            if (!LEAN && p.act == ACT_SOFTARGMIN && p_ok) p.out[pix_g] = __fdividef(sm_w, sm_s);
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);
            PROF_ADD(4);                                   // slot 4: epilogue work
        };
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
            if (t + (int)gridDim.x >= total_tiles) {     // last tile: the next kernel's launch + prologue overlap
                pdl_trigger();                     // this CTA's last epilogue
                triggered = true;
            }
            tc = tile_coord(prob, n_prob, t);
            epilogue(MULTI ? prob[tc.pi] : B.pr[0]);
        }
        if (warp == kProdWarps && lane == 0) { PROF_FLUSH(3); PROF_FLUSH(4); }
    } else {
        if (warp == kLoadWarp && lane == 0) {
        // ================================ weight loader (bulk async copy) ========================
            uint32_t it = 0;
            PROF_DECL();
            PROF_T0();
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(prob, n_prob, t);
                const uint8_t *src = nullptr;
                int nkb = 0;
                auto weights_of = [&](const ConvParams &p) {
                    nkb = p.KB;
                    src = reinterpret_cast<const uint8_t *>(p.wpack) +
                          (size_t)(tc.grp * p.n_tiles_n + tc.nt) * p.KB * (2 * Cfg::kBTileBytes);
                };
                weights_of(MULTI ? prob[tc.pi] : B.pr[0]);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    umma::mbar_wait_sleep(&bar_empty[s], ph ^ 1);
                    PROF_ADD(5);                           // slot 5: loader waiting for a free stage
                    umma::mbar_expect_tx(&bar_full[s], 2 * Cfg::kBTileBytes);
                    umma::bulk_g2s(smem + (size_t)s * Cfg::kStageBytes + 2 * kATileBytes,
                                   src + (size_t)kb * 2 * Cfg::kBTileBytes, 2 * Cfg::kBTileBytes, &bar_full[s]);
                    PROF_ADD(6);
                }
            }
            PROF_FLUSH(5); PROF_FLUSH(6);
        } else if (warp == kMmaWarp) {
        // ================================ MMA issuer ==============================================
        // The whole warp walks the loop (uniform control flow) and one elected lane issues the tcgen05 instructions:
        // under `if (lane == 0)` the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY retry loop.  Wrapping stage
        // counter instead of % and /, one barrier wait and one commit per K block.
            constexpr uint32_t idesc = umma::make_idesc_tf32(kUM, BN);
            constexpr uint32_t idesc2 = umma::make_idesc_tf32(kUM, 2 * BN);   // stacked [B_hi | B_lo]
            uint32_t ti = 0, stage = 0, phase = 0;
            const uint32_t smem0 = umma::smem_u32(smem);
            PROF_DECL();
            PROF_T0();
            const long long prof_start = clock64();
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                const int nkb = (MULTI ? prob[tile_problem(prob, n_prob, t)].KB : B.pr[0].KB);
                umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);   // epilogue drained this accumulator
                umma::tc_fence_after();
                PROF_ADD(7);                               // slot 7: MMA waiting for a drained accumulator
                const uint32_t d_tmem = tmem_base + a * Cfg::kAccStride;
#pragma unroll 1
                for (int kb = 0; kb < nkb; ++kb) {
                    umma::mbar_wait_sleep(&bar_full[stage], phase);
                    umma::tc_fence_after();
                    PROF_ADD(8);                           // slot 8: MMA waiting for A and B
                    if (umma::elect_one()) {
                        const uint32_t a0 = smem0 + stage * (uint32_t)Cfg::kStageBytes;
                        const uint64_t a_hi = umma::make_desc_sw128(a0), a_lo = umma::make_desc_sw128(a0 + kATileBytes);
                        const uint64_t b_hi = umma::make_desc_sw128(a0 + 2 * kATileBytes);   // B_lo follows it in smem
                        umma::mma_tf32(d_tmem, a_hi, b_hi, idesc2, kb != 0);
                        umma::mma_tf32(d_tmem, a_lo, b_hi, idesc, 1);
#pragma unroll
                        for (int k = 1; k < kUK / 8; ++k) {
                            const uint32_t adv = k * 32;     // 8 tf32 = 32 bytes along K inside the swizzle row
                            umma::mma_tf32(d_tmem, umma::desc_advance(a_hi, adv), umma::desc_advance(b_hi, adv), idesc2, 1);
                            umma::mma_tf32(d_tmem, umma::desc_advance(a_lo, adv), umma::desc_advance(b_hi, adv), idesc, 1);
                        }
                        umma::tc_commit(&bar_empty[stage]);      // frees this stage when the MMAs above retire
                    }
                    __syncwarp();
                    PROF_ADD(10);                          // slot 10: issuing MMAs
                    if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
                }
                if (umma::elect_one()) umma::tc_commit(&bar_acc_full[a]);       // accumulator of this tile complete
                __syncwarp();
            }
#ifdef AANET_PROFILE
            prof_acc[0] = clock64() - prof_start; prof_acc[11] = ti;
#endif
            if (lane == 0) { PROF_FLUSH(0); PROF_FLUSH(7); PROF_FLUSH(8); PROF_FLUSH(10); PROF_FLUSH(11); }
        }
    }
    if (!triggered) pdl_trigger();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        umma::tc_fence_after();
        umma::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int BN, int MODE, bool MULTI, bool RES, bool LEAN>
static int launch_inst(const ConvBatch &batch, cudaStream_t stream) {
    constexpr size_t smem = EngineCfg<BN>::kSmemBytes;
    cudaFuncSetAttribute(conv_umma_kernel<BN, MODE, MULTI, RES, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    // Persistent grid: the fewest CTAs that still finish in ceil(tiles / #SMs) rounds (416 tiles -> 139 CTAs x 3
    // tiles instead of 148 CTAs of which 28 would idle in the last round): the SMs left free run the coarse-scale
    // kernels that the fused executor issues on parallel streams.
    const int rounds = ceil_div(batch.total_tiles, num_sms());
    const int grid = ceil_div(batch.total_tiles, rounds);
    // Programmatic dependent launch: consecutive engine kernels of a stream overlap launch latency and prologue
    // with the predecessor's last epilogue (see pdl_wait / pdl_trigger in the kernel).
    return launch_pdl(conv_umma_kernel<BN, MODE, MULTI, RES, LEAN>, dim3(grid), dim3(kUThreads), smem, stream, batch);
}

template <int BN, int MODE, bool MULTI, bool RES>
static int launch_lean(const ConvBatch &batch, bool lean, cudaStream_t stream) {
    return lean ? launch_inst<BN, MODE, MULTI, RES, true>(batch, stream)
                : launch_inst<BN, MODE, MULTI, RES, false>(batch, stream);
}

template <int BN, int MODE>
static int launch_one(const ConvBatch &batch, cudaStream_t stream) {
    bool res = false, lean = true;
    for (int i = 0; i < batch.n; ++i) {
        const ConvParams &p = batch.pr[i];
        res |= p.residual != nullptr;
        lean &= !p.out_nchw && p.act != ACT_OFFSET_MASK && p.act != ACT_SOFTARGMIN && p.d.Og % 16 == 0 && (p.d.Cout & 3) == 0;
    }
    if (batch.n > 1)
        return res ? launch_lean<BN, MODE, true, true>(batch, lean, stream)
                   : launch_lean<BN, MODE, true, false>(batch, lean, stream);
    return res ? launch_lean<BN, MODE, false, true>(batch, lean, stream)
               : launch_lean<BN, MODE, false, false>(batch, lean, stream);
}

// One translation unit per MODE (conv_umma_m0.cu .. m3.cu) instantiates the kernels of that mode for every N tile
// width, so the 64 instantiations compile in parallel.
#define AANET_DEFINE_CONV_MODE(MODE_)                                                          \
    int conv_umma_launch_mode##MODE_(const ConvBatch &batch, int BN, cudaStream_t stream) {     \
        switch (BN) {                                                                          \
            case 16: return launch_one<16, MODE_>(batch, stream);                              \
            case 32: return launch_one<32, MODE_>(batch, stream);                              \
            case 48: return launch_one<48, MODE_>(batch, stream);                              \
            case 64: return launch_one<64, MODE_>(batch, stream);                              \
        }                                                                                      \
        return AANET_ERR_UNSUPPORTED;                                                          \
    }

#ifdef AANET_PROFILE
// read-and-clear of this translation unit's counters
#define AANET_DEFINE_PROFILE_READ(MODE_)                                                                          \
    extern "C" __attribute__((visibility("default"))) int aanet_profile_read_m##MODE_(long long *host_dst) {      \
        int rc = (int)cudaMemcpyFromSymbol(host_dst, g_prof, sizeof(long long) * 148 * 16);                       \
        static long long zeros[148 * 16];                                                                         \
        if (!rc) rc = (int)cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros));                                      \
        return rc;                                                                                                \
    }
#else
#define AANET_DEFINE_PROFILE_READ(MODE_)
#endif

}  // namespace aanet
