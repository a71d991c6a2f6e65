// Modulated deformable convolution (DCNv2, the ISA operator; reference deform_conv_cuda_kernel.cu:570-633 +
// deform_conv_cuda.cpp:539-561) -- round-2 kernel: the sampled operand never touches shared memory.
//
//   input halo   ONE cp.async.bulk.tensor (4-D tensor map over the channels-last activation, SWIZZLE_128B) per
//                (tile, 32-channel block) stages the patch a 16 x 8 tile can reach -- the dilated tap grid plus a
//                margin for the learned offsets and the bilinear corner -- in shared memory as 128-byte lines
//                [pixel][32 channels].  Out-of-image pixels are zero-filled by the TMA unit, which IS the operator's
//                padding rule (every corner outside the image contributes 0, cu:467-497), so the fast path has no
//                validity logic.  Two slots: the next block's halo streams in under the current one's taps.
//   gather       thread = output pixel = TMEM lane.  Per K block (tap, 32 channels of one deformable group) a
//                producer thread computes its own bilinear sample (no sharing, no shuffles), reads its four corner
//                lines with LDS.128 (the 128-byte swizzle makes 8 neighbouring pixels hit 8 different bank groups),
//                combines, splits into tf32 hi + lo and writes its row of the A operand straight into TENSOR MEMORY
//                with tcgen05.st (32x32b: one lane per thread).  Samples whose 2 x 2 footprint leaves the staged
//                patch fall back to a global gather with the reference's validity rules, per row.
//   MMA          tcgen05.mma kind::tf32 with A in TMEM, B (pre-packed [B_hi | B_lo] weight block, cp.async.bulk)
//                in shared memory: A_hi x [B_hi | B_lo] (N = 2 BN) + A_lo x B_hi (N = BN) per 8-channel step, i.e. the
//                same 3xTF32 arithmetic as conv_umma_kernel.cuh, but the tensor core no longer re-reads a 32 KB
//                hi/lo tile from shared memory per K block and the producers no longer store one.
// Why: the round-1 gather engine and the shared-memory-halo variant (deform_halo.cu) are both bound by shared-memory
// bandwidth -- per K block ~64 KB of corner reads + 32 KB of operand stores + 56 KB of tensor-core operand reads at
// 128 B/clk (profiles/r02: deform_halo 50 % LSU + tensor reads).  Keeping A in TMEM removes 88 KB of the 180 KB.
//
// Roles (640 threads): warps 0-3 epilogue (TMEM lane quarters), 4-15 producers (3 groups x 4 warps, warp % 4 =
// lane quarter; group g owns A stage g), 16 halo TMA, 17 weight loader, 18 MMA warp (warp-uniform loop, one elected
// lane issues), 19 idle.  setmaxnreg: producers 112, epilogue 96, control warps 40.
// TMEM (512 columns): two accumulators of 2 BN columns, three A stages of 64 columns (32 hi + 32 lo), 64 columns for
// the fused tail's accumulator.
// What the template covers (round 2):
//   DENSE = false  DCNv2, stride 1, deformable groups of >= 32 channels (SUB = 1) or of 16 (SUB = 2: two samples per
//                  K block -- the 1/6 scale); offsets / mask of the next K block arrive through cp.async slots
//   DENSE = true   ordinary convolutions over 32-channel blocks: stride 1 or 2, any tap count including 1x1 (fewer
//                  taps than producer groups: the groups pass every patch slot in order), optional channels-last
//                  residual, the offset / mask head epilogue, the soft-argmin epilogue (non-LEAN)
//   TAIL           + the bottleneck's trailing 1x1 convolution + bn + identity + activation from the activated tile
//   FUSE           the 1x1 convolution's input is the CSA cross-scale sum, produced by the A producers from TMA boxes
//                  of the terms (csa_conv1_tmem_launch)
// Everything else (channel counts that are not multiples of 32, NCHW output of the DCN, multi-problem launches)
// takes the round-1 engine.
#include <stdio.h>
#include <stdlib.h>
#include "conv_engine.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace aanet {

constexpr int kTM = 128, kTTW = 16, kTTH = 8;          // 16 x 8 output pixels per tile
constexpr int kTProdWarp0 = 4;
constexpr int kTAccCol = 0, kTACol = 256, kTAStageCols = 64;      // TMEM column map
// G producer groups of 4 warps, one A stage (64 TMEM columns) and one weight stage each.  The per-K-block chain
// "stage free -> gather -> tcgen05.st -> wait::st -> arrive -> MMA -> commit" is latency-bound, so the number of
// stages in rotation sets the K-block rate: G = 3 (640 threads, 96 registers) or G = 4 (768 threads, 80 registers).
template <int G> struct TCfg {
    static constexpr int kProdWarps = 4 * G;
    static constexpr int kTmaWarp = kTProdWarp0 + kProdWarps, kLoadWarp = kTmaWarp + 1, kMmaWarp = kTmaWarp + 2;
    static constexpr int kThreads = ((kMmaWarp + 1 + 3) / 4) * 4 * 32;
    static_assert(kTACol + G * kTAStageCols <= 512, "TMEM: two accumulators + G operand stages");
};
constexpr int kTSmemBudget = 219 * 1024;   // dynamic; + 1 KB alignment slack + ~6.5 KB static (barriers, tables, geometry slots) <= 227 KB
constexpr int kTAcc3Col = 448;           // TAIL: accumulator of the fused 1x1 convolution (<= 64 columns)

// FUSE: the 1x1 convolution's input does not exist in memory -- it is the cross-scale sum of the CSA stage,
//   x[p, c] = LeakyReLU( sum_k resize_k(term_k)[p, c] )        (nets/aggregation.py:387-400)
// produced by the A producers from TMA-staged tiles / bilinear source patches of the terms and written out once (it
// is the next bottleneck's identity) while it goes into tensor memory as the operand of conv1 (deform.py:164-170).
constexpr int kFuseMaxTerms = 3;
struct FuseGeom {
    int n;                                   // 0: not a fused launch
    int th[kFuseMaxTerms], tw[kFuseMaxTerms];   // term sizes; (H, W): same-size term, else bilinear up-sampling source
    int ph[kFuseMaxTerms], pw[kFuseMaxTerms];   // TMA box of term k in lines: the tile (8 x 16) or the source patch
    int off[kFuseMaxTerms];                  // byte offset of term k's region inside a halo slot (1024-aligned)
    int bytes;                               // sum of the boxes
    float slope;
    float *out;                              // the sum, channels-last [B][H*W][C]
};

struct DeformTmemParams {
    ConvParams p;
    FuseGeom fz;
    int HH, HWd, lines, slot_bytes;    // halo box (rows, pixels per row), lines = HH * HWd, bytes rounded to 1024
    int margin_y, margin_x, n_cb, prof;   // pixels of offset the halo covers above/below and left/right
    int spin;                             // experiment: the MMA thread polls its barriers instead of sleeping on them
    int ns_log2;                          // patch slots: 2 (log2 = 1) or 4 (log2 = 2; the 1x1 layers, whose slots hold ONE K block)
    int sb;                               // weight-ring slots: S (a block is requested when its stage frees) or 2 S
                                          // (requested one use of the stage earlier: hides the ~1500-cycle L2 fetch)
    int rot;                              // CTA b walks the taps of a channel block starting at tap b % T: the CTAs of a
                                          // wave then fetch different weight blocks at any one time (L2 hot spot)
};

struct TItem { int grp, nt, b, ty, tx; };

// AANET_HALO_PROF=2: block 0 records clock64 stamps of the first kTraceKB K blocks (producer: stage free seen / stores
// issued / wait::st done / arrived; MMA thread: A seen / B seen / issued + committed) and prints them at exit.
// Both need a -DAANET_TMEM_PROF build (AANET_NVCC_DEFS): the clock64 reads and counters sit on the critical path of the
// single MMA-issuing thread (measured: ~490 of its ~980 cycles per K block were instrumentation and barrier overhead).
constexpr int kTraceKB = 48;
#ifdef AANET_TMEM_PROF
__device__ long long g_ttrace[kTraceKB][8];
#define TP(...) __VA_ARGS__
#else
#define TP(...)
#endif

__device__ __forceinline__ TItem t_item(const ConvParams &p, int t) {
    TItem it;
    const int pt = t % p.n_ptiles, gn = t / p.n_ptiles;
    it.grp = gn / p.n_tiles_n; it.nt = gn - it.grp * p.n_tiles_n;
    it.b = pt / p.tiles_per_img;
    const int r = pt - it.b * p.tiles_per_img;
    it.ty = r / p.tiles_x; it.tx = r - it.ty * p.tiles_x;
    return it;
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// thread t of the warp writes 8 consecutive 32-bit columns of TMEM lane (base lane + t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(dst_smem)), "l"(src) : "memory");
}

// Packed fp32x2 arithmetic (sm_100: two FMAs per issued instruction; the producers are issue / latency bound).
__device__ __forceinline__ uint64_t pack2(float x, float y) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t p, float &x, float &y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// w0 * q0 + w1 * q1 + w2 * q2 + w3 * q3 for the four channels of a 16-byte chunk, as two packed pairs; evaluated in
// the scalar code's order (mul, then three fused multiply-adds), so the values are bit-identical to it
__device__ __forceinline__ void combine4(const float4 &q0, const float4 &q1, const float4 &q2, const float4 &q3,
                                         uint64_t W0, uint64_t W1, uint64_t W2, uint64_t W3, uint64_t &xy, uint64_t &zw) {
    xy = mul2(W0, pack2(q0.x, q0.y)); zw = mul2(W0, pack2(q0.z, q0.w));
    xy = fma2(W1, pack2(q1.x, q1.y), xy); zw = fma2(W1, pack2(q1.z, q1.w), zw);
    xy = fma2(W2, pack2(q2.x, q2.y), xy); zw = fma2(W2, pack2(q2.z, q2.w), zw);
    xy = fma2(W3, pack2(q3.x, q3.y), xy); zw = fma2(W3, pack2(q3.z, q3.w), zw);
}
// (hi, lo) tf32 split of a packed pair: hi = 13 low mantissa bits cleared, lo = x - hi (exact)
__device__ __forceinline__ void split2(uint64_t v, float &h0, float &h1, float &l0, float &l1) {
    const uint64_t h = v & 0xffffe000ffffe000ull;
    unpack2(h, h0, h1);
    unpack2(sub2(v, h), l0, l1);
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// DENSE = false: DCNv2 (bilinear gather, offsets + mask).  DENSE = true: an ordinary stride-1 convolution through the
// same pipeline -- the "sample" of tap (ki, kj) is the input pixel itself, one line per K block and thread (the
// offset/mask head of nets/deform.py:70-72 and the 3x3 of SimpleBottleneck, nets/deform.py:164-184).
// LEAN = true: channels-last output in whole 16-channel chunks, no offset/mask head epilogue (see conv_umma_kernel.cuh).
// TAIL = true: the bottleneck's trailing 1x1 convolution (conv3 + bn3 + identity + ReLU, nets/deform.py:177-183,
// :229-235) is fused: the epilogue warps activate the main accumulator (bn2 + ReLU), split it into tf32 hi + lo and
// write it back IN PLACE into the accumulator's TMEM columns, where it is the A operand of 3 x (Cm / 8) more MMAs
// against the resident tail weights; a second epilogue pass adds bn3, the residual and the activation and stores.
// MMA order D(0) D(1) C(0) D(2) C(1) ... so the activation pass of tile i runs under the main loop of tile i + 1.
// SUB = deformable groups per 32-channel K block: 1 (>= 32 channels per deformable group) or 2 (16 channels per group,
// the 1/6 scale of the pyramid: a thread then takes two bilinear samples per K block, one per half of its row).
// ATen area_pixel_compute_source_index + guard (float arithmetic), as csa_fuse.cu
__device__ __forceinline__ void fuse_src_index(int dst, int in, float scale, int &i0, int &i1, float &l0, float &l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in - 1);
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

template <int BN, bool DENSE, bool LEAN, int G, bool TAIL = false, int SUB = 1, bool FUSE = false>
__global__ void __launch_bounds__(TCfg<G>::kThreads, 1)
deform_tmem_kernel(const __grid_constant__ DeformTmemParams hp, const __grid_constant__ CUtensorMap tm,
                   const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2) {
    constexpr int S = G, kTGroups = G, kTProdWarps = TCfg<G>::kProdWarps, kTTmaWarp = TCfg<G>::kTmaWarp,
                  kTLoadWarp = TCfg<G>::kLoadWarp, kTMmaWarp = TCfg<G>::kMmaWarp;
    constexpr int kBTile = 2 * BN * 32 * 4;                   // [B_hi | B_lo] of one K block
    static_assert(2 * BN <= 128, "accumulator stride is 128 columns");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_halo_full[4], bar_halo_empty[4];
    const uint32_t ns_log2 = (uint32_t)hp.ns_log2, ns_mask = (1u << ns_log2) - 1u;      // patch slot = hs & mask, phase = hs >> log2
    // Stage barriers come in PAIRS, indexed [stage + S * (use & 1)]: use p of stage s signals full / empty barrier
    // (s, p & 1), whose own phase is p >> 1.  One wait and one commit per K block for the MMA warp as before, but a
    // barrier now completes every second use of its stage, so the weight loader can wait for use p - 2 (and request
    // the block of use p into the spare slot (s, p & 1) of a 2 S-deep ring) without the parity test aliasing.
    __shared__ __align__(8) uint64_t bar_full[2 * S], bar_empty[2 * S];   // full: 4 producer warps + the weight block's bytes
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ __align__(8) uint64_t bar_y2[2], bar_acc3_full, bar_acc3_empty, bar_tailw;      // TAIL only
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_aff[2][BN];
    __shared__ __align__(16) float s_aff3[2][64];
    __shared__ int2 s_tapoff[64];                             // per tap: (ki * dil - pad, kj * dil - pad)
    // offset / mask values of every producer thread's NEXT K block, written by cp.async (no register is live across
    // the K block for them: ptxas spilled the prefetched values right after the load, i.e. waited for the DRAM miss)
    __shared__ float s_geom[3 * SUB][DENSE ? 1 : 128 * G];
    __shared__ int s_dgk[64];                                 // per 32-channel block of the input: first offset channel of its deformable group

    const ConvParams &p = hp.p;
    const MdcnDims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024 - (umma::smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t *halo0 = smem + (size_t)hp.sb * kBTile;           // two halo slots behind the weight ring
    uint8_t *tailw = halo0 + ((size_t)hp.slot_bytes << hp.ns_log2);       // TAIL: resident packed weights of the 1x1 convolution
    if (tid < d.K) s_tapoff[tid] = make_int2((tid / d.kw) * d.dil - d.pad, (tid % d.kw) * d.dil - d.pad);
    if (!DENSE && tid < 64) s_dgk[tid] = ((tid * 32) / max(d.Cd, 1)) * d.K;     // SUB == 2: the block's second group follows at + K
    const int T = d.K, n_cb = hp.n_cb, total = p.total_tiles;
    const int rot = hp.rot ? (int)(blockIdx.x % (unsigned)T) : 0;
    auto phys = [&](int t_) { const int q_ = t_ + rot; return q_ >= T ? q_ - T : q_; };   // walk position -> tap

    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            umma::mbar_init(&bar_halo_full[s], 1);            // expect_tx + TMA bytes
            // every producer warp, after its last tap of the slot -- or, with fewer taps than groups (the 1x1
            // convolutions), when it passes a slot it has no tap in: all groups then see every phase of both slot
            // barriers in order, so a parity wait can never alias (a group that SKIPPED slots would test phase n of a
            // barrier that is still in phase n - 2: found the hard way)
            umma::mbar_init(&bar_halo_empty[s], kTProdWarps);
        }
        for (int s = 0; s < 2 * S; ++s) {
            umma::mbar_init(&bar_full[s], 5);                 // the four warps of the filling group + arrive.expect_tx of
                                                              // the weight loader: ONE wait per K block for the MMA thread
                                                              // (a barrier operation costs that thread ~100 cycles; a
                                                              // separate, deeper weight ring measured slower: 25 -> 30 us)
            umma::mbar_init(&bar_empty[s], 1);                // tcgen05.commit: A stage (TMEM) and B stage (smem) free
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(&bar_acc_full[a], 1);
            umma::mbar_init(&bar_acc_empty[a], 4);
            umma::mbar_init(&bar_y2[a], 4);
        }
        umma::mbar_init(&bar_acc3_full, 1);
        umma::mbar_init(&bar_acc3_empty, 4);
        umma::mbar_init(&bar_tailw, 1);
        umma::fence_mbar_init();
    }
    if (warp == kTMmaWarp) umma::tmem_alloc<512>(&s_tmem);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    // Register rebalancing (warpgroup-collective): 640 threads cap a uniform allocation at 96 registers, with which the
    // deformable producers spill (110 bytes once the instrumentation left the loop: 45.6 -> 53 us).  Same split the
    // round-1 engine runs with: producers 112, epilogue 96, control warps 40.
    // (issued at the top of the role branches: values live across a setmaxnreg must fit the smaller budget)
    pdl_wait();
    bool triggered = false;

    if (warp < 4 && TAIL) {
        // ================================ epilogue with fused 1x1 tail ==========================
        // pass 1 (tile j): main accumulator -> bn2 + act -> tf32 hi / lo, written back in place (columns [0, BN) hi,
        // [BN, 2 BN) lo of the accumulator = the A operand of the tail MMAs); pass 2 (tile j - 1): tail accumulator ->
        // bn3 + residual + act -> global.  groups == 1 and one N tile: the affine tables are constant.
        const int q = warp, row = q * 32 + lane;
        const int Ct = p.tail_cout;
        if (tid < BN) {
            float sc = 1.f, sh = 0.f;
            if (p.scale) { sc = __ldg(p.scale + tid); sh = __ldg(p.shift + tid); }
            if (p.bias) sh = fmaf(__ldg(p.bias + tid), sc, sh);
            s_aff[0][tid] = sc; s_aff[1][tid] = sh;
        }
        if (tid < 64) {
            s_aff3[0][tid] = (tid < Ct && p.tail_scale) ? __ldg(p.tail_scale + tid) : 1.f;
            s_aff3[1][tid] = (tid < Ct && p.tail_shift) ? __ldg(p.tail_shift + tid) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int n_my = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) ++n_my;
        for (int j = 0; j <= n_my; ++j) {
            if (j < n_my) {
                const int a = j & 1;
                umma::mbar_wait_sleep(&bar_acc_full[a], (j >> 1) & 1);
                umma::tc_fence_after();
                const uint32_t acc0 = tmem_base + ((uint32_t)(q * 32) << 16) + kTAccCol + a * 128;
#pragma unroll 1
                for (int n0 = 0; n0 < BN; n0 += 16) {
                    float acc[16], acc2[16];
                    umma::tmem_ld16(acc0 + n0, acc);
                    umma::tmem_ld16(acc0 + BN + n0, acc2);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = fmaf(acc[i] + acc2[i], s_aff[0][n0 + i], s_aff[1][n0 + i]);
                        if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
                        else if (p.act == ACT_LEAKY) v = v > 0.f ? v : v * p.slope;
                        umma::split_tf32(v, acc[i], acc2[i]);
                    }
                    tmem_st16(acc0 + n0, acc);
                    tmem_st16(acc0 + BN + n0, acc2);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_y2[a]);
            }
            if (j >= 1) {
                const int jj = j - 1;
                const int t = blockIdx.x + jj * gridDim.x;
                if (jj == n_my - 1) { pdl_trigger(); triggered = true; }
                const TItem it = t_item(p, t);
                int e_oh = it.ty * kTTH + (row >> 4), e_ow = it.tx * kTTW + (row & 15);
                const bool p_ok = e_oh < d.Ho && e_ow < d.Wo;
                e_oh = min(e_oh, d.Ho - 1); e_ow = min(e_ow, d.Wo - 1);
                const long pix_g = (long)it.b * d.P + (long)e_oh * d.Wo + e_ow;
                umma::mbar_wait_sleep(&bar_acc3_full, jj & 1);
                umma::tc_fence_after();
#pragma unroll 1
                for (int n0 = 0; n0 < 64; n0 += 16) {
                    if (n0 >= Ct) break;                    // uniform
                    float res[16];
                    if (p.tail_residual && p_ok) {
                        const float4 *rp = reinterpret_cast<const float4 *>(p.tail_residual + pix_g * Ct + n0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 r4 = __ldg(rp + i);
                            res[4 * i] = r4.x; res[4 * i + 1] = r4.y; res[4 * i + 2] = r4.z; res[4 * i + 3] = r4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) res[i] = 0.f;
                    }
                    float acc[16];
                    umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kTAcc3Col + n0, acc);
                    if (!p_ok) continue;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = fmaf(acc[i], s_aff3[0][n0 + i], s_aff3[1][n0 + i]) + res[i];
                        if (p.tail_act == ACT_RELU) v = fmaxf(v, 0.f);
                        else if (p.tail_act == ACT_LEAKY) v = v > 0.f ? v : v * p.slope;
                        acc[i] = v;
                    }
                    float4 *dst = reinterpret_cast<float4 *>(p.out + pix_g * Ct + n0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_acc3_empty);
            }
        }
    } else if (warp < 4) {
        // ================================ epilogue (channels-last, affine + activation) ===========
        const int q = warp, row = q * 32 + lane;
        uint32_t ti = 0;
        int cur_gn = -1;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
            if (t + (int)gridDim.x >= total) { pdl_trigger(); triggered = true; }
            const TItem it = t_item(p, t);
            const int a = ti & 1;
            int e_oh = it.ty * kTTH + (row >> 4), e_ow = it.tx * kTTW + (row & 15);
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
            const int pix = e_oh * d.Wo + e_ow;
            const long pix_g = (long)it.b * d.P + pix;
            const bool vec_ok = !p.out_nchw && ((d.Cout | o_base) & 3) == 0;
            umma::mbar_wait_sleep(&bar_acc_full[a], (ti >> 1) & 1);
            umma::tc_fence_after();
            float sm_m = -INFINITY, sm_s = 0.f, sm_w = 0.f;   // ACT_SOFTARGMIN: running max / sum / weighted sum
#pragma unroll 1
            for (int n0 = 0; n0 < BN; n0 += 16) {
                float acc[16], acc2[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kTAccCol + a * 128 + n0, acc);
                umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kTAccCol + a * 128 + BN + n0, acc2);
                if (!p_ok || n0 >= n_valid) continue;
                const bool full = LEAN || (vec_ok && n0 + 16 <= n_valid);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 sc = *reinterpret_cast<const float4 *>(&s_aff[0][n0 + i]);
                    const float4 sh = *reinterpret_cast<const float4 *>(&s_aff[1][n0 + i]);
                    acc[i] = fmaf(acc[i] + acc2[i], sc.x, sh.x); acc[i + 1] = fmaf(acc[i + 1] + acc2[i + 1], sc.y, sh.y);
                    acc[i + 2] = fmaf(acc[i + 2] + acc2[i + 2], sc.z, sh.z); acc[i + 3] = fmaf(acc[i + 3] + acc2[i + 3], sc.w, sh.w);
                }
                if (p.residual) {                   // channels-last, same shape as the output (host-checked)
                    if (full) {
                        const float4 *rp = reinterpret_cast<const float4 *>(p.residual + pix_g * d.Cout + o_base + n0);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 r4 = __ldg(rp + i);
                            acc[4 * i] += r4.x; acc[4 * i + 1] += r4.y; acc[4 * i + 2] += r4.z; acc[4 * i + 3] += r4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + i < n_valid) acc[i] += __ldg(p.residual + pix_g * d.Cout + o_base + n0 + i);
                    }
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
                    // online softmax over my pixel's disparity candidates (= output channels, one N tile), as the
                    // round-1 engine's epilogue and softargmin_fwd_kernel
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
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
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
            if (!LEAN && p.act == ACT_SOFTARGMIN && p_ok) p.out[pix_g] = __fdividef(sm_w, sm_s);
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_acc_empty[a]);
        }
    } else if (warp < kTProdWarp0 + kTProdWarps) {
        // ================================ A producers: thread = pixel = TMEM lane =================
        if (G == 3) umma::setmaxnreg_inc<112>();      // 384 x 112 + 128 x 96 + 128 x 40 <= 640 x 96
        const int pw = warp - kTProdWarp0;
        const int grpi = pw >> 2, q = pw & 3;                 // producer group; TMEM lane quarter (== warp % 4)
        const int row = q * 32 + lane;                        // row of the tile this thread produces
        TP(long long c_wait_halo = 0, c_wait_stage = 0; const long long pt0 = clock64();)

        // Walk of this group's K blocks (every G-th of the CTA's (tile, channel block, tap) sequence).  State is kept
        // small on purpose: per-tile values in `tl`, (cb, tap, hs, ph) for the position, and only the three prefetched
        // geometry values live across the body.  (An earlier cur / nxt pair of full states made ptxas spill the
        // prefetched offsets right after the load, i.e. wait for the DRAM miss it was meant to hide: 45 -> 53 us.)
        struct Tile {
            int t, b, grp, oh, ow, hy0, hx0; bool ok;
            const float *off, *msk;          // offset / mask pointers of my pixel, channel 0
        };
        auto decode_tile = [&](Tile &c) {
            const TItem item = t_item(p, c.t);
            c.b = item.b; c.grp = item.grp;
            c.oh = item.ty * kTTH + (row >> 4); c.ow = item.tx * kTTW + (row & 15);
            c.ok = c.oh < d.Ho && c.ow < d.Wo;
            c.oh = min(c.oh, d.Ho - 1); c.ow = min(c.ow, d.Wo - 1);
            c.hy0 = item.ty * kTTH * d.stride - d.pad - hp.margin_y; c.hx0 = item.tx * kTTW * d.stride - d.pad - hp.margin_x;
            const long pc = (long)c.oh * d.Wo + c.ow;
            c.off = DENSE ? nullptr : p.offset + (long)item.b * p.off_bs + pc * p.off_ps;
            c.msk = (!DENSE && p.mask) ? p.mask + (long)item.b * p.mask_bs + pc * p.mask_ps : nullptr;
        };
        // s_dgk[grp * n_cb + cb] = first offset channel of the deformable group the 32-channel block belongs to
        const int ptid = DENSE ? 0 : pw * 32 + lane;
        auto prefetch_geom = [&](const Tile &c, int cb_, int tap_) {
            if (DENSE) return;
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                const long ch = s_dgk[c.grp * n_cb + cb_] + u * d.K + phys(tap_);
                cp_async4(&s_geom[3 * u + 0][ptid], c.off + (ch * 2) * p.off_cs);
                cp_async4(&s_geom[3 * u + 1][ptid], c.off + (ch * 2 + 1) * p.off_cs);
                if (c.msk) cp_async4(&s_geom[3 * u + 2][ptid], c.msk + ch * p.mask_cs);
            }
        };

        Tile tl;
        tl.t = (int)blockIdx.x;
        int cb = 0, tap = grpi;                               // the group's first K block: tap grpi of block 0 ...
        uint32_t hs = 0u;                                     // halo slot sequence number
        int n_my = 0;                                         // halo slots this CTA walks = its tiles x channel blocks
        for (int t = blockIdx.x; t < total; t += gridDim.x) n_my += n_cb;
        auto pass_slot = [&](uint32_t h) {                    // a slot I have no tap in: see it filled, release it
            if ((int)h >= n_my) return;
            umma::mbar_wait(&bar_halo_full[h & ns_mask], (h >> ns_log2) & 1);
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_halo_empty[h & ns_mask]);
        };
        while (tap >= T) {                                    // ... or, with fewer taps than groups, a later block / tile
            pass_slot(hs);
            tap -= T; ++hs;
            if (++cb == n_cb) { cb = 0; tl.t += (int)gridDim.x; }
        }
        int use = 0;                                          // how often my stage has been filled (see bar_full / bar_empty)
        TP(uint32_t it_seq = grpi;)
        const int s = grpi;                                   // S == G: group g always refills A stage g
        if (tl.t < total) { decode_tile(tl); prefetch_geom(tl, cb, tap); }
        int2 tapo = s_tapoff[phys(tap)];
        uint32_t ready_hs = 0xffffffffu;
        while (tl.t < total) {
            // offsets / mask of my NEXT K block: each (tap, deformable group) plane is touched once per tile, so
            // these loads are DRAM misses and must be in flight while the current K block is produced
            const bool last_in_slot = tap + G >= T;           // my last tap inside this halo slot
            const int hslot = (int)(hs & ns_mask);
            const uint32_t halo = umma::smem_u32(halo0 + (size_t)hslot * hp.slot_bytes);
            if (DENSE) {
                // the tap's input pixel of my output pixel: always inside the staged patch (margin 0); out-of-image
                // pixels were zero-filled by the TMA unit (= the convolution's zero padding)
                if (ready_hs != hs) {
                    umma::mbar_wait(&bar_halo_full[hslot], (hs >> ns_log2) & 1);
                    ready_hs = hs;
                }
                // the line is read into registers BEFORE the stage is awaited: the stage (tensor memory) is then held
                // only for the split + tcgen05.st, and the loads overlap the MMAs that still read it
                const uint32_t a_col = tmem_base + ((uint32_t)(q * 32) << 16) + kTACol + s * kTAStageCols;
                // (stride 2, the CSA down-sampling convs: neighbouring pixels read every other line -> 2-way bank
                // conflicts on 8 loads per K block, irrelevant next to the L2 round trips of the gather engine)
                float4 qd[8];
                if (FUSE) {
                    // my pixel's 32 channels of the cross-scale sum: same-size terms are a line of their tile box,
                    // up-sampled terms four corner lines of their source patch (weights as csa_fuse.cu: products of
                    // the row and column weights, corners in the order 00 01 10 11, terms added in order)
                    const FuseGeom &fz = hp.fz;
#pragma unroll
                    for (int k = 0; k < kFuseMaxTerms; ++k) {
                        if (k >= fz.n) break;
                        const uint32_t reg = halo + (uint32_t)fz.off[k];
                        if (fz.th[k] == d.H && fz.tw[k] == d.W) {
                            const uint32_t P = (reg + (uint32_t)row * 128) | ((uint32_t)(row & 7) << 4);
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 v = lds128(P ^ (uint32_t)(c << 4));
                                if (k == 0) qd[c] = v;
                                else { qd[c].x += v.x; qd[c].y += v.y; qd[c].z += v.z; qd[c].w += v.w; }
                            }
                        } else {
                            const float shh = (float)fz.th[k] / (float)d.H, sww = (float)fz.tw[k] / (float)d.W;
                            int h0, h1, x0, x1, r_lo, c_lo, t_; float a0, a1, b0, b1, u0, u1;
                            fuse_src_index(tl.oh, fz.th[k], shh, h0, h1, a0, a1);
                            fuse_src_index(tl.ow, fz.tw[k], sww, x0, x1, b0, b1);
                            fuse_src_index(tl.hy0, fz.th[k], shh, r_lo, t_, u0, u1);      // patch origin (pad = 0: hy0 = tile row 0)
                            fuse_src_index(tl.hx0, fz.tw[k], sww, c_lo, t_, u0, u1);
                            const int l00 = (h0 - r_lo) * fz.pw[k] + (x0 - c_lo), l01 = (h0 - r_lo) * fz.pw[k] + (x1 - c_lo);
                            const int l10 = (h1 - r_lo) * fz.pw[k] + (x0 - c_lo), l11 = (h1 - r_lo) * fz.pw[k] + (x1 - c_lo);
                            const uint32_t P00 = (reg + (uint32_t)l00 * 128) | ((uint32_t)(l00 & 7) << 4);
                            const uint32_t P01 = (reg + (uint32_t)l01 * 128) | ((uint32_t)(l01 & 7) << 4);
                            const uint32_t P10 = (reg + (uint32_t)l10 * 128) | ((uint32_t)(l10 & 7) << 4);
                            const uint32_t P11 = (reg + (uint32_t)l11 * 128) | ((uint32_t)(l11 & 7) << 4);
                            const float w00 = a0 * b0, w01 = a0 * b1, w10 = a1 * b0, w11 = a1 * b1;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const uint32_t x = (uint32_t)(c << 4);
                                const float4 v00 = lds128(P00 ^ x), v01 = lds128(P01 ^ x), v10 = lds128(P10 ^ x), v11 = lds128(P11 ^ x);
                                float4 v;
                                v.x = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
                                v.y = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
                                v.z = w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
                                v.w = w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
                                if (k == 0) qd[c] = v;
                                else { qd[c].x += v.x; qd[c].y += v.y; qd[c].z += v.z; qd[c].w += v.w; }
                            }
                        }
                    }
                    float4 *dst = reinterpret_cast<float4 *>(fz.out + ((long)tl.b * d.HW + (long)tl.oh * d.W + tl.ow) * d.Cin + cb * 32);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        qd[c].x = qd[c].x > 0.f ? qd[c].x : qd[c].x * fz.slope; qd[c].y = qd[c].y > 0.f ? qd[c].y : qd[c].y * fz.slope;
                        qd[c].z = qd[c].z > 0.f ? qd[c].z : qd[c].z * fz.slope; qd[c].w = qd[c].w > 0.f ? qd[c].w : qd[c].w * fz.slope;
                        if (tl.ok && fz.out) dst[c] = qd[c];
                    }
                } else {
                const int l0 = (tl.oh * d.stride + tapo.x - tl.hy0) * hp.HWd + (tl.ow * d.stride + tapo.y - tl.hx0);
                const uint32_t P0 = (halo + (uint32_t)l0 * 128) | ((uint32_t)(l0 & 7) << 4);
#pragma unroll
                for (int c = 0; c < 8; ++c) qd[c] = lds128(P0 ^ (uint32_t)(c << 4));
                }
                // previous use of my stage retired?  (use - 1 = -1 on the first pass: parity 1 of a fresh barrier passes)
                umma::mbar_wait(&bar_empty[s + S * ((use - 1) & 1)], (uint32_t)(((use - 1) >> 1) & 1));
                TP(const bool tr = hp.prof == 2 && blockIdx.x == 0 && q == 0 && lane == 0 && it_seq < (uint32_t)kTraceKB;
                   if (tr) g_ttrace[it_seq][0] = clock64();)
                umma::tc_fence_after();
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float hi[8], lo[8];
                    split2(pack2(qd[2 * c8].x, qd[2 * c8].y), hi[0], hi[1], lo[0], lo[1]);
                    split2(pack2(qd[2 * c8].z, qd[2 * c8].w), hi[2], hi[3], lo[2], lo[3]);
                    split2(pack2(qd[2 * c8 + 1].x, qd[2 * c8 + 1].y), hi[4], hi[5], lo[4], lo[5]);
                    split2(pack2(qd[2 * c8 + 1].z, qd[2 * c8 + 1].w), hi[6], hi[7], lo[6], lo[7]);
                    tmem_st8(a_col + c8 * 8, hi);
                    tmem_st8(a_col + 32 + c8 * 8, lo);
                }
                TP(if (tr) g_ttrace[it_seq][1] = clock64();)
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                TP(if (tr) g_ttrace[it_seq][2] = clock64();)
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    umma::mbar_arrive(&bar_full[s + S * (use & 1)]);
                    if (last_in_slot) umma::mbar_arrive(&bar_halo_empty[hslot]);
                }
                TP(if (tr) g_ttrace[it_seq][3] = clock64();)
                {   // step to my next K block
                    tap += G; ++use; TP(it_seq += G;)
                    bool moved = false, mine = true;
                    while (tap >= T) {                        // (several wraps when a block has fewer taps than groups)
                        if (!mine) pass_slot(hs);             // the slot just left was mine only on the first wrap
                        mine = false;
                        tap -= T; ++hs;
                        if (++cb == n_cb) { cb = 0; tl.t += (int)gridDim.x; moved = true; }
                    }
                    if (moved && tl.t < total) decode_tile(tl);
                    tapo = s_tapoff[phys(tap)];
                }
                continue;
            }
            // ---- my bilinear sample(s) for this (tap, deformable group [pair])
            asm volatile("cp.async.wait_all;" ::: "memory");
            float py[SUB], px[SUB], lh[SUB], lw[SUB], m[SUB], ry[SUB], rx[SUB];
            bool inside = true;
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                const float gh = s_geom[3 * u][ptid], gw = s_geom[3 * u + 1][ptid], gm = tl.msk ? s_geom[3 * u + 2][ptid] : 1.f;
                py[u] = (float)(tl.oh + tapo.x) + gh;
                px[u] = (float)(tl.ow + tapo.y) + gw;
                const float fy = floorf(py[u]), fx = floorf(px[u]);
                lh[u] = py[u] - fy; lw[u] = px[u] - fx;
                m[u] = tl.ok ? gm : 0.f;
                ry[u] = fy - (float)tl.hy0; rx[u] = fx - (float)tl.hx0;         // top-left corner inside the halo?
                inside = inside && ry[u] >= 0.f && rx[u] >= 0.f && ry[u] <= (float)(hp.HH - 2) && rx[u] <= (float)(hp.HWd - 2);
            }
            {   // the values are in registers (consumed above): request my next K block's into the same slots
                const int ntap = last_in_slot ? tap + G - T : tap + G;
                const int ncb = last_in_slot ? cb + 1 : cb;
                if (ncb < n_cb) {
                    prefetch_geom(tl, ncb, ntap);
                } else if (tl.t + (int)gridDim.x < total) {   // first K block of my next tile
                    Tile tn;
                    tn.t = tl.t + (int)gridDim.x;
                    decode_tile(tn);
                    prefetch_geom(tn, 0, ntap);
                }
            }
            if (ready_hs != hs) {
                TP(const long long t0 = clock64();)
                umma::mbar_wait(&bar_halo_full[hslot], (hs >> ns_log2) & 1);
                TP(c_wait_halo += clock64() - t0;)
                ready_hs = hs;
            }
            const uint32_t a_col = tmem_base + ((uint32_t)(q * 32) << 16) + kTACol + s * kTAStageCols;

            // Fast path: corner line L sits at L * 128 bytes of the slot; its 16-byte chunk c at (c ^ (L & 7)) * 16
            // (SWIZZLE_128B).  With P = line address | ((L & 7) << 4) the chunk address is P ^ (c << 4).
            // Slow path (a footprint leaves the staged patch): global gather with the reference's validity rules.
            // The 32 sampled values are produced into REGISTERS before the A stage is awaited: a group owns one stage,
            // so whatever happens while it holds the stage is serial with the MMAs that consume it.  (First version:
            // gather + combine + store all under the stage, ~1500 cycles per K block and group on top of the MMA
            // turnaround.)  Now the stage is held for the hi / lo split and eight tcgen05.st only.
            const bool all_inside = __all_sync(0xffffffffu, inside);
            uint64_t v[16];                                   // 32 sampled channels as packed pairs
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                constexpr int CH = 8 / SUB;                   // 16-byte chunks of the K row that sample u covers
                const int c0 = u * CH;
                float w0 = (1.f - lh[u]) * (1.f - lw[u]) * m[u], w1 = (1.f - lh[u]) * lw[u] * m[u];
                float w2 = lh[u] * (1.f - lw[u]) * m[u], w3 = lh[u] * lw[u] * m[u];
                if (all_inside) {
                    // the whole warp reads from the staged patch (the common case): no branches in the loop; the four
                    // corner chunks of chunk c+1 are requested before chunk c is combined
                    const int l0 = (int)ry[u] * hp.HWd + (int)rx[u], l2 = l0 + hp.HWd;
                    const uint32_t p0 = halo + (uint32_t)l0 * 128, p2 = halo + (uint32_t)l2 * 128;
                    const uint32_t P0 = p0 | ((uint32_t)(l0 & 7) << 4), P1 = (p0 + 128) | ((uint32_t)((l0 + 1) & 7) << 4);
                    const uint32_t P2 = p2 | ((uint32_t)(l2 & 7) << 4), P3 = (p2 + 128) | ((uint32_t)((l2 + 1) & 7) << 4);
                    const uint64_t W0 = pack2(w0, w0), W1 = pack2(w1, w1), W2 = pack2(w2, w2), W3 = pack2(w3, w3);
                    float4 qb[2][4];
                    {
                        const uint32_t x = (uint32_t)(c0 << 4);
                        qb[0][0] = lds128(P0 ^ x); qb[0][1] = lds128(P1 ^ x); qb[0][2] = lds128(P2 ^ x); qb[0][3] = lds128(P3 ^ x);
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int b = c & 1;
                        if (c + 1 < CH) {
                            const uint32_t x = (uint32_t)((c0 + c + 1) << 4);
                            qb[b ^ 1][0] = lds128(P0 ^ x); qb[b ^ 1][1] = lds128(P1 ^ x);
                            qb[b ^ 1][2] = lds128(P2 ^ x); qb[b ^ 1][3] = lds128(P3 ^ x);
                        }
                        combine4(qb[b][0], qb[b][1], qb[b][2], qb[b][3], W0, W1, W2, W3, v[2 * (c0 + c)], v[2 * (c0 + c) + 1]);
                    }
                } else {
                    // per-lane: from the patch where my footprints are inside it, else from global memory
                    uint32_t P0 = 0, P1 = 0, P2 = 0, P3 = 0;
                    const float4 *g0 = nullptr, *g1 = nullptr, *g2 = nullptr, *g3 = nullptr;
                    if (inside) {
                        const int l0 = (int)ry[u] * hp.HWd + (int)rx[u], l2 = l0 + hp.HWd;
                        const uint32_t p0 = halo + (uint32_t)l0 * 128, p2 = halo + (uint32_t)l2 * 128;
                        P0 = p0 | ((uint32_t)(l0 & 7) << 4); P1 = (p0 + 128) | ((uint32_t)((l0 + 1) & 7) << 4);
                        P2 = p2 | ((uint32_t)(l2 & 7) << 4); P3 = (p2 + 128) | ((uint32_t)((l2 + 1) & 7) << 4);
                    } else {
                        const Sample sm = make_sample(py[u], px[u], d.H, d.W);
                        w0 = sm.w[0] * m[u]; w1 = sm.w[1] * m[u]; w2 = sm.w[2] * m[u]; w3 = sm.w[3] * m[u];
                        const float *x_b = p.x + (long)tl.b * d.HW * d.Cin + tl.grp * d.Cg + cb * 32;
                        g0 = reinterpret_cast<const float4 *>(x_b + (long)sm.i[0] * d.Cin);
                        g1 = reinterpret_cast<const float4 *>(x_b + (long)sm.i[1] * d.Cin);
                        g2 = reinterpret_cast<const float4 *>(x_b + (long)sm.i[2] * d.Cin);
                        g3 = reinterpret_cast<const float4 *>(x_b + (long)sm.i[3] * d.Cin);
                    }
                    const uint64_t W0 = pack2(w0, w0), W1 = pack2(w1, w1), W2 = pack2(w2, w2), W3 = pack2(w3, w3);
#pragma unroll
                    for (int c = c0; c < c0 + CH; ++c) {
                        float4 q0, q1, q2, q3;
                        if (inside) {
                            const uint32_t x = (uint32_t)(c << 4);
                            q0 = lds128(P0 ^ x); q1 = lds128(P1 ^ x); q2 = lds128(P2 ^ x); q3 = lds128(P3 ^ x);
                        } else {
                            q0 = __ldg(g0 + c); q1 = __ldg(g1 + c); q2 = __ldg(g2 + c); q3 = __ldg(g3 + c);
                        }
                        combine4(q0, q1, q2, q3, W0, W1, W2, W3, v[2 * c], v[2 * c + 1]);
                    }
                    __syncwarp();
                }
            }
            {
                TP(const long long t0 = clock64();)
                // the MMAs that read this A stage (its previous use) have retired
                umma::mbar_wait(&bar_empty[s + S * ((use - 1) & 1)], (uint32_t)(((use - 1) >> 1) & 1));
                TP(c_wait_stage += clock64() - t0;)
            }
            TP(const bool tr = hp.prof == 2 && blockIdx.x == 0 && q == 0 && lane == 0 && it_seq < (uint32_t)kTraceKB;
               if (tr) g_ttrace[it_seq][0] = clock64();)
            umma::tc_fence_after();
            // tcgen05.st is warp-collective (.sync.aligned): issued by the whole, reconverged warp
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) split2(v[c8 * 4 + i], hi[2 * i], hi[2 * i + 1], lo[2 * i], lo[2 * i + 1]);
                tmem_st8(a_col + c8 * 8, hi);
                tmem_st8(a_col + 32 + c8 * 8, lo);
            }
            TP(if (tr) g_ttrace[it_seq][1] = clock64();)
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            TP(if (tr) g_ttrace[it_seq][2] = clock64();)
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                umma::mbar_arrive(&bar_full[s + S * (use & 1)]);
                // last K block of this group inside the halo slot: this warp has read everything it needs from it
                if (last_in_slot) umma::mbar_arrive(&bar_halo_empty[hslot]);
            }
            TP(if (tr) g_ttrace[it_seq][3] = clock64();)
            {   // step to my next K block
                tap += G; ++use; TP(it_seq += G;)
                if (tap >= T) {
                    tap -= T; ++hs;
                    if (++cb == n_cb) { cb = 0; tl.t += (int)gridDim.x; if (tl.t < total) decode_tile(tl); }
                }
                tapo = s_tapoff[phys(tap)];
            }
        }
        TP(if (hp.prof == 1 && blockIdx.x == 0 && lane == 0 && q == 0)
               printf("deform tmem producer group %d: total %lld cycles, wait halo %lld, wait stage %lld\n", grpi,
                      clock64() - pt0, c_wait_halo, c_wait_stage);)
    } else {
      if (G == 3) umma::setmaxnreg_dec<40>();
      if (warp == kTTmaWarp) {
        if (lane == 0) {
            // ================================ halo loader (tensor-map TMA) ========================
            uint32_t hs = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TItem it = t_item(p, t);
                for (int cb = 0; cb < n_cb; ++cb, ++hs) {
                    const int s = (int)(hs & ns_mask);
                    umma::mbar_wait_sleep(&bar_halo_empty[s], ((hs >> ns_log2) & 1) ^ 1);
                    if (FUSE) {
                        const FuseGeom &fz = hp.fz;
                        umma::mbar_expect_tx(&bar_halo_full[s], fz.bytes);
                        for (int k = 0; k < fz.n; ++k) {
                            const CUtensorMap *tk = k == 0 ? &tm : k == 1 ? &tm1 : &tm2;
                            int x0 = it.tx * kTTW, y0 = it.ty * kTTH;
                            if (fz.th[k] != d.H || fz.tw[k] != d.W) {      // first source row / column of the patch
                                int i1; float l0, l1;
                                fuse_src_index(it.ty * kTTH, fz.th[k], (float)fz.th[k] / (float)d.H, y0, i1, l0, l1);
                                fuse_src_index(it.tx * kTTW, fz.tw[k], (float)fz.tw[k] / (float)d.W, x0, i1, l0, l1);
                            }
                            umma::tma_load_4d(halo0 + (size_t)s * hp.slot_bytes + fz.off[k], tk, cb * 32, x0, y0, it.b,
                                              &bar_halo_full[s]);
                        }
                        continue;
                    }
                    umma::mbar_expect_tx(&bar_halo_full[s], hp.lines * 128);
                    umma::tma_load_4d(halo0 + (size_t)s * hp.slot_bytes, &tm, it.grp * d.Cg + cb * 32,
                                      it.tx * kTTW * d.stride - d.pad - hp.margin_x,
                                      it.ty * kTTH * d.stride - d.pad - hp.margin_y, it.b,
                                      &bar_halo_full[s]);
                }
            }
        }
    } else if (warp == kTLoadWarp) {
        if (lane == 0) {
            // ================================ weight loader ========================================
            if (TAIL) {                                  // the 1x1 tail's packed weights stay resident
                const uint32_t bytes = (uint32_t)(BN / 32) * 2u * (uint32_t)p.tail_cout * 128u;
                umma::mbar_expect_tx(&bar_tailw, bytes);
                umma::bulk_g2s(tailw, p.tail_wpack, bytes, &bar_tailw);
            }
            // K block i = use u of stage s.  Its weight block goes to ring slot (s, u & 1) [2 S slots] or s [S slots]
            // and may be requested as soon as the slot's previous block has been consumed: use u - 2 / u - 1.
            TP(uint32_t itc = 0;)
            int s = 0, use = 0;
            const bool deep = hp.sb == 2 * S;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const TItem it = t_item(p, t);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.wpack) +
                                     (size_t)(it.grp * p.n_tiles_n + it.nt) * p.KB * kBTile;
                for (int cb = 0; cb < n_cb; ++cb)
                    for (int tap = 0; tap < T; ++tap) {
                        const int prev = deep ? use - 2 : use - 1;           // negative: parity 1 of a fresh barrier passes
                        umma::mbar_wait(&bar_empty[s + S * (prev & 1)], (uint32_t)((prev >> 1) & 1));
                        TP(if (hp.prof == 2 && blockIdx.x == 0 && itc < (uint32_t)kTraceKB) g_ttrace[itc][7] = clock64(); ++itc;)
                        uint64_t *full = &bar_full[s + S * (use & 1)];
                        uint8_t *dst = smem + (size_t)(deep ? s + S * (use & 1) : s) * kBTile;
                        umma::mbar_expect_tx(full, kBTile);
                        umma::bulk_g2s(dst, src + (size_t)(phys(tap) * n_cb + cb) * kBTile, kBTile, full);
                        if (++s == S) { s = 0; ++use; }
                    }
            }
        }
      } else if (warp == kTMmaWarp) {
        {
            // ================================ MMA issuer (A from tensor memory) ====================
            // The whole warp walks the loop (uniform control flow); one elected lane issues the tcgen05 instructions.
            // Under `if (lane == 0)` the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY retry loop.
            constexpr uint32_t idesc = umma::make_idesc_tf32(kTM, BN);
            constexpr uint32_t idesc2 = umma::make_idesc_tf32(kTM, 2 * BN);
            uint32_t ti = 0;
            TP(const bool l0 = lane == 0; uint32_t itc = 0; long long c_acc = 0, c_a = 0, c_issue = 0, t0 = clock64(); const long long t_start = t0;)
#define TPROF(acc_) TP(do { const long long t1 = clock64(); acc_ += t1 - t0; t0 = t1; } while (0))
            const int nkb = n_cb * T;
            // TAIL: the 1x1 convolution of tile j: A = activated tile in TMEM (hi at the accumulator's columns
            // [0, BN), lo at [BN, 2 BN)), B = resident tail weights ([B_hi | B_lo] blocks of 32 input channels)
            auto issue_tail = [&](uint32_t j) {
                const int a = j & 1;
                const int Ct = p.tail_cout;
                const uint32_t idesc_t = umma::make_idesc_tf32(kTM, Ct);
                if (j == 0) umma::mbar_wait_sleep(&bar_tailw, 0);
                umma::mbar_wait_sleep(&bar_y2[a], (j >> 1) & 1);
                umma::mbar_wait_sleep(&bar_acc3_empty, (j & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t a0 = tmem_base + kTAccCol + a * 128, d3 = tmem_base + kTAcc3Col;
                if (umma::elect_one()) {
#pragma unroll 1
                    for (int kb = 0; kb < BN / 32; ++kb) {
                        const uint32_t wb = umma::smem_u32(tailw) + (uint32_t)kb * 2u * (uint32_t)Ct * 128u;
                        const uint64_t b_hi = umma::make_desc_sw128(wb), b_lo = umma::make_desc_sw128(wb + (uint32_t)Ct * 128u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t ah = a0 + kb * 32 + k * 8, al = a0 + BN + kb * 32 + k * 8;
                            mma_tf32_ts(d3, ah, umma::desc_advance(b_hi, k * 32), idesc_t, (kb | k) != 0);
                            mma_tf32_ts(d3, ah, umma::desc_advance(b_lo, k * 32), idesc_t, 1);
                            mma_tf32_ts(d3, al, umma::desc_advance(b_hi, k * 32), idesc_t, 1);
                        }
                    }
                    umma::tc_commit(&bar_acc3_full);
                }
                __syncwarp();
            };
            // The loop below is the critical path of the whole CTA: one thread feeds the tensor pipe, and everything it
            // executes between two K blocks is time the pipe drains (trace of the first version: 490 of 980 cycles per
            // K block).  Hence ONE barrier per stage (A arrivals + weight bytes), a wrapping stage counter instead of
            // % and /, no instrumentation unless built for it.
            uint32_t stage = 0, use = 0;                      // K block = use `use` of stage `stage` (see bar_full)
            const uint32_t b_base = umma::smem_u32(smem);
            const uint32_t deep = hp.sb == 2 * S ? (uint32_t)S : 0u;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
                const int a = ti & 1;
                if (!TAIL) {    // TAIL: tile i-2's tail MMAs (which follow its activation pass) precede this tile in the pipe
                    umma::mbar_wait_sleep(&bar_acc_empty[a], ((ti >> 1) & 1) ^ 1);
                    umma::tc_fence_after();
                }
                TPROF(c_acc);
                const uint32_t d_tmem = tmem_base + kTAccCol + a * 128;
#pragma unroll 1
                for (int kb = 0; kb < nkb; ++kb) {
                    TP(const bool tr = l0 && hp.prof == 2 && blockIdx.x == 0 && itc < (uint32_t)kTraceKB;)
                    const uint32_t bi = stage + S * (use & 1u);       // barrier pair member of this use
                    if (hp.spin == 1) umma::mbar_wait(&bar_full[bi], (use >> 1) & 1u);
                    else umma::mbar_wait_sleep(&bar_full[bi], (use >> 1) & 1u);
                    umma::tc_fence_after();
                    TPROF(c_a);
                    TP(if (tr) g_ttrace[itc][4] = g_ttrace[itc][5] = t0;)
                    const uint32_t a_col = tmem_base + kTACol + stage * kTAStageCols;
                    const uint64_t b_hi = umma::make_desc_sw128(b_base + (stage + deep * (use & 1u)) * (uint32_t)kBTile);
                    if (umma::elect_one()) {
                        mma_tf32_ts(d_tmem, a_col, b_hi, idesc2, kb != 0);
                        mma_tf32_ts(d_tmem, a_col + 32, b_hi, idesc, 1);
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            mma_tf32_ts(d_tmem, a_col + k * 8, umma::desc_advance(b_hi, k * 32), idesc2, 1);
                            mma_tf32_ts(d_tmem, a_col + 32 + k * 8, umma::desc_advance(b_hi, k * 32), idesc, 1);
                        }
                        umma::tc_commit(&bar_empty[bi]);
                    }
                    __syncwarp();
                    TPROF(c_issue);
                    TP(if (tr) g_ttrace[itc][6] = t0; ++itc;)
                    if (++stage == S) { stage = 0; ++use; }
                }
                if (umma::elect_one()) umma::tc_commit(&bar_acc_full[a]);
                __syncwarp();
                if (TAIL && ti >= 1) issue_tail(ti - 1);
            }
            if (TAIL && ti >= 1) issue_tail(ti - 1);
            TP(if (l0 && hp.prof == 1 && blockIdx.x == 0)
                   printf("deform tmem MMA thread: %u tiles, total %lld cycles; wait acc %lld, wait A+B %lld, issue %lld\n",
                          ti, clock64() - t_start, c_acc, c_a, c_issue);)
        }
    }
    }
    if (!triggered) pdl_trigger();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == kTMmaWarp) {
        umma::tc_fence_after();
        umma::tmem_dealloc<512>(tmem_base);
    }
#ifdef AANET_TMEM_PROF
    if (hp.prof == 2 && blockIdx.x == 0 && tid == 0) {
        const long long z = g_ttrace[0][0];
        printf("# K block: producer [stage free seen, stores issued, wait::st done, arrived]  MMA thread [A + B seen, issued + committed]  weight copy issued\n");
        for (int i = 0; i < kTraceKB && i < n_cb * T * 3; ++i)
            printf("kb %2d: %6lld %6lld %6lld %6lld | %6lld %6lld | %6lld\n", i, g_ttrace[i][0] - z, g_ttrace[i][1] - z,
                   g_ttrace[i][2] - z, g_ttrace[i][3] - z, g_ttrace[i][4] - z, g_ttrace[i][6] - z, g_ttrace[i][7] - z);
    }
#endif
}

// --------------------------------------------------------------------------------------------- host side
static size_t tail_bytes(const ConvParams &p, int BN) {
    return p.tail_wpack ? (size_t)(BN / 32) * 2 * p.tail_cout * 128 : 0;
}

template <int BN, bool DENSE, bool LEAN, int G, bool TAIL = false, int SUB = 1, bool FUSE = false>
static int tmem_launch_g(const DeformTmemParams &hp, const CUtensorMap &tm, cudaStream_t stream,
                         const CUtensorMap *tm1 = nullptr, const CUtensorMap *tm2 = nullptr) {
    const size_t smem = (size_t)hp.sb * 2 * BN * 32 * 4 + ((size_t)hp.slot_bytes << hp.ns_log2) + tail_bytes(hp.p, BN) + 1024;
    cudaFuncSetAttribute(deform_tmem_kernel<BN, DENSE, LEAN, G, TAIL, SUB, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rounds = ceil_div(hp.p.total_tiles, num_sms());
    const int grid = ceil_div(hp.p.total_tiles, rounds);
    return launch_pdl(deform_tmem_kernel<BN, DENSE, LEAN, G, TAIL, SUB, FUSE>, dim3(grid), dim3(TCfg<G>::kThreads), smem, stream, hp,
                      tm, tm1 ? *tm1 : tm, tm2 ? *tm2 : tm);
}

// the fused-tail instantiations exist for the 3-group pipeline, lean epilogue, BN = 32 / 64
template <bool DENSE>
static int tmem_launch_tail(const DeformTmemParams &hp, const CUtensorMap &tm, int BN, cudaStream_t stream) {
    return BN == 64 ? tmem_launch_g<64, DENSE, true, 3, true>(hp, tm, stream)
                    : tmem_launch_g<32, DENSE, true, 3, true>(hp, tm, stream);
}

static int tmem_groups(bool dense) {
    const char *e = getenv(dense ? "AANET_DENSE_GROUPS" : "AANET_DEFORM_GROUPS");
    const int g = e ? atoi(e) : 3;      // measured: 4 groups buy nothing (dense) or cost spills (deformable)
    return g == 4 && dense ? 4 : 3;     // (deformable: the geometry slots of a 4th group do not fit the static smem)
}

template <int BN, bool DENSE, bool LEAN>
static int tmem_launch_inst(const DeformTmemParams &hp, const CUtensorMap &tm, int groups, cudaStream_t stream) {
    return groups == 4 ? tmem_launch_g<BN, DENSE, LEAN, 4>(hp, tm, stream) : tmem_launch_g<BN, DENSE, LEAN, 3>(hp, tm, stream);
}

// Fused 1x1 tail: one conv group, the main output is exactly one N tile of 32 / 64 channels (it becomes the tail's
// K = 1 or 2 blocks), lean main epilogue (no NCHW output / residual / offset-mask head), tail width a multiple of 16.
static bool tmem_tail_ok(const ConvParams &p, int BN) {
    const MdcnDims &d = p.d;
    return d.groups == 1 && d.Cout == BN && (BN == 32 || BN == 64) && !p.out_nchw && !p.residual &&
           p.act != ACT_OFFSET_MASK && p.tail_cout >= 16 && p.tail_cout <= 64 && p.tail_cout % 16 == 0 &&
           aligned16(p.tail_wpack) && (p.tail_act == ACT_NONE || p.tail_act == ACT_RELU || p.tail_act == ACT_LEAKY);
}

// Shared plan of the two modes; margin = pixels of learned offset the staged patch covers (0 for DENSE).
static int tmem_plan(const ConvParams &src, int BN, int margin_req, int groups, DeformTmemParams &hp, CUtensorMap &tm) {
    const MdcnDims &d = src.d;
    hp.p = src;
    hp.fz = FuseGeom{};
    const size_t ring = (size_t)groups * 2 * BN * 32 * 4 + tail_bytes(src, BN) + (src.offset && d.Cd == 16 ? 5 * 1024 : 0);   // (+ the second sample's geometry slots, static)
    // Halo plan.  The pitch (pixels per halo row) is rounded up to a multiple of 8 lines: a sample whose row index
    // jitters by one (sub-pixel offsets of either sign) then keeps its swizzle key (line & 7), so neighbouring
    // threads keep hitting different bank groups; the extra columns widen the horizontal margin.  Vertical margin:
    // the largest (<= the request) for which two slots fit next to the weight ring.
    const char *ep8 = getenv("AANET_DEFORM_PITCH8");
    const bool pitch8 = !(ep8 && ep8[0] == '0');
    const int corner = margin_req > 0 || src.offset ? 1 : 0;      // +1 line / column for the bilinear corner
    int margin = margin_req, mx = 0;
    for (; margin >= 0; --margin) {
        hp.HH = (kTTH - 1) * d.stride + 1 + (d.kh - 1) * d.dil + 2 * margin + corner;
        const int wmin = (kTTW - 1) * d.stride + 1 + (d.kw - 1) * d.dil + 2 * margin + corner;
        hp.HWd = (pitch8 && d.stride == 1) ? (wmin + 7) / 8 * 8 : wmin;
        mx = margin + (hp.HWd - wmin) / 2;
        hp.lines = hp.HH * hp.HWd;
        hp.slot_bytes = (hp.lines * 128 + 1023) & ~1023;
        if (hp.HH <= 256 && hp.HWd <= 256 && ring + 2 * (size_t)hp.slot_bytes <= (size_t)kTSmemBudget) break;
    }
    if (margin < 0) return AANET_ERR_UNSUPPORTED;
    hp.margin_y = margin;
    hp.margin_x = src.offset ? mx : 0;
    {   // weight ring: twice as deep where the patch slots leave room (the dense layers; the deformable layer's margin
        // is worth more than the ring depth).  AANET_TMEM_DEEP=0: A/B switch.
        const char *ed = getenv("AANET_TMEM_DEEP");
        const size_t extra = (size_t)groups * 2 * BN * 32 * 4;
        hp.sb = groups;
        hp.ns_log2 = 1;
        // fewer taps than producer groups (1x1): a slot holds one K block, so two slots mean two K blocks in flight per
        // CTA -- take four
        const char *e4 = getenv("AANET_TMEM_SLOTS4");
        if (d.K < groups && !(e4 && e4[0] == '0') && ring + 4 * (size_t)hp.slot_bytes <= (size_t)kTSmemBudget) hp.ns_log2 = 2;
        if (!(ed && ed[0] == '0') && ring + extra + ((size_t)hp.slot_bytes << hp.ns_log2) <= (size_t)kTSmemBudget) hp.sb = 2 * groups;
    }
    hp.n_cb = d.Cg / 32;
    { const char *ep = getenv("AANET_HALO_PROF"); hp.prof = ep ? atoi(ep) : 0; }
    { const char *ep = getenv("AANET_MMA_SPIN"); hp.spin = ep ? atoi(ep) : 0; }
    { const char *ep = getenv("AANET_TMEM_ROT"); hp.rot = ep ? atoi(ep) : 0; }
    ConvParams &p = hp.p;
    p.n_tiles_n = ceil_div(d.Og, BN);
    p.K = d.K * d.Cg;
    p.KB = p.K / 32;
    p.tiles_x = ceil_div(d.Wo, kTTW);
    p.tiles_per_img = p.tiles_x * ceil_div(d.Ho, kTTH);
    p.n_ptiles = d.B * p.tiles_per_img;
    const long total = (long)d.groups * p.n_tiles_n * p.n_ptiles;
    if (total > 0x3fffffffL) return AANET_ERR_UNSUPPORTED;
    p.total_tiles = (int)total;
    const uint64_t dims[4] = {(uint64_t)d.Cin, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    const uint64_t strides[3] = {(uint64_t)d.Cin * 4, (uint64_t)d.W * d.Cin * 4, (uint64_t)d.HW * d.Cin * 4};
    const uint32_t box[4] = {32, (uint32_t)hp.HWd, (uint32_t)hp.HH, 1};
    return make_tensor_map_f32(&tm, src.x, 4, dims, strides, box, 1);
}

// DCNv2.  Returns AANET_ERR_UNSUPPORTED when the problem should take another kernel.
int deform_tmem_launch(const ConvParams &src, int BN, cudaStream_t stream) {
    { const char *e = getenv("AANET_DEFORM_TMEM"); if (e && e[0] == '0') return AANET_ERR_UNSUPPORTED; }   // A/B switch
    const MdcnDims &d = src.d;
    // deformable groups of >= 32 channels (one bilinear sample per 32-channel K block) or of exactly 16 (two samples)
    const bool sub2 = d.Cd == 16;
    if (d.stride != 1 || d.Cg % 32 || (d.Cd % 32 && !sub2) || d.Cin > 64 * 32 || !aligned16(src.x)) return AANET_ERR_UNSUPPORTED;
    if (sub2 && BN != 32) return AANET_ERR_UNSUPPORTED;      // instantiated for the 32-channel scale
    if (src.out_nchw || src.residual || d.Og % 16 || (d.Cout & 3) || src.act == ACT_OFFSET_MASK) return AANET_ERR_UNSUPPORTED;
    if (BN != 32 && BN != 64) return AANET_ERR_UNSUPPORTED;
    const bool tail = src.tail_wpack != nullptr;
    if (tail && !tmem_tail_ok(src, BN)) return AANET_ERR_UNSUPPORTED;
    const int groups = tail ? 3 : tmem_groups(false);
    if (d.K < groups) return AANET_ERR_UNSUPPORTED;         // every producer group must own a tap in every halo slot
    const char *em = getenv("AANET_DEFORM_MARGIN");
    DeformTmemParams hp;
    CUtensorMap tm;
    const int rc = tmem_plan(src, BN, em ? atoi(em) : 4, groups, hp, tm);
    if (rc) return rc;
    if (sub2) return tail ? tmem_launch_g<32, false, true, 3, true, 2>(hp, tm, stream)
                          : tmem_launch_g<32, false, true, 3, false, 2>(hp, tm, stream);
    if (tail) return tmem_launch_tail<false>(hp, tm, BN, stream);
    return BN == 64 ? tmem_launch_inst<64, false, true>(hp, tm, groups, stream)
                    : tmem_launch_inst<32, false, true>(hp, tm, groups, stream);
}

// Stride-1 dense convolution with at least as many taps as producer groups and 32-channel blocks (3x3 layers of the ISA blocks and
// the offset/mask head).  Returns AANET_ERR_UNSUPPORTED when the problem should take the round-1 engine.
int dense_tmem_launch(const ConvParams &src, int BN, cudaStream_t stream) {
    { const char *e = getenv("AANET_DENSE_TMEM"); if (e && e[0] == '0') return AANET_ERR_UNSUPPORTED; }   // A/B switch
    const MdcnDims &d = src.d;
    if ((d.stride != 1 && d.stride != 2) || d.Cg % 32 || !aligned16(src.x) || src.offset) return AANET_ERR_UNSUPPORTED;
    // residual: channels-last only, added after the affine and before the activation (as the round-1 engine)
    if (src.residual && (src.out_nchw || src.tail_wpack || !aligned16(src.residual) || (d.Cout & 3))) return AANET_ERR_UNSUPPORTED;
    if (BN != 32 && BN != 48 && BN != 64) return AANET_ERR_UNSUPPORTED;
    const bool tail = src.tail_wpack != nullptr;
    if (tail && !tmem_tail_ok(src, BN)) return AANET_ERR_UNSUPPORTED;
    const int groups = tail ? 3 : tmem_groups(true);
    // (fewer taps than producer groups -- the 1x1 convolutions -- are fine: the groups then skip halo slots)
    if (tail && d.K < groups) return AANET_ERR_UNSUPPORTED;
    DeformTmemParams hp;
    CUtensorMap tm;
    const int rc = tmem_plan(src, BN, 0, groups, hp, tm);
    if (rc) return rc;
    if (tail) return tmem_launch_tail<true>(hp, tm, BN, stream);
    const bool lean = !src.out_nchw && src.act != ACT_OFFSET_MASK && d.Og % 16 == 0 && (d.Cout & 3) == 0;
    switch (BN) {
        case 32: return lean ? tmem_launch_inst<32, true, true>(hp, tm, groups, stream) : tmem_launch_inst<32, true, false>(hp, tm, groups, stream);
        case 48: return lean ? tmem_launch_inst<48, true, true>(hp, tm, groups, stream) : tmem_launch_inst<48, true, false>(hp, tm, groups, stream);
        case 64: return lean ? tmem_launch_inst<64, true, true>(hp, tm, groups, stream) : tmem_launch_inst<64, true, false>(hp, tm, groups, stream);
    }
    return AANET_ERR_UNSUPPORTED;
}

bool tmem_tail_supported(const ConvParams &src, bool deform) {
    const MdcnDims &d = src.d;
    const int BN = conv_umma_pick_bn(d.Og);
    if (!src.tail_wpack || !tmem_tail_ok(src, BN) || d.stride != 1 || d.Cg % 32 || d.K < 3 || !aligned16(src.x)) return false;
    { const char *e = getenv(deform ? "AANET_DEFORM_TMEM" : "AANET_DENSE_TMEM"); if (e && e[0] == '0') return false; }
    { const char *e = getenv("AANET_TAIL_FUSION"); if (e && e[0] == '0') return false; }
    if (deform && d.Cd % 32 && !(d.Cd == 16 && BN == 32)) return false;
    DeformTmemParams hp;
    CUtensorMap tm;
    return tmem_plan(src, BN, deform ? 4 : 0, 3, hp, tm) == AANET_OK;
}

// CSA resize-and-sum + LeakyReLU fused into the next bottleneck's conv1 (1x1 + bn1 + ReLU): `terms` are channels-last
// [B][th][tw][C] tensors, each either (H, W)-sized or smaller (bilinear, align_corners = False); the sum goes to
// fused_out [B][H*W][C] and conv1 of it to conv.out.  conv: a ConvParams of the 1x1 convolution (x unused).
int csa_conv1_tmem_launch(const float *const *terms, const int *th, const int *tw, int n_terms, float slope,
                          float *fused_out, const ConvParams &conv, cudaStream_t stream) {
    const MdcnDims &d = conv.d;
    if (n_terms < 1 || n_terms > kFuseMaxTerms || d.K != 1 || d.stride != 1 || d.pad != 0 || d.groups != 1) return AANET_ERR_UNSUPPORTED;
    if (d.Cin % 32 || d.Cin > 64 * 32 || (d.Cout != 32 && d.Cout != 64) || conv.out_nchw || conv.residual || conv.tail_wpack ||
        conv.act == ACT_OFFSET_MASK)
        return AANET_ERR_UNSUPPORTED;
    // fused_out == NULL: the sum is only consumed by the convolution (the last module: final 1x1 + soft-argmin)
    if (fused_out && !aligned16(fused_out)) return AANET_ERR_UNSUPPORTED;
    DeformTmemParams hp;
    hp.p = conv;
    hp.p.x = terms[0];
    FuseGeom &fz = hp.fz;
    fz = FuseGeom{};
    fz.n = n_terms; fz.slope = slope; fz.out = fused_out;
    CUtensorMap tms[kFuseMaxTerms];
    int off = 0;
    for (int k = 0; k < n_terms; ++k) {
        if (!terms[k] || !aligned16(terms[k]) || th[k] <= 0 || tw[k] <= 0 || th[k] > d.H || tw[k] > d.W) return AANET_ERR_UNSUPPORTED;
        const bool same = th[k] == d.H && tw[k] == d.W;
        if (!same && (th[k] == d.H || tw[k] == d.W)) return AANET_ERR_UNSUPPORTED;      // both axes resized, or none
        fz.th[k] = th[k]; fz.tw[k] = tw[k];
        // bilinear footprint of kTTH (kTTW) consecutive destination pixels: span * scale + both neighbours
        fz.ph[k] = same ? kTTH : (int)((long)(kTTH - 1) * th[k] / d.H) + 3;
        fz.pw[k] = same ? kTTW : (int)((long)(kTTW - 1) * tw[k] / d.W) + 3;
        fz.off[k] = off;
        fz.bytes += fz.ph[k] * fz.pw[k] * 128;
        off += (fz.ph[k] * fz.pw[k] * 128 + 1023) & ~1023;
        const uint64_t dims[4] = {(uint64_t)d.Cin, (uint64_t)tw[k], (uint64_t)th[k], (uint64_t)d.B};
        const uint64_t strides[3] = {(uint64_t)d.Cin * 4, (uint64_t)tw[k] * d.Cin * 4, (uint64_t)th[k] * tw[k] * d.Cin * 4};
        const uint32_t box[4] = {32, (uint32_t)fz.pw[k], (uint32_t)fz.ph[k], 1};
        const int rc = make_tensor_map_f32(&tms[k], terms[k], 4, dims, strides, box, 1);
        if (rc) return rc;
    }
    const int BN = d.Cout;
    hp.HH = kTTH; hp.HWd = kTTW; hp.lines = kTTH * kTTW; hp.slot_bytes = off;
    hp.margin_y = hp.margin_x = 0; hp.n_cb = d.Cin / 32;
    hp.prof = 0; hp.spin = 0; hp.rot = 0; hp.sb = 3;
    if ((size_t)3 * 2 * BN * 32 * 4 + 2 * (size_t)off > (size_t)kTSmemBudget) return AANET_ERR_UNSUPPORTED;
    { const char *e4 = getenv("AANET_TMEM_SLOTS4");
      hp.ns_log2 = (!(e4 && e4[0] == '0') && (size_t)3 * 2 * BN * 32 * 4 + 4 * (size_t)off <= (size_t)kTSmemBudget) ? 2 : 1; }
    ConvParams &p = hp.p;
    p.n_tiles_n = 1;
    p.K = d.Cin; p.KB = p.K / 32;
    p.tiles_x = ceil_div(d.W, kTTW);
    p.tiles_per_img = p.tiles_x * ceil_div(d.H, kTTH);
    p.n_ptiles = d.B * p.tiles_per_img;
    if ((long)p.n_ptiles > 0x3fffffffL) return AANET_ERR_UNSUPPORTED;
    p.total_tiles = p.n_ptiles;
    const CUtensorMap *t1 = n_terms > 1 ? &tms[1] : nullptr, *t2 = n_terms > 2 ? &tms[2] : nullptr;
    if (conv.act == ACT_SOFTARGMIN)     // non-lean epilogue: reduces the candidates to one disparity per pixel
        return BN == 64 ? tmem_launch_g<64, true, false, 3, false, 1, true>(hp, tms[0], stream, t1, t2)
                        : tmem_launch_g<32, true, false, 3, false, 1, true>(hp, tms[0], stream, t1, t2);
    return BN == 64 ? tmem_launch_g<64, true, true, 3, false, 1, true>(hp, tms[0], stream, t1, t2)
                    : tmem_launch_g<32, true, true, 3, false, 1, true>(hp, tms[0], stream, t1, t2);
}

}  // namespace aanet
