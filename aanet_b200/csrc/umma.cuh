// Thin inline-PTX layer for the Blackwell primitives the conv engine uses: mbarrier, 1-D bulk
// async copy, tcgen05 (alloc / mma kind::tf32 / commit / ld / fences) and the UMMA descriptors.
// sm_100a only.  Encodings follow the PTX ISA "tcgen05" chapter (shared-memory matrix descriptor,
// instruction descriptor); field positions are spelled out next to each builder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aanet {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait for the phase with the given parity to complete.  A bounded spin turns a protocol bug into a
// trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}

// Same, for the single-thread roles (MMA issuer, weight loader, epilogue): the suspend-time hint lets
// the thread sleep in hardware instead of spinning through issue slots the producer warps need
// (ncu on the first version: 36 % of all issued instructions were try_wait loops).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    // fast path: one plain try_wait, no timer read (SR_GLOBALTIMER costs ~100 cycles and the MMA issuer passes
    // through three waits per K block, most of them already satisfied)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    const unsigned long long t0 = globaltimer_ns();
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
        if ((spin & 63) == 63 && globaltimer_ns() - t0 > 4000000000ull) __trap();   // 4 s: protocol bug
    }
}

// Register rebalancing between warp roles; executed by every warp of an aligned 4-warp group.
template <int REGS> __device__ __forceinline__ void setmaxnreg_inc() {
#ifndef AANET_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}
template <int REGS> __device__ __forceinline__ void setmaxnreg_dec() {
#ifndef AANET_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copy engine)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------- TMEM
template <uint32_t COLS>   // power of two >= 32; executed by one full warp
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// All previously issued tcgen05.mma of this thread arrive (once) on the mbarrier when they retire.
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// One lane of a converged warp (elect.sync); the branch it guards is how a warp-uniform role issues the
// single-thread tcgen05 instructions without the compiler's per-instruction ELECT retry loops.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate; one thread issues for the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp reads TMEM lane (base_lane + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand tile stored as rows of 128 bytes (32 tf32) with the
// 128-byte swizzle (16-byte chunk index XOR (row & 7)); 8-row groups are 1024 bytes apart.
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride-dim byte offset >> 4 (= 1024 >> 4)   [46,48) descriptor version = 1 (sm_100)
//   [49,52) base offset = 0 (tile is 1024-byte aligned) [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Advance a descriptor by `bytes` inside the swizzle atom (K step of 8 tf32 = 32 bytes).
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (bytes >> 4); }

// Instruction descriptor for kind::tf32, FP32 accumulator, both operands K-major, dense:
//   [4,6) D format: 1 = F32   [7,10) A format: 2 = TF32   [10,13) B format: 2 = TF32
//   [15] A major: 0 = K       [16] B major: 0 = K          [17,23) N >> 3      [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp32 -> (hi, lo): hi = x with the 13 low mantissa bits cleared (exactly representable in tf32, so the tensor
// core reads it unchanged), lo = x - hi (exact in fp32, |lo| < 2^-10 |x|).  hi*hi + hi*lo + lo*hi then misses
// only lo*lo < 2^-20 |x||w|.  Truncation instead of cvt.rna.tf32 because the latter expands to ~5 integer
// instructions and the A producers split 32 values per thread per K block (ncu: 929 instructions per K block
// per producer warp, 60 % of them integer).
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

}  // namespace umma
}  // namespace aanet
