// Tensor-map TMA (cp.async.bulk.tensor) helpers: host-side encoding through the driver entry point (no -lcuda:
// cudaGetDriverEntryPoint) and the device-side tile loads.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "umma.cuh"

namespace aanet {

// Encode a tiled fp32 tensor map of `rank` dimensions (innermost first).  strides_bytes has rank-1 entries (the
// innermost stride is the element size).  swizzle128: box rows of 128 bytes are stored with the 128-byte swizzle
// (16-byte chunk index XOR (line & 7)), the layout tcgen05 descriptors of type SWIZZLE_128B read.  Out-of-bounds
// elements are filled with zeros.  Returns an aanet_status.
// swizzle: 0 = none, 1 = SWIZZLE_128B, 2 = SWIZZLE_128B_ATOM_32B (the only layout tcgen05 accepts for MN-major
// tf32 operands: UMMA layout type SWIZZLE_128B_BASE32B).
int make_tensor_map_f32(CUtensorMap *tm, const void *base, int rank, const uint64_t *dims,
                        const uint64_t *strides_bytes, const uint32_t *box, int swizzle);

namespace umma {

__device__ __forceinline__ void tma_load_4d(void *dst_smem, const CUtensorMap *tm, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst_smem)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// K-major SWIZZLE_128B descriptor with an explicit stride-byte-offset between 8-row groups.  The start address
// may be any multiple of 128 bytes inside a swizzled buffer whose base is 1024-byte aligned: the hardware applies
// the XOR to absolute shared-memory address bits, so a window (shifted rows, arbitrary row-group pitch) into a
// TMA-written halo is a valid operand with base offset 0 (profiles/probes/tma_umma_probe.cu, measured on the B200).
__device__ __forceinline__ uint64_t make_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MN-major SWIZZLE_128B descriptor: rows of 128 bytes hold 32 consecutive M (or N) elements of one k; 8 such rows
// (8 k) form the 1024-byte swizzle atom; lbo_bytes = distance between consecutive 32-element groups along M/N,
// sbo_bytes = distance between consecutive 8-k atoms.
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MN-major tf32 operand in the SWIZZLE_128B_BASE32B layout (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of
// 128 bytes hold 32 consecutive M (or N) elements of one k; 4 such rows form the swizzle atom; lbo_bytes = distance
// between consecutive 32-element groups along M/N, sbo_bytes = distance between consecutive 4-k atoms.
__device__ __forceinline__ uint64_t make_desc_mn_tf32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// kind::tf32 instruction descriptor with selectable operand majors (bit 15 / 16: 1 = MN-major).
__host__ __device__ constexpr uint32_t make_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace umma
}  // namespace aanet
