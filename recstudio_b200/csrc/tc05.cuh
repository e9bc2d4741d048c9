// tc05.cuh -- hand-written tcgen05 / TMEM primitives for sm_100a (inline PTX, no CUTLASS).
//
// Shared-memory operand layout used throughout (UMMA "K-major, SWIZZLE_NONE / interleaved"
// canonical layout ((8,m),(T,2)) : ((1T,SBO),(1,LBO)), T = 8 bf16 = 16 bytes):
//     element (row r, col c) of an [R x K] bf16 tile lives at byte
//         (c / 8) * (R * 16)  +  r * 16  +  (c % 8) * 2
//   i.e. the tile is a sequence of K/8 column groups, each holding all R rows as 16-byte chunks.
//   An 8-row x 16-byte core matrix is 128 contiguous bytes; SBO (next 8 rows) = 128 B,
//   LBO (next 8 columns) = R * 16 B.  One tcgen05.mma (K = 16) consumes two column groups.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace rsb {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) inside an [R x K] interleaved K-major bf16 tile
__device__ __forceinline__ uint32_t kmajor_off(int r, int c, int R) { return (uint32_t)((c >> 3) * (R * 16) + r * 16 + (c & 7) * 2); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version = 1, no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    return d;                                      // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// 32-bit instruction descriptor: bf16 x bf16 -> f32, both operands K-major, dense (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4)                 // c_format  = F32
         | (1u << 7)                 // a_format  = BF16
         | (1u << 10)                // b_format  = BF16
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one full warp executes these) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA issue (ONE thread) --------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; accumulate = 0 overwrites D
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_addr(bar)) : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: returns false (instead of hanging the GPU) if the phase does not complete
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_spins = 20000000u) {
    const uint32_t a = smem_addr(bar);
    for (uint32_t i = 0; i < max_spins; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// ---- TMEM -> registers: thread t of warp w gets 32 consecutive fp32 columns of lane 32w + t ------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of lane 32w + t
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
}  // namespace rsb
