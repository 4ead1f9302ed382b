// f2_umma.cuh -- the few tcgen05 / TMEM primitives the CNN kernels need, as inline PTX (sm_100a).
//
// Operand convention used throughout f2_cnn.cu: both MMA operands are K-major, no swizzle, stored in
// shared memory as PLANES of 8 K-values: element (row r, k) lives at
//     base + (k / 8) * plane_bytes + r * 16 + (k % 8) * 2            (bf16)
// i.e. one 16-byte chunk per (row, group of 8 K-values), rows back to back inside a plane.  This is the
// canonical "interleaved" K-major layout of the UMMA shared-memory descriptor -- 8 rows x 16 bytes core
// matrices, consecutive 8-row groups SBO = 128 bytes apart, the two K-halves of one K = 16 instruction
// LBO = plane_bytes apart -- with the property the convolutions are built on: the rows are LINEAR in
// memory (16 bytes per row), so "the same matrix, s rows further down" is the same descriptor with its
// start address advanced by 16*s bytes.  A 3x3 convolution tap (dy, dx) over an activation tensor laid
// out [channel / 8][pixel][8 channels] is exactly such a shift: s = dy * W + dx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f2_common.cuh"

namespace f2 {
namespace umma {

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit): start address, leading / stride byte offsets (all >> 4),
// descriptor version 1 (Blackwell) in bits 46..47, swizzle mode 0 (none) in bits 61..63.
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor of tcgen05.mma.kind::f16 (32 bit): D = fp32, A = B = bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
    return (1u << 4)                      // D format: F32
           | (1u << 7)                    // A format: BF16
           | (1u << 10)                   // B format: BF16
           | (0u << 15) | (0u << 16)      // A, B: K-major
           | ((uint32_t)(N >> 3) << 17)   // N / 8
           | ((uint32_t)(M >> 4) << 24);  // M / 16
}

// ---- tensor memory ---------------------------------------------------------------------------------
// Warp-collective.  `cols`: power of two, 32 .. 512.  The base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared by ordinary threads) -> visible to the async proxy
// (tcgen05.mma reads its operands through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a fully active warp (elect.sync): the tensor-core instructions are issued from warp-uniform
// code by the elected lane, which lets the compiler keep descriptors in uniform registers and emit the
// UTCHMMA without a per-thread election loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- MMA: D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread ---------------------------------------
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// All MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- accumulator -> registers: 32 lanes (this warp's quarter of the 128) x 32 / 16 consecutive columns ---
// taddr: (lane << 16) | column; the lane field must be the first lane of the calling warp's quarter.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier wait with a bounded spin (a wrong descriptor must end in an error code, not a hang) -----
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t max_polls = 1u << 22) {
    for (uint32_t i = 0; i < max_polls; ++i) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}

}  // namespace umma
}  // namespace f2
