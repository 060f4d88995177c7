// Blackwell (sm_100a) tensor-path primitives as inline PTX: mbarrier, TMA tensor loads, tensor memory (TMEM)
// allocation / load / store and tcgen05.mma with shared-memory matrix descriptors.  No CUTLASS: the bit layouts
// below are the documented descriptor formats (PTX ISA "tcgen05 matrix descriptors"; the same fields CUTLASS names
// in cute/arch/mma_sm100_desc.hpp — start_address [0,14), leading_byte_offset [16,30), stride_byte_offset [32,46),
// version [46,48), layout_type [61,64)).
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the driver entry point is resolved at run time, libcuda is not linked)
#include <stdint.h>

namespace ktb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// blocks until the phase with the given parity has completed (a fresh barrier passes parity 1 at once)
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (TMA, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled load: box (c0 .. c0+box0, c1 .. c1+box1) -> shared memory, completion (bytes) on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(m), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------- TMEM
// One warp allocates `cols` (power of two >= 32) columns; the base address lands in shared memory at `dst`.
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address: bits [31:16] lane, [15:0] column.  A warp can only touch the 32 lanes of its sub-partition
// (warp_id % 4); thread i of the warp gets lane (32 * (warp_id % 4) + i), registers r[0..N) = N consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]),
        "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(taddr)
        : "memory");
}

// ---------------------------------------------------------------------------------------------- descriptors
enum : uint64_t { kLayoutNone = 0, kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutSw32 = 6 };

// Shared-memory matrix descriptor.  `addr` is a shared::cta byte address (16-byte aligned), lbo / sbo in bytes.
//   K-major,  swizzle 128B : rows of 128 B (64 bf16 / 128 int8 along K), 8-row groups `sbo` = 1024 B apart; lbo unused
//   K-major,  no swizzle   : core matrix = 8 rows x 16 B contiguous; next core matrix along K at `lbo`, next 8 rows at `sbo`
//   MN-major, swizzle 128B : 128 B (64 bf16) contiguous along MN, 8 K-rows per 1024-B atom; next MN atom at `lbo`,
//                            next 8 K-rows at `sbo`
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46) | (layout << 61);
}
// Instruction descriptor (upper 32 bits of the PTX idesc operand): kind::f16 / kind::i8, dense.
//   c_format [4,6): 0 f16, 1 f32, 2 s32 ; a_format [7,10), b_format [10,13): f16 kinds 0 f16 / 1 bf16, i8 kind 0 u8 / 1 s8 ;
//   a_major bit 15, b_major bit 16 (0 K-major, 1 MN-major) ; n_dim [17,23) = N >> 3 ; m_dim [24,29) = M >> 4
__host__ __device__ constexpr uint32_t instr_desc(int c_fmt, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major, int M, int N) {
    return ((uint32_t)c_fmt << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem].  Called by ALL lanes of a converged warp with warp-uniform operands; ONE elected lane
// issues the instruction for the whole CTA.  (Issuing from inside `if (lane == 0)` makes the compiler treat every operand
// as per-thread: each MMA then costs ~20 SASS instructions — register -> uniform-register moves and an ELECT loop —
// and a dependent ~100-cycle issue chain; measured in profiles/mla_trace_r02.txt.)  bf16/f16 inputs: K = 16 per instruction.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 r;\n"
        "elect.sync r|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// int8 inputs, int32 accumulate: K = 32 per instruction (sm_100a has the integer tensor path; sm_103a does not)
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 r;\n"
        "elect.sync r|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// e4m3 / e5m2 / 6- and 4-bit float inputs, fp32 accumulate: K = 32 per instruction for the 8-bit types
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 r;\n"
        "elect.sync r|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// makes `bar` complete (one arrival) when all tcgen05 operations issued so far have finished; called by a converged
// warp like mma_f16 (the same elected lane issues it)
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .b32 r;\n"
        "elect.sync r|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(bar)
        : "memory");
}

}  // namespace umma
}  // namespace ktb
