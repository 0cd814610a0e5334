// umma.cuh -- tcgen05 (5th-gen tensor core) building blocks for fp32-faithful GEMM tiles.
//
// The reference computes every Linear in full fp32 (torch's allow_tf32 default is off and it is
// never changed).  tcgen05 has no fp32 kind, so the product uses the error-compensated
// 3xTF32 scheme: each fp32 operand is split a = hi + lo with hi = round_tf32(a), lo = a - hi (exact)
// and   a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b   accumulated in fp32 in tensor memory.
// The dropped lo*lo term is <= 2^-22 |a||b|, i.e. at fp32 rounding level.
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout:
//   atom  = 8 rows x 128 B (32 fp32 of K); 16-byte chunk c of row r is stored at chunk c ^ (r & 7);
//   a tile of R rows x K fp32 = K/32 "k-atoms" blocks of R*128 B each, rows grouped by 8 (1024 B).
// One tcgen05.mma.kind::tf32 consumes K = 8 fp32 (32 B) per operand row.
#pragma once

#include "common.cuh"

namespace ltgnn {
namespace umma {

using ptx::smem_u32;

// ---- fp32 -> (hi, lo) TF32 split ------------------------------------------------------------
// hi = x rounded to TF32 (10-bit mantissa) with two integer ops (add half an ulp of TF32, clear the 13 low
// bits: round-half-away, finite inputs); lo = x - hi is exact in fp32 and is handed to the tensor core as
// is -- kind::tf32 reads the upper 19 bits of the container, so lo is truncated to TF32 there, an error of
// at most 2^-11 |lo| <= 2^-23 |x|.  (cvt.rna.tf32.f32 expands to ~4 SASS instructions with NaN handling;
// the loaders are instruction-bound, so this matters.)
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float tf32_rna(float x) { return tf32_hi(x); }
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

// byte offset of 16-byte chunk `c16` (0..K/4-1) of row `r` in an R-row K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int c16, int rows) {
    const int katom = c16 >> 3;
    const int c = c16 & 7;
    return static_cast<uint32_t>(katom * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// ---- descriptors --------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4)  // start address   [0,14)
           | (1ull << 16)                                    // LBO (unused for swizzled K-major)
           | (64ull << 32)                                   // SBO = 1024 B >> 4 [32,46)
           | (1ull << 46)                                    // descriptor version (sm_100)
           | (2ull << 61);                                   // SWIZZLE_128B
}
// instruction descriptor: D fp32, A/B TF32, both K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- tensor memory -------------------------------------------------------------------------
// whole warp: allocate `cols` (power of two >= 32) TMEM columns, base address written to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, one K=8 step, issued by ONE thread.
// The single issuing thread is a serial bottleneck (a few dozen dependent integer instructions per MMA
// cost more than the MMA itself at N = 64), so descriptors are passed as their 32-bit low word (the only
// part that changes: start address >> 4) plus a constant high word, and assembled inside the asm block.
constexpr uint32_t kDescHi = (64u /*SBO*/) | (1u << 14) /*version*/ | (2u << 29) /*SWIZZLE_128B*/;
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

__device__ __forceinline__ void mma_tf32_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 3xTF32 product of one (rows_a x 32) by (rows_b x 32) k-atom: 4 K-steps x 3 MMAs.
// a_hi/a_lo/b_hi/b_lo: shared addresses of the k-atom blocks (1024-byte aligned); first==1 overwrites D.
__device__ __forceinline__ void mma_katom_3x(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                             uint32_t b_lo, uint32_t idesc, bool first) {
    const uint32_t ah = desc_lo(a_hi), al = desc_lo(a_lo), bh = desc_lo(b_hi), bl = desc_lo(b_lo);
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {  // one K = 8 step is 32 bytes = 2 descriptor units
        mma_tf32_lo(d_tmem, al + 2 * k, bh + 2 * k, idesc, (first && k == 0) ? 0u : 1u);  // small terms first
        mma_tf32_lo(d_tmem, ah + 2 * k, bl + 2 * k, idesc, 1u);
        mma_tf32_lo(d_tmem, ah + 2 * k, bh + 2 * k, idesc, 1u);
    }
}

// warp-collective: 16 consecutive fp32 columns of this thread's TMEM lane -> registers.
// Load and wait sit in ONE asm statement so no use of the registers can be scheduled in between.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Several 16-column loads in flight, then one wait -- as separate asm statements, the way CUTLASS issues them: ptxas
// knows that tcgen05.wait::ld orders the destination registers of the preceding loads.  What must not happen is the
// C++ compiler moving a USE of the values above the wait statement; `tmem_pin16` after the wait routes every value
// through an (empty) volatile asm, and volatile asm statements keep their order, so no use can precede the wait.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_pin16(float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+f"(v[i]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
          "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}

// Three (or two) 8-column loads with ONE wait, all in one asm statement: the loads are in flight together and no use
// of the registers can be scheduled before the wait (separate asm statements would not guarantee that).
__device__ __forceinline__ void tmem_ld8x3(uint32_t ta, uint32_t tb, uint32_t tc, float* a, float* b, float* c) {
    uint32_t r[24];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(ta), "r"(tb), "r"(tc)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[8 + i]);
        c[i] = __uint_as_float(r[16 + i]);
    }
}
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float* a, float* b) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta), "r"(tb)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[8 + i]);
    }
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// 32 columns with one wait (two dependent x16 loads pay the tensor-memory latency twice)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace ltgnn
