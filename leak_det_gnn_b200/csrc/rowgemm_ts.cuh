// rowgemm_ts.cuh -- building blocks for GEMMs whose A operand lives in TENSOR MEMORY (tcgen05 "TS" form), 3xTF32.
//
//   D[128 x N] = A_tile[128 x K] * B[N x K]^T
//
// Why: in the shared-memory ("SS") skeleton of rowgemm.cuh every K = 8 step re-reads a 128 x 32 B slice of A
// three times (hi/lo products) and the loaders first have to write both copies -- at small N that traffic, not
// the tensor pipe, is the limit, and the A ring eats the shared memory the whole weight would need.  With A in
// TMEM a thread keeps a row (= TMEM lane), splits it in registers and `tcgen05.st`s hi and lo; the MMA reads A from
// TMEM and only B from shared memory, which leaves room to keep the WHOLE weight resident.
// Users: the pipe-head kernels (heads.cu) and the GRU kernels (gru.cu) build their own warp-specialised loops on
// top of these pieces (mma_tf32_ts, tmem_st32 / tmem_wait_st, fill_b).
#pragma once

#include "umma.cuh"

namespace ltgnn {
namespace rowgemm_ts {

using namespace ltgnn::ptx;
using namespace ltgnn::umma;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
          "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
          "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
          "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}

// B[n][k] = W[n * ldw + k] (transposed = 0) or W[k * ldw + n] (transposed = 1); resident, K-major SW128, hi/lo
__device__ inline void fill_b(uint8_t* b_hi, uint8_t* b_lo, const float* __restrict__ W, int ldw, int transposed, int K,
                              int N, int tid, int nthreads, float scale = 1.f) {
    if (!transposed) {
        const int k4 = K >> 2;
        for (int i = tid; i < N * k4; i += nthreads) {
            const int r = i / k4, c = i - r * k4;
            float4 hi, lo;
            float4 w = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(r) * ldw) + c);
            w.x *= scale; w.y *= scale; w.z *= scale; w.w *= scale;
            split4(w, hi, lo);
            const uint32_t off = sw128_offset(r, c, N);
            *reinterpret_cast<float4*>(b_hi + off) = hi;
            *reinterpret_cast<float4*>(b_lo + off) = lo;
        }
    } else {
        for (int i = tid; i < N * K; i += nthreads) {
            const int k = i / N, n = i - k * N;
            const float w = __ldg(W + static_cast<size_t>(k) * ldw + n) * scale;
            const float hi = tf32_hi(w), lo = w - hi;
            const uint32_t off = sw128_offset(n, k >> 2, N) + (k & 3) * 4;
            *reinterpret_cast<float*>(b_hi + off) = hi;
            *reinterpret_cast<float*>(b_lo + off) = lo;
        }
    }
}

}  // namespace rowgemm_ts
}  // namespace ltgnn
