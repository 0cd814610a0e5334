// patch.cuh -- warp-private shared-memory transposition between the two ways a warp can hold 32 rows x 128 bytes.
//
// tcgen05.ld / tcgen05.st (32x32b) want ROW form: lane l owns row l (a TMEM lane).  Global memory wants COALESCED
// form: 8 lanes cover the 128 contiguous bytes of one row, 4 rows per instruction -- 4 LSU wavefronts instead of the
// 32 a row-per-thread access costs (ncu: l1tex data-pipe wavefronts were the limiter of every row-per-thread kernel).
#pragma once

#include "common.cuh"

namespace ltgnn {
namespace patch {

constexpr uint32_t kPatchBytes = 4096;

// A 32-row x 128-byte patch, 16-byte chunk c of row r stored at chunk c ^ (r & 7): conflict-free for both access
// patterns below.  "Coalesced" form: g[k] = chunk (l & 7) of patch row 4 k + (l >> 3).  "Row" form: lane l holds
// the 8 chunks of patch row l.  Addresses are 2 + 1 registers (row 4 k + sub has (r & 7) = sub or sub + 4).
struct Patch {
    uint8_t* w0;  // coalesced-form address for even k (+ 512 k)
    uint8_t* w1;  // ... for odd k
    uint8_t* rd;  // row-form address of chunk 0; chunk j sits at rd ^ (j << 4)
    __device__ __forceinline__ Patch(uint8_t* scr, int lane) {
        const int sub = lane >> 3, ch = lane & 7;
        w0 = scr + sub * 128 + ((ch ^ sub) << 4);
        w1 = scr + sub * 128 + ((ch ^ sub ^ 4) << 4);
        rd = scr + lane * 128 + ((lane & 7) << 4);
    }
    __device__ __forceinline__ float4* co(int k) const { return reinterpret_cast<float4*>(((k & 1) ? w1 : w0) + k * 512); }
    __device__ __forceinline__ float4* row(int j) const {
        return reinterpret_cast<float4*>(reinterpret_cast<uintptr_t>(rd) ^ static_cast<uintptr_t>(j << 4));
    }
};
// coalesced form -> row form
__device__ __forceinline__ void transpose_in(const Patch& pt, const float4 (&g)[8], float (&v)[32]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) *pt.co(k) = g[k];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = *pt.row(j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    __syncwarp();
}
// row form -> coalesced form
__device__ __forceinline__ void transpose_out(const Patch& pt, const float (&v)[32], float4 (&g)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) *pt.row(j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = *pt.co(k);
    __syncwarp();
}

}  // namespace patch
}  // namespace ltgnn
