// tgrad.cu -- tensor-core weight gradients (see tgrad.cuh) for
//   * GCNConv.lin:   dW[Do, Di] = G^T X over the B*N node rows      (autograd of models/detector.py:199)
//   * EdgeHead.mlp.0: dW1[H, 3D] = dpre^T feat and db1 = column sums of dpre over the B*P pipe rows
//                     (autograd of models/detector.py:79-87), with dpre and feat formed on the fly from the
//                     saved hidden activations / the node states -- neither (B*P, H) nor (B*P, 3D) exists.
#include "functors.cuh"
#include "tgrad.cuh"

using namespace ltgnn;
using namespace ltgnn::functors;

namespace {

// ---- pipe head.  d pre[row, j] = g[row] * w2[j] * live[row, j] with g = dlogit * scale is rank one up to the 0/1 gate,
// so with  A[j, :] = sum_rows live[row, j] * g[row] * feat[row, :]  and  a[j] = sum_rows live[row, j] * g[row]:
//      dW1[j, :] = w2[j] * A[j, :]        db1[j] = w2[j] * a[j]
//      dw2[j]    = sum_rows dlogit * hidden[row, j] = sum_c W1[j, c] * A[j, c] + b1[j] * a[j]
// (hidden = live * scale * (W1 feat + b1)): the saved hidden activations are not needed at all -- the forward keeps 1 bit
// per hidden unit instead of 4 bytes.  The GEMM operands: G = live (0 / 1: exact in TF32, so no lo copy and one MMA
// less per K step), X = F' = g * feat.  The 192 result columns are split over three CTAs per row range -- h_u, h_v and
// |h_u - h_v| -- so that each keeps its 128 x 64 totals in registers (tgrad.cuh, register-total form; a 128 x 192 total in
// tensor memory cost 3 x 96 KB of tensor-memory traffic per 32 rows and ran at 1.7 ms).  Both loaders are raw-style: with
// 6 loads per thread the kernel is latency bound unless the next chunk's loads are in flight while this one is stored
// (a first version that computed inside the fetch ran at 5.4 ms).
constexpr int kHeadSlices = 3;
struct HeadLive {
    static constexpr bool kRowFast = true;  // a warp = 32 rows x one chunk pair
    static constexpr bool kExact = true;
    using Raw = uint32_t;   // the 32-unit gate word that holds this chunk's 4 hidden units
    using Side = float4;    // this thread's share of a[4 c .. 4 c + 3] (slice-0 CTAs only)
    const uint32_t* hmask;  // [Mp, 4]: hidden unit j of a row <-> bit 31 - j % 32 of word j / 32 (heads.cu)
    const float* dlogit;    // [M]
    float scale;
    float* side_part;       // [2 * gridDim.x][128]: one partial of a[] per (CTA, loader group)
    __device__ __forceinline__ Raw raw(uint32_t row, int c) const { return __ldg(hmask + static_cast<size_t>(row) * 4 + (c >> 3)); }
    __device__ __forceinline__ float aux(uint32_t row) const { return __ldg(dlogit + row); }
    __device__ __forceinline__ float4 convert(const Raw& w, float d, int c, Side& side) const {
        const uint32_t sh = 28 - 4 * (c & 7);  // units 4 (c % 8) .. + 3 sit at bits 31 - 4 (c % 8) .. 28 - 4 (c % 8)
        const float4 live = make_float4((w >> (sh + 3)) & 1u ? 1.f : 0.f, (w >> (sh + 2)) & 1u ? 1.f : 0.f,
                                        (w >> (sh + 1)) & 1u ? 1.f : 0.f, (w >> sh) & 1u ? 1.f : 0.f);
        if (blockIdx.x % kHeadSlices == 0) {
            const float g = d * scale;
            side.x = fmaf(g, live.x, side.x); side.y = fmaf(g, live.y, side.y);
            side.z = fmaf(g, live.z, side.z); side.w = fmaf(g, live.w, side.w);
        }
        return live;
    }
    // the 32 lanes of a loader warp hold 32 different rows of the same columns: butterfly sum, lane 0 writes
    __device__ __forceinline__ void finish(Side* side, const int* cols, int n, uint32_t cta, int group, int lane) const {
        for (int j = 0; j < n; ++j) {
            float4 t = side[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
            }
            if (lane == 0) reinterpret_cast<float4*>(side_part + (static_cast<size_t>(cta) * 2 + group) * 128)[cols[j]] = t;
        }
    }
};
struct FeatRaw {
    float4 a, b;
};
struct HeadFeatSlice {
    static constexpr bool kRowFast = false;
    static constexpr int kSlices = kHeadSlices;
    using Raw = FeatRaw;
    const float4* x;   // node states [B*N, 16]
    const int2* ends;
    const float* dlogit;
    float scale;
    uint32_t P, N;
    uint64_t magic;
    __device__ __forceinline__ Raw raw(uint32_t row, int c) const {  // c < 16: chunk of this CTA's 64-column slice
        const uint32_t b = magic ? ptx::fastdiv(row, magic) : row;
        const int2 e = __ldg(ends + (row - b * P));
        const float4* xb = x + static_cast<int64_t>(b) * N * 16 + c;
        const int slice = blockIdx.x % kHeadSlices;
        Raw r;
        r.a = __ldg(xb + (slice == 1 ? e.y : e.x) * 16);
        r.b = slice == 2 ? __ldg(xb + e.y * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
        return r;
    }
    __device__ __forceinline__ float aux(uint32_t row) const { return __ldg(dlogit + row); }
    __device__ __forceinline__ float4 convert(const Raw& r, float d, int) const {
        const float g = d * scale;
        if (blockIdx.x % kHeadSlices == 2)
            return make_float4(g * fabsf(r.a.x - r.b.x), g * fabsf(r.a.y - r.b.y), g * fabsf(r.a.z - r.b.z), g * fabsf(r.a.w - r.b.w));
        return make_float4(g * r.a.x, g * r.a.y, g * r.a.z, g * r.a.w);
    }
};

// dW1[j, :] = w2[j] * A[j, :], db1[j] = w2[j] * a[j], dw2[j] = W1[j, :] . A[j, :] + b1[j] * a[j]; one warp per hidden unit j
__global__ void __launch_bounds__(128)
head_param_epilogue_kernel(const float* __restrict__ A, const float* __restrict__ a, const float* __restrict__ W1,
                           const float* __restrict__ b1, const float* __restrict__ w2, float* __restrict__ dW1,
                           float* __restrict__ db1, float* __restrict__ dw2, int H, int F) {
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= H) return;
    const float wj = __ldg(w2 + j);
    float dot = 0.f;
    for (int c = lane; c < F; c += 32) {
        const float v = A[j * F + c];
        dW1[j * F + c] = wj * v;
        dot = fmaf(__ldg(W1 + j * F + c), v, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        const float aj = a[j];
        db1[j] = wj * aj;
        dw2[j] = fmaf(__ldg(b1 + j), aj, dot);
    }
}

}  // namespace

extern "C" int64_t ltgnn_pipe_head_ws_floats(int device) {
    const DeviceInfo* di = device_info(device);  // tgrad partials [sm][128][64] + partials of a[] [2 sm][128] + A [128][192] + a [128]
    return di ? static_cast<int64_t>(di->sm_count) * (tgrad::kMo * 64 + 2 * 128) + 128 * 192 + 128 : -1;
}

extern "C" int64_t ltgnn_tgrad_ws_floats(int device, int32_t No) {
    const DeviceInfo* di = device_info(device);
    return di ? static_cast<int64_t>(di->sm_count) * tgrad::kMo * No : -1;
}

// dW[Do, Di] = G[M, Do]^T X[M, Di];  (Do, Di) in {(64, 64), (128, 128), (128, 64), (64, 128), ...}: Do in {64, 128},
// Di a multiple of 32 <= 256.  ws: ltgnn_tgrad_ws_floats(device, Di) floats.
extern "C" int ltgnn_wgrad_tc(int device, int64_t M, int32_t Do, int32_t Di, const float* G, const float* X, float* dW,
                              int accumulate, float* ws, void* stream_) {
    LTGNN_REQUIRE(M >= 0, LTGNN_E_ARG, "wgrad_tc: M=%lld", static_cast<long long>(M));
    LTGNN_REQUIRE(Do == 64 || Do == 128, LTGNN_E_SHAPE, "wgrad_tc: Do=%d must be 64 or 128", Do);
    LTGNN_REQUIRE(Di % 32 == 0 && Di > 0 && Di <= 192, LTGNN_E_SHAPE, "wgrad_tc: Di=%d must be a multiple of 32, <= 192", Di);
    LTGNN_REQUIRE(G && X && dW && ws, LTGNN_E_ARG, "wgrad_tc: null tensor");
    LTGNN_REQUIRE(aligned16(G) && aligned16(X), LTGNN_E_ALIGN, "wgrad_tc: G/X must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (M == 0) {
        if (!accumulate) LTGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * Do * Di, stream));
        return LTGNN_OK;
    }
    // Do = 64 fills only half of the 128 accumulator rows: the other half is fed with zeros
    StackedRows g{reinterpret_cast<const float4*>(G), nullptr, Do / 4, 0};
    StackedRows x{reinterpret_cast<const float4*>(X), nullptr, Di / 4, 0};
    int grid = 0;
    // the GCN shape (64 x 64) takes the narrow variant: half of operand G is never stored, loads are pipelined
    int rc = (Do == 64 && Di == 64) ? tgrad::launch<1, 2, 2, -1>(device, g, x, ws, M, Di, &grid, stream, "wgrad_tc")
                                    : tgrad::launch(device, g, x, ws, M, Di, &grid, stream, "wgrad_tc");
    if (rc) return rc;
    return tgrad::gather(ws, grid, Di, 0, Do, 0, Di, dW, Di, accumulate, stream);
}

// Pipe-head parameter gradients: dW1 [128, 192], db1 [128], dw2 [128] from the 1-bit gate of the hidden layer, dlogit and
// the node states (see the comment above the loaders).  ws: ltgnn_pipe_head_ws_floats(device) floats.
extern "C" int ltgnn_pipe_head_bwd_w(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                                     const int32_t* ends, const float* W1, const float* b1, const float* w2,
                                     const uint32_t* hmask, const float* dlogit, float gate_scale, float* dW1, float* db1,
                                     float* dw2, float* ws, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P > 0, LTGNN_E_ARG, "pipe_head_bwd_w: B=%lld N=%d P=%d", static_cast<long long>(B), N, P);
    LTGNN_REQUIRE(D == 64 && H == 128, LTGNN_E_SHAPE, "pipe_head_bwd_w: D=%d H=%d (64 / 128 only)", D, H);
    LTGNN_REQUIRE(X && ends && W1 && b1 && w2 && hmask && dlogit && dW1 && db1 && dw2 && ws, LTGNN_E_ARG,
                  "pipe_head_bwd_w: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(hmask) && aligned16(ws), LTGNN_E_ALIGN, "pipe_head_bwd_w: alignment");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B == 0) {
        LTGNN_CUDA_TRY(cudaMemsetAsync(dW1, 0, sizeof(float) * H * 3 * D, stream));
        LTGNN_CUDA_TRY(cudaMemsetAsync(db1, 0, sizeof(float) * H, stream));
        LTGNN_CUDA_TRY(cudaMemsetAsync(dw2, 0, sizeof(float) * H, stream));
        return LTGNN_OK;
    }
    const int64_t M = B * P;
    const DeviceInfo* di = device_info(device);
    LTGNN_REQUIRE(di, LTGNN_E_CUDA, "pipe_head_bwd_w: device %d", device);
    const int No = D;  // one 64-column slice of the 192 feature columns per CTA
    float* part = ws + static_cast<size_t>(di->sm_count) * tgrad::kMo * No;  // [2 sm][128] partials of a[]
    float* A = part + static_cast<size_t>(di->sm_count) * 2 * 128;           // [128][192]
    float* a = A + 128 * 192;                                                // [128]
    HeadLive g{hmask, dlogit, gate_scale, part};
    HeadFeatSlice x{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), dlogit, gate_scale,
                    static_cast<uint32_t>(P), static_cast<uint32_t>(N), P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull};
    int grid = 0;
    int rc = tgrad::launch<1, 4, 2, -1>(device, g, x, ws, M, No, &grid, stream, "pipe_head_bwd_w");
    if (rc) return rc;
    for (int sl = 0; sl < kHeadSlices; ++sl) {  // A[:, 64 sl : 64 sl + 64] = sum over the row ranges of slice sl
        rc = tgrad::gather(ws, grid, No, 0, H, 0, No, A + sl * No, 3 * D, 0, stream, tgrad::kMo, sl, kHeadSlices);
        if (rc) return rc;
    }
    rc = reduce_parts(part, 128, a, 2 * grid, H, 0, stream);
    if (rc) return rc;
    head_param_epilogue_kernel<<<(H + 3) / 4, 128, 0, stream>>>(A, a, W1, b1, w2, dW1, db1, dw2, H, 3 * D);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
