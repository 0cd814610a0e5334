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

// ---- pipe head.  d pre[row, j] = dlogit[row] * scale * w2[j] * live[row, j] is rank one up to the 0/1 mask, so
//      dW1[j, :] = w2[j] * sum_rows live[row, j] * (dlogit[row] * scale * feat[row, :]):
//      G = live (exactly representable: no lo copy, one MMA less per K step), X = F' = dlogit * scale * feat, and the
//      row scale w2[j] is applied when the CTA partials are added.  The 192 result columns are split over three CTAs
//      per row range -- h_u, h_v and |h_u - h_v| -- so that every CTA keeps its 128 x 64 totals in registers
//      (tgrad.cuh, register-total form); the three read the same saved activations at the same time (one from HBM, two
//      from L2).  d w2 and d b1 ride on the slice-0 CTAs' pass over the activations.
constexpr int kHeadSlices = 3;
struct HeadSide {
    float4 dw2, db1;
};
struct HeadLive {
    static constexpr bool kRowFast = true;  // hpost is stored blocked-32
    static constexpr bool kExact = true;
    using Side = HeadSide;
    const float4* hpost;  // blocked-32 [Mp, 128]
    const float* dlogit;  // [M]
    const float4* w2;     // [32]
    float scale;
    float* side_part;     // [2 * gridDim.x][256]: d w2 | d b1, one partial per (CTA, loader group)
    __device__ __forceinline__ float4 load(uint32_t row, int c, Side& side) const {
        const float4 h = ptx::ldg_stream(hpost + ptx::b32(row, c, 32));
        const float4 live = make_float4(h.x > 0.f ? 1.f : 0.f, h.y > 0.f ? 1.f : 0.f, h.z > 0.f ? 1.f : 0.f,
                                        h.w > 0.f ? 1.f : 0.f);
        if (blockIdx.x % kHeadSlices == 0) {
            const float4 w = __ldg(w2 + c);
            const float d = __ldg(dlogit + row), g = d * scale;
            side.dw2.x = fmaf(d, h.x, side.dw2.x); side.dw2.y = fmaf(d, h.y, side.dw2.y);
            side.dw2.z = fmaf(d, h.z, side.dw2.z); side.dw2.w = fmaf(d, h.w, side.dw2.w);
            side.db1.x = fmaf(g * w.x, live.x, side.db1.x); side.db1.y = fmaf(g * w.y, live.y, side.db1.y);
            side.db1.z = fmaf(g * w.z, live.z, side.db1.z); side.db1.w = fmaf(g * w.w, live.w, side.db1.w);
        }
        return live;
    }
    // the 32 lanes of a loader warp hold 32 different rows of the same columns: butterfly sum, lane 0 writes
    __device__ __forceinline__ void finish(Side* side, const int* cols, int n, uint32_t cta, int group, int lane) const {
        for (int j = 0; j < n; ++j) {
            float4 t = side[j].dw2, u = side[j].db1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
                u.x += __shfl_xor_sync(0xffffffffu, u.x, o); u.y += __shfl_xor_sync(0xffffffffu, u.y, o);
                u.z += __shfl_xor_sync(0xffffffffu, u.z, o); u.w += __shfl_xor_sync(0xffffffffu, u.w, o);
            }
            if (lane == 0) {
                float4* dst = reinterpret_cast<float4*>(side_part + (static_cast<size_t>(cta) * 2 + group) * 256);
                dst[cols[j]] = t;
                dst[32 + cols[j]] = u;
            }
        }
    }
};
struct HeadFeatSlice {
    static constexpr bool kRowFast = false;
    static constexpr int kSlices = kHeadSlices;
    const float4* x;   // node states [B*N, 16]
    const int2* ends;
    const float* dlogit;
    float scale;
    uint32_t P, N;
    uint64_t magic;
    __device__ __forceinline__ float4 operator()(uint32_t row, int c) const {  // c < 16: one 64-column slice
        const uint32_t b = magic ? ptx::fastdiv(row, magic) : row;
        const int2 e = __ldg(ends + (row - b * P));
        const float g = __ldg(dlogit + row) * scale;
        const int slice = blockIdx.x % kHeadSlices;
        const float4* xb = x + static_cast<int64_t>(b) * N * 16 + c;
        float4 v;
        if (slice == 0) {
            v = __ldg(xb + e.x * 16);
        } else if (slice == 1) {
            v = __ldg(xb + e.y * 16);
        } else {
            const float4 a = __ldg(xb + e.x * 16), d = __ldg(xb + e.y * 16);
            v = make_float4(fabsf(a.x - d.x), fabsf(a.y - d.y), fabsf(a.z - d.z), fabsf(a.w - d.w));
        }
        return make_float4(g * v.x, g * v.y, g * v.z, g * v.w);
    }
};

}  // namespace

extern "C" int64_t ltgnn_pipe_head_ws_floats(int device) {
    const DeviceInfo* di = device_info(device);  // tgrad partials [sm][128][64] + (d w2 | d b1) partials [2 sm][256]
    return di ? static_cast<int64_t>(di->sm_count) * (tgrad::kMo * 64 + 2 * 256) : -1;
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

// Pipe-head parameter gradients: dW1 [128, 192], db1 [128], dw2 [128].  ws: ltgnn_pipe_head_ws_floats(device) floats.
extern "C" int ltgnn_pipe_head_bwd_w(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                                     const int32_t* ends, const float* w2, const float* hpost, const float* dlogit,
                                     float gate_scale, float* dW1, float* db1, float* dw2, float* ws, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P > 0, LTGNN_E_ARG, "pipe_head_bwd_w: B=%lld N=%d P=%d", static_cast<long long>(B), N, P);
    LTGNN_REQUIRE(D == 64 && H == 128, LTGNN_E_SHAPE, "pipe_head_bwd_w: D=%d H=%d (64 / 128 only)", D, H);
    LTGNN_REQUIRE(X && ends && w2 && hpost && dlogit && dW1 && db1 && dw2 && ws, LTGNN_E_ARG, "pipe_head_bwd_w: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(w2) && aligned16(hpost), LTGNN_E_ALIGN, "pipe_head_bwd_w: alignment");
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
    float* part = ws + static_cast<size_t>(di->sm_count) * tgrad::kMo * No;  // [2 sm][256]
    HeadLive g{reinterpret_cast<const float4*>(hpost), dlogit, reinterpret_cast<const float4*>(w2), gate_scale, part};
    HeadFeatSlice x{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), dlogit, gate_scale,
                    static_cast<uint32_t>(P), static_cast<uint32_t>(N), P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull};
    int grid = 0;
    int rc = tgrad::launch<1, 4, 2, -1>(device, g, x, ws, M, No, &grid, stream, "pipe_head_bwd_w");
    if (rc) return rc;
    for (int sl = 0; sl < kHeadSlices; ++sl) {  // dW1[:, 64 sl : 64 sl + 64] = w2[j] * sum over the row ranges of slice sl
        rc = tgrad::gather(ws, grid, No, 0, H, 0, No, dW1 + sl * No, 3 * D, 0, stream, tgrad::kMo, sl, kHeadSlices, w2);
        if (rc) return rc;
    }
    rc = reduce_parts(part, 256, dw2, 2 * grid, H, 0, stream);
    if (rc) return rc;
    return reduce_parts(part + 128, 256, db1, 2 * grid, H, 0, stream);
}
