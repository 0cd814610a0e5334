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

// ---- pipe head: G = d loss / d pre (from the saved post-activation), X = feat; d w2 and d b1 ride on the same pass.
// (Tried and dropped: G = the 0/1 live mask (exact in TF32: no lo copy), X = dlogit * feat, the 192 result columns split
// over three CTAs per row range so that each keeps its totals in registers -- accurate, but with 6 loads per loader
// thread the kernel became latency bound: 5.4 ms against 1.7 ms for this form.)
struct HeadSide {
    float4 dw2, db1;
};
struct HeadDpre {
    static constexpr bool kRowFast = true;  // hpost is stored blocked-32
    using Side = HeadSide;  // this thread's share of d w2 = sum_rows dlogit * hpost and of d b1 = sum_rows d pre
    const float4* hpost;  // blocked-32 [Mp, 128]
    const float* dlogit;  // [M]
    const float4* w2;     // [32]
    float scale;
    float* side_part;     // [2 * gridDim.x][256]: d w2 | d b1, one partial per (CTA, loader group)
    __device__ __forceinline__ float4 load(uint32_t row, int c, Side& side) const {
        const float4 h = ptx::ldg_stream(hpost + ptx::b32(row, c, 32));
        const float4 w = __ldg(w2 + c);
        const float d = __ldg(dlogit + row);
        side.dw2.x = fmaf(d, h.x, side.dw2.x); side.dw2.y = fmaf(d, h.y, side.dw2.y);
        side.dw2.z = fmaf(d, h.z, side.dw2.z); side.dw2.w = fmaf(d, h.w, side.dw2.w);
        const float g = d * scale;
        const float4 dp = make_float4(h.x > 0.f ? g * w.x : 0.f, h.y > 0.f ? g * w.y : 0.f, h.z > 0.f ? g * w.z : 0.f,
                                      h.w > 0.f ? g * w.w : 0.f);
        side.db1.x += dp.x; side.db1.y += dp.y; side.db1.z += dp.z; side.db1.w += dp.w;
        return dp;
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
struct HeadFeat {
    static constexpr bool kRowFast = false;
    const float4* x;   // node states [B*N, 16]
    const int2* ends;
    uint32_t P, N;
    uint64_t magic;
    __device__ __forceinline__ float4 operator()(uint32_t row, int c) const {
        const uint32_t b = magic ? ptx::fastdiv(row, magic) : row;
        const int2 e = __ldg(ends + (row - b * P));
        const int seg = c >> 4, cc = c & 15;
        const float4* xb = x + static_cast<int64_t>(b) * N * 16 + cc;
        if (seg == 0) return __ldg(xb + e.x * 16);
        if (seg == 1) return __ldg(xb + e.y * 16);
        const float4 a = __ldg(xb + e.x * 16), d = __ldg(xb + e.y * 16);
        return make_float4(fabsf(a.x - d.x), fabsf(a.y - d.y), fabsf(a.z - d.z), fabsf(a.w - d.w));
    }
};

}  // namespace

extern "C" int64_t ltgnn_pipe_head_ws_floats(int device) {
    const DeviceInfo* di = device_info(device);  // tgrad partials [sm][128][192] + (d w2 | d b1) partials [2 sm][256]
    return di ? static_cast<int64_t>(di->sm_count) * (tgrad::kMo * 192 + 2 * 256) : -1;
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
    const int No = 3 * D;  // the 192 feature columns; d b1 and d w2 are accumulated by the loaders on the side
    float* part = ws + static_cast<size_t>(di->sm_count) * tgrad::kMo * No;  // [2 sm][256]
    HeadDpre g{reinterpret_cast<const float4*>(hpost), dlogit, reinterpret_cast<const float4*>(w2), gate_scale, part};
    HeadFeat x{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), static_cast<uint32_t>(P),
               static_cast<uint32_t>(N), P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull};
    int grid = 0;
    int rc = tgrad::launch<1, 4, 6>(device, g, x, ws, M, No, &grid, stream, "pipe_head_bwd_w");
    if (rc) return rc;
    rc = tgrad::gather(ws, grid, No, 0, H, 0, No, dW1, No, 0, stream);
    if (rc) return rc;
    rc = reduce_parts(part, 256, dw2, 2 * grid, H, 0, stream);
    if (rc) return rc;
    return reduce_parts(part + 128, 256, db1, 2 * grid, H, 0, stream);
}
