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

// ---- pipe head: G = d loss / d pre (from the saved post-activation), X = [feat | 1 | 0...]
struct HeadDpre {
    static constexpr bool kRowFast = true;  // hpost is stored blocked-32
    const float4* hpost;  // blocked-32 [Mp, 128]
    const float* dlogit;  // [M]
    const float4* w2;     // [32]
    float scale;
    __device__ __forceinline__ float4 operator()(uint32_t row, int c) const {
        const float4 h = ptx::ldg_stream(hpost + ptx::b32(row, c, 32));
        const float4 w = __ldg(w2 + c);
        const float g = __ldg(dlogit + row) * scale;
        return make_float4(h.x > 0.f ? g * w.x : 0.f, h.y > 0.f ? g * w.y : 0.f, h.z > 0.f ? g * w.z : 0.f,
                           h.w > 0.f ? g * w.w : 0.f);
    }
};
struct HeadFeatOnes {
    static constexpr bool kRowFast = false;
    const float4* x;   // node states [B*N, 16]
    const int2* ends;
    uint32_t P, N;
    uint64_t magic;
    __device__ __forceinline__ float4 operator()(uint32_t row, int c) const {
        if (c >= 48) return make_float4(c == 48 ? 1.f : 0.f, 0.f, 0.f, 0.f);  // column 192 = 1 -> db1
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

namespace {
// dw2[j] = sum_rows dlogit[row] * hpost[row, j]: one streaming pass over the blocked-32 activations.
// warp w of a CTA owns float4 columns w, w + 8, w + 16, w + 24; lane = row inside a 32-row block.
__global__ void __launch_bounds__(256)
head_dw2_kernel(const float4* __restrict__ hpost, const float* __restrict__ dlogit, float* __restrict__ part, uint32_t M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t n_blk = (M + 31) / 32;
    for (uint32_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const uint32_t row = blk * 32 + lane;
        const float d = row < M ? __ldg(dlogit + row) : 0.f;
        float4 h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = ptx::ldg_stream(hpost + (static_cast<size_t>(blk) * 32 + warp + 8 * k) * 32 + lane);
        if (row >= M) {  // the padding rows of the last block are never written
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[k].x = fmaf(d, h[k].x, acc[k].x); acc[k].y = fmaf(d, h[k].y, acc[k].y);
            acc[k].z = fmaf(d, h[k].z, acc[k].z); acc[k].w = fmaf(d, h[k].w, acc[k].w);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o); acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
            acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o); acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
        }
        if (lane == 0) reinterpret_cast<float4*>(part + static_cast<size_t>(blockIdx.x) * 128)[warp + 8 * k] = acc[k];
    }
}
}  // namespace

extern "C" int64_t ltgnn_pipe_head_ws_floats(int device) {
    const DeviceInfo* di = device_info(device);  // tgrad partials [sm][128][224] + dw2 partials [8 sm][128]
    return di ? static_cast<int64_t>(di->sm_count) * tgrad::kMo * (224 + 8) : -1;
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
    LTGNN_REQUIRE(Di % 32 == 0 && Di > 0 && Di <= 256, LTGNN_E_SHAPE, "wgrad_tc: Di=%d must be a multiple of 32, <= 256", Di);
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
    int rc = (Do == 64 && Di == 64) ? tgrad::launch<1, 2, 2>(device, g, x, ws, M, Di, &grid, stream, "wgrad_tc")
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
    HeadDpre g{reinterpret_cast<const float4*>(hpost), dlogit, reinterpret_cast<const float4*>(w2), gate_scale};
    HeadFeatOnes x{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), static_cast<uint32_t>(P),
                   static_cast<uint32_t>(N), P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull};
    const int No = 224;  // 192 feature columns + a ones column (-> db1) padded to a whole 32-column block
    int grid = 0;
    int rc = tgrad::launch<1, 4, 7>(device, g, x, ws, M, No, &grid, stream, "pipe_head_bwd_w");
    if (rc) return rc;
    rc = tgrad::gather(ws, grid, No, 0, H, 0, 3 * D, dW1, 3 * D, 0, stream);
    if (rc) return rc;
    rc = tgrad::gather(ws, grid, No, 0, H, 3 * D, 1, db1, 1, 0, stream);
    if (rc) return rc;
    const DeviceInfo* di = device_info(device);
    LTGNN_REQUIRE(di, LTGNN_E_CUDA, "pipe_head_bwd_w: device %d", device);
    float* part = ws + static_cast<size_t>(di->sm_count) * tgrad::kMo * No;  // [8 sm][128]
    const int n_blk = static_cast<int>((M + 31) / 32);
    const int grid2 = n_blk < 8 * di->sm_count ? n_blk : 8 * di->sm_count;
    head_dw2_kernel<<<grid2, 256, 0, stream>>>(reinterpret_cast<const float4*>(hpost), dlogit, part, static_cast<uint32_t>(M));
    LTGNN_CUDA_TRY(cudaGetLastError());
    return reduce_parts(part, H, dw2, grid2, H, 0, stream);
}
