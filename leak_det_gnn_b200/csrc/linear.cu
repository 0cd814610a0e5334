// linear.cu -- row-wise dense layer  Y = gate(act(X op(W) + b))  on tcgen05 tensor cores (3xTF32).
//
// Replaces the plain Linear layers of the detector's message-passing path that are not fused into a
// neighbouring kernel: GCNConv's `self.lin` (PyG Linear, models/detector.py:199) in the forward, its
// input-gradient GEMM dX = dXW * W in the backward (autograd of the same line), NoLeakHead
// (models/detector.py:94-99).  Built on the pipelined skeleton in rowgemm.cuh; also the validation
// vehicle for the tensor-core building blocks (tests/test_linear_gpu.py checks it against fp64).
#include "rowgemm.cuh"

using namespace ltgnn;

namespace {

struct RowLoader {
    const float4* x;
    int k4;
    __device__ __forceinline__ float4 operator()(uint32_t row, int c) const {
        return ptx::ldg_stream(x + static_cast<int64_t>(row) * k4 + c);
    }
};

// bias -> ReLU -> optional gate: y *= (gate[row, col] > 0) ? gate_scale : 0
// (the gate is the output of an upstream ReLU(+dropout): its positivity IS that layer's backward mask)
struct StoreEpilogue {
    float* y;
    const float* bias;
    const float* gate;
    float gate_scale;
    int n;
    int relu;
    template <class Pull>
    __device__ __forceinline__ void operator()(uint32_t row, bool valid, int /*var*/, Pull&& pull) const {
        for (int c0 = 0; c0 < n; c0 += 16) {
            float v[16];
            pull(c0, v);
            if (valid) chunk(row, c0, v);
        }
    }
    __device__ __forceinline__ void chunk(int64_t row, int c0, float (&v)[16]) const {
        if (bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(bias + c0 + j);
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (gate) {
            const float4* g = reinterpret_cast<const float4*>(gate + row * n + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 m = ptx::ldg_stream(g + j);
                v[4 * j + 0] = m.x > 0.f ? v[4 * j + 0] * gate_scale : 0.f;
                v[4 * j + 1] = m.y > 0.f ? v[4 * j + 1] * gate_scale : 0.f;
                v[4 * j + 2] = m.z > 0.f ? v[4 * j + 2] * gate_scale : 0.f;
                v[4 * j + 3] = m.w > 0.f ? v[4 * j + 3] * gate_scale : 0.f;
            }
        }
        float4* out = reinterpret_cast<float4*>(y + row * n + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
};

}  // namespace

extern "C" int ltgnn_linear(int device, int64_t M, int32_t K, int32_t N, const float* X, const float* W,
                            int w_transposed, const float* bias, int relu, const float* gate, float gate_scale, float* Y,
                            void* stream_) {
    LTGNN_REQUIRE(M >= 0 && K > 0 && N > 0, LTGNN_E_ARG, "linear: M=%lld K=%d N=%d", static_cast<long long>(M), K, N);
    if (M == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && W && Y, LTGNN_E_ARG, "linear: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(W) && aligned16(Y) && aligned16(gate), LTGNN_E_ALIGN,
                  "linear: 16-byte alignment required");
    RowLoader ld{reinterpret_cast<const float4*>(X), K / 4};
    StoreEpilogue ep{Y, bias, gate, gate_scale, N, relu};
    rowgemm::BSpec bs{W, w_transposed ? N : K, w_transposed, 1};
    return rowgemm::launch(device, ld, ep, bs, M, K, N, static_cast<cudaStream_t>(stream_), "linear");
}
