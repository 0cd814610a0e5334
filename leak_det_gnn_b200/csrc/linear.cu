// linear.cu -- row-wise dense layer  Y = gate(act(X op(W) + b))  on tcgen05 tensor cores (3xTF32).
//
// Replaces the plain Linear layers of the detector's message-passing path that are not fused into a
// neighbouring kernel: GCNConv's `self.lin` (PyG Linear, models/detector.py:199) in the forward, its
// input-gradient GEMM dX = dXW * W in the backward (autograd of the same line), NoLeakHead
// (models/detector.py:94-99).  Built on the pipelined skeleton in rowgemm.cuh; also the validation
// vehicle for the tensor-core building blocks (tests/test_linear_gpu.py checks it against fp64).
#include "functors.cuh"
#include "rowgemm.cuh"

using namespace ltgnn;
using namespace ltgnn::functors;

namespace {

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
