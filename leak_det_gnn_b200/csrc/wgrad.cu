// wgrad.cu -- weight gradient of a row-wise linear map:  dW[Do, Di] = sum_rows G[row, :]^T X[row, :]
//
// Replaces autograd's `grad_weight = grad_output.t() @ input` for GCNConv.lin (PyG Linear,
// models/detector.py:199) over the M = B*N node rows of a batch.  It is a tall-skinny reduction
// (M ~ 2.7e6 rows into a 64x64 result) and is HBM-bound: both operands are streamed exactly once.
//
// Each CTA walks a contiguous range of 32-row chunks, double-buffered into shared memory with
// cp.async; 256 threads are split into groups, every thread keeps an 8x8 block of dW in registers
// (64 FMAs per two LDS.128 pairs), groups take alternate rows.  Per-CTA partials go to a workspace
// and a second kernel adds them in a fixed order, so the result is run-to-run deterministic.
#include "common.cuh"

using namespace ltgnn;

namespace {

constexpr int kThreads = 256;
constexpr int kChunkRows = 32;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// G: [M, Do], X: [M, Di], ws: [gridDim.x, Do*Di]
__global__ void __launch_bounds__(kThreads, 2)
wgrad_kernel(const float* __restrict__ G, const float* __restrict__ X, float* __restrict__ ws, int64_t M, int Do, int Di) {
    extern __shared__ __align__(16) float smem_f[];
    const int row_floats = Do + Di;
    float* buf[2] = {smem_f, smem_f + kChunkRows * row_floats};

    const int tid = threadIdx.x;
    const int tj = Di >> 3;                 // thread columns
    const int tpg = (Do >> 3) * tj;         // threads per group
    const int n_groups = kThreads / tpg;
    const int grp = tid / tpg, t = tid - grp * tpg;
    const int o0 = (t / tj) * 8, i0 = (t % tj) * 8;

    const int64_t n_chunks = (M + kChunkRows - 1) / kChunkRows;
    const int64_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const int64_t c_begin = blockIdx.x * per;
    const int64_t c_end = c_begin + per < n_chunks ? c_begin + per : n_chunks;

    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    const int g4 = Do >> 2, x4 = Di >> 2, r4 = g4 + x4;  // float4 per row: G part then X part
    auto issue = [&](int64_t chunk, float* dst) {
        const int64_t row0 = chunk * kChunkRows;
        for (int i = tid; i < kChunkRows * r4; i += kThreads) {
            const int r = i / r4, c = i - r * r4;
            float* d = dst + r * row_floats + c * 4;
            if (row0 + r < M) {
                const float* src = c < g4 ? G + (row0 + r) * Do + c * 4 : X + (row0 + r) * Di + (c - g4) * 4;
                cp_async16(d, src);
            } else {
                *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };

    if (c_begin < c_end) issue(c_begin, buf[0]);
    for (int64_t c = c_begin; c < c_end; ++c) {
        const int cur = static_cast<int>((c - c_begin) & 1);
        if (c + 1 < c_end) {
            issue(c + 1, buf[cur ^ 1]);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* base = buf[cur];
        for (int r = grp; r < kChunkRows; r += n_groups) {
            const float* row = base + r * row_floats;
            const float4 ga = *reinterpret_cast<const float4*>(row + o0);
            const float4 gb = *reinterpret_cast<const float4*>(row + o0 + 4);
            const float4 xa = *reinterpret_cast<const float4*>(row + Do + i0);
            const float4 xb = *reinterpret_cast<const float4*>(row + Do + i0 + 4);
            const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(gv[a], xv[b], acc[a][b]);
        }
        __syncthreads();
    }

    // combine the groups in a fixed order through shared memory, then write this CTA's partial
    float* red = smem_f;  // Do*Di floats <= 2 * kChunkRows * row_floats for every supported shape
    for (int gsel = 0; gsel < n_groups; ++gsel) {
        if (grp == gsel) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float* p = red + (o0 + a) * Di + i0 + b;
                    *p = gsel == 0 ? acc[a][b] : *p + acc[a][b];
                }
        }
        __syncthreads();
    }
    float* out = ws + static_cast<size_t>(blockIdx.x) * Do * Di;
    for (int i = tid; i < Do * Di; i += kThreads) out[i] = red[i];
}

int wgrad_grid(const DeviceInfo* di) { return di->sm_count * 2; }

}  // namespace

extern "C" int64_t ltgnn_wgrad_ws_floats(int device, int32_t Do, int32_t Di) {
    const DeviceInfo* di = device_info(device);
    return di ? static_cast<int64_t>(wgrad_grid(di)) * Do * Di : -1;
}

extern "C" int ltgnn_wgrad(int device, int64_t M, int32_t Do, int32_t Di, const float* G, const float* X, float* dW,
                           int accumulate, float* ws, void* stream_) {
    LTGNN_REQUIRE(M >= 0 && Do > 0 && Di > 0, LTGNN_E_ARG, "wgrad: M=%lld Do=%d Di=%d", static_cast<long long>(M), Do, Di);
    LTGNN_REQUIRE(Do % 8 == 0 && Di % 8 == 0, LTGNN_E_SHAPE, "wgrad: Do=%d, Di=%d must be multiples of 8", Do, Di);
    const int tpg = (Do / 8) * (Di / 8);
    LTGNN_REQUIRE(tpg <= kThreads && kThreads % tpg == 0, LTGNN_E_SHAPE,
                  "wgrad: (Do/8)*(Di/8)=%d must divide %d (e.g. 64x64, 128x128, 64x128)", tpg, kThreads);
    LTGNN_REQUIRE(G && X && dW && ws, LTGNN_E_ARG, "wgrad: null tensor");
    LTGNN_REQUIRE(aligned16(G) && aligned16(X), LTGNN_E_ALIGN, "wgrad: G/X must be 16-byte aligned");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "wgrad: device is sm_%d%d, need sm_100", di->cc_major, di->cc_minor);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    LTGNN_USE_DEVICE(device);
    const int n = Do * Di;
    size_t smem = 2ull * kChunkRows * (Do + Di) * 4;
    if (smem < static_cast<size_t>(n) * 4) smem = static_cast<size_t>(n) * 4;
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "wgrad: needs %zu B of shared memory", smem);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int grid = wgrad_grid(di);
    wgrad_kernel<<<grid, kThreads, smem, stream>>>(G, X, ws, M, Do, Di);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return reduce_parts(ws, n, dW, grid, n, accumulate, stream);
}
