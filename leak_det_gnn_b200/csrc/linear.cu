// linear.cu -- row-wise dense layer  Y = act(X W^T + b)  on tcgen05 tensor cores (3xTF32, fp32-faithful).
//
// Replaces the plain Linear layers of the detector's message-passing path when they are not fused into
// a neighbouring kernel: GCNConv's `self.lin` (PyG Linear, models/detector.py:199), NoLeakHead
// (models/detector.py:94-99).  Also the validation vehicle for the tensor-core building blocks in
// umma.cuh: tests/test_linear_gpu.py checks it against an fp64 matmul.
//
// One CTA (128 threads) owns 128-row tiles (persistent, grid-strided).  W is split into TF32 hi/lo
// once per CTA; each X tile is loaded with coalesced 128-bit loads, split, and stored K-major
// SWIZZLE_128B; one thread issues the tcgen05.mma chain into a TMEM accumulator; the 4 warps read
// their 32 TMEM lanes back (row = lane), add bias / ReLU, and store.
#include "umma.cuh"

using namespace ltgnn;
using namespace ltgnn::ptx;
using namespace ltgnn::umma;

namespace {

constexpr int kTileM = 128;

__global__ void __launch_bounds__(128)
linear_tf32x3_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ Y, int64_t M, int K, int N, int relu, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t a_bytes = kTileM * K * 4, b_bytes = N * K * 4;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k4 = K >> 2;  // 16-byte chunks per row

    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < N * k4; i += 128) {
        const int r = i / k4, c = i - r * k4;
        float4 hi, lo;
        split4(__ldg(reinterpret_cast<const float4*>(W) + i), hi, lo);
        const uint32_t off = sw128_offset(r, c, N);
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_lo + off) = lo;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = idesc_tf32(kTileM, N);
    const int64_t n_tiles = (M + kTileM - 1) / kTileM;
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        for (int i = tid; i < kTileM * k4; i += 128) {
            const int r = i / k4, c = i - r * k4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < M) v = ldg_stream(reinterpret_cast<const float4*>(X) + (row0 + r) * k4 + c);
            float4 hi, lo;
            split4(v, hi, lo);
            const uint32_t off = sw128_offset(r, c, kTileM);
            *reinterpret_cast<float4*>(a_hi + off) = hi;
            *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        fence_before_sync();       // previous tile's TMEM reads ordered before the barrier
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            for (int ka = 0; ka < (K >> 5); ++ka)
                mma_katom_3x(tmem_base, smem_u32(a_hi) + ka * kTileM * 128, smem_u32(a_lo) + ka * kTileM * 128,
                             smem_u32(b_hi) + ka * N * 128, smem_u32(b_lo) + ka * N * 128, idesc, ka == 0);
            commit(&mma_bar);
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        fence_after_sync();

        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const int64_t grow = row0 + warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (bias) v[j] += __ldg(bias + c0 + j);
                if (relu) v[j] = fmaxf(v[j], 0.f);
            }
            if (grow < M) {
                float4* y = reinterpret_cast<float4*>(Y + grow * N + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

extern "C" int ltgnn_linear(int device, int64_t M, int32_t K, int32_t N, const float* X, const float* W,
                            const float* bias, int relu, float* Y, void* stream_) {
    LTGNN_REQUIRE(M >= 0 && K > 0 && N > 0, LTGNN_E_ARG, "linear: M=%lld K=%d N=%d", static_cast<long long>(M), K, N);
    LTGNN_REQUIRE(K % 32 == 0 && K <= 256, LTGNN_E_SHAPE, "linear: K=%d must be a multiple of 32, <= 256", K);
    LTGNN_REQUIRE(N % 16 == 0 && N <= 256, LTGNN_E_SHAPE, "linear: N=%d must be a multiple of 16, <= 256", N);
    if (M == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && W && Y, LTGNN_E_ARG, "linear: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(W) && aligned16(Y), LTGNN_E_ALIGN, "linear: 16-byte alignment required");
    LTGNN_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    LTGNN_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    LTGNN_REQUIRE(prop.major == 10, LTGNN_E_UNSUPPORTED, "linear: device is sm_%d%d, need sm_100", prop.major, prop.minor);
    const size_t smem = 1024 + 2ull * kTileM * K * 4 + 2ull * N * K * 4;
    LTGNN_REQUIRE(smem <= prop.sharedMemPerBlockOptin, LTGNN_E_SHAPE, "linear: K=%d N=%d needs %zu B of shared memory", K,
                  N, smem);
    uint32_t cols = 32;
    while (cols < static_cast<uint32_t>(N)) cols <<= 1;
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    const int per_sm = static_cast<int>(prop.sharedMemPerMultiprocessor / (smem + 1024)) > 0
                           ? static_cast<int>(prop.sharedMemPerMultiprocessor / (smem + 1024)) : 1;
    int64_t grid = static_cast<int64_t>(prop.multiProcessorCount) * per_sm;
    if (grid > tiles) grid = tiles;
    linear_tf32x3_kernel<<<static_cast<int>(grid), 128, smem, static_cast<cudaStream_t>(stream_)>>>(
        X, W, bias, Y, M, K, N, relu, cols);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
