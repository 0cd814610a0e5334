// rowgemm_ts.cuh -- row-tile GEMM with the A operand in TENSOR MEMORY (tcgen05 "TS" form), 3xTF32.
//
//   D[128 x N] = A_tile[128 x K] * B[N x K]^T       persistent CTAs, one per SM
//
// Why: in the shared-memory ("SS") skeleton of rowgemm.cuh every K = 8 step re-reads a 128 x 32 B slice of A
// three times (hi/lo products) and the loaders first have to write both copies -- at N = 64 that traffic, not
// the tensor pipe, is the limit, and the A ring eats the shared memory the whole weight would need.  Here the
// loaders keep a row per thread (thread = TMEM lane), split it in registers and `tcgen05.st` hi and lo straight
// into tensor memory; the MMA reads A from TMEM and only B from shared memory.  That frees enough shared memory
// to keep the WHOLE weight resident (N up to 192 at K = 128, N = 128 at K = 192), so a tile is produced once.
//
//   warps 0-15  LOADERS  groups of 4 warps (= 128 TMEM lanes); group g owns A stage g.  A `Loader` functor
//                        fills 32 consecutive K values of one row (gathers, on-the-fly gradients, ...)
//   warp  16    MMA      elected lane: per stage 4 K-steps x 3 tcgen05.mma (A from TMEM, B descriptor)
//   warps 17-20 EPILOGUE tcgen05.ld + functor, double-buffered accumulators
//
// TMEM budget (512 columns): 2 accumulators of N columns + stages of 64 columns (32 hi + 32 lo).
#pragma once

#include "umma.cuh"

namespace ltgnn {
namespace rowgemm_ts {

using namespace ltgnn::ptx;
using namespace ltgnn::umma;

constexpr int kTileM = 128;
constexpr int kLoaderWarps = 16;
constexpr int kMmaWarp = kLoaderWarps;
constexpr int kThreads = (kLoaderWarps + 1 + 4) * 32;
constexpr int kMaxStages = 4;
constexpr int kStageCols = 64;  // 32 hi + 32 lo columns of A per stage

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

__host__ inline size_t smem_bytes(int K, int N) { return 1024 + 2ull * N * K * 4; }
__host__ inline int stages_for(int N) {
    int s = (512 - 2 * N) / kStageCols;
    return s > kMaxStages ? kMaxStages : s;
}

// Loader:   void operator()(uint32_t row, int kg, float (&v)[32]) const   -- K values [32 kg, 32 kg + 32) of `row`
// Epilogue: void operator()(uint32_t row, bool valid, int var, Pull&& pull) const   -- as in rowgemm.cuh (var = 0)
template <class Loader, class Epilogue>
__global__ void __launch_bounds__(kThreads, 1)
rowgemm_ts_kernel(const Loader loader, const Epilogue epilogue, const float* __restrict__ W, int ldw, int transposed,
                  uint32_t M, int K, int N, int n_stages) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_hi = smem;
    uint8_t* b_lo = b_hi + N * K * 4;
    const int n_kg = K >> 5;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&bar_full[s], 4);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], 4);
        }
        fence_mbar_init();
    }
    fill_b(b_hi, b_lo, W, ldw, transposed, K, N, tid, kThreads);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t acc_base = tmem_base;                 // 2 accumulators of N columns
    const uint32_t a_base = tmem_base + 2 * N;           // then the A stages
    const uint32_t n_tiles = (M + kTileM - 1) / kTileM;

    if (warp < 4 * n_stages) {
        // ------------------------------- loaders: one row (TMEM lane) per thread -------------------------------
        const int grp = warp >> 2;
        const int r_in_tile = (warp & 3) * 32 + lane;
        const uint32_t st_addr = a_base + grp * kStageCols + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t it = 0, use = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t row = tile * kTileM + r_in_tile;
            for (int kg = 0; kg < n_kg; ++kg, ++it) {
                if (static_cast<int>(it % n_stages) != grp) continue;
                float v[32];
                if (row < M) {
                    loader(row, kg, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                float lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = tf32_hi(v[j]);
                    lo[j] = v[j] - h;
                    v[j] = h;
                }
                mbar_wait(&bar_empty[grp], (use & 1) ^ 1);  // stage consumed by the tensor core
                ++use;
                fence_after_sync();
                tmem_st32(st_addr, v);
                tmem_st32(st_addr + 32, lo);
                tmem_wait_st();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[grp]);
            }
        }
    } else if (warp < kLoaderWarps) {
        // spare warps when fewer than 4 stages fit in tensor memory
    } else if (warp == kMmaWarp) {
        // ------------------------------- MMA issuer -------------------------------
        const uint32_t idesc = idesc_tf32(kTileM, N);
        const uint32_t bh = desc_lo(smem_u32(b_hi)), bl = desc_lo(smem_u32(b_lo));
        const uint32_t kg_units = static_cast<uint32_t>(N) * 128u >> 4;  // one 32-column k-atom block of B, in 16 B units
        uint32_t it = 0, t = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t a = t & 1;
            mbar_wait(&bar_tempty[a], ((t >> 1) & 1) ^ 1);
            const uint32_t d = acc_base + a * N;
            for (int kg = 0; kg < n_kg; ++kg, ++it) {
                const uint32_t s = it % n_stages;
                mbar_wait(&bar_full[s], (it / n_stages) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = a_base + s * kStageCols, a_lo = a_hi + 32;
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t boff = kg * kg_units + 2 * k;
                        mma_tf32_ts(d, a_lo + 8 * k, bh + boff, idesc, (kg == 0 && k == 0) ? 0u : 1u);
                        mma_tf32_ts(d, a_hi + 8 * k, bl + boff, idesc, 1u);
                        mma_tf32_ts(d, a_hi + 8 * k, bh + boff, idesc, 1u);
                    }
                    commit(&bar_empty[s]);
                    if (kg == n_kg - 1) commit(&bar_tfull[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3;
        uint32_t t = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const uint32_t a = t & 1;
            mbar_wait(&bar_tfull[a], (t >> 1) & 1);
            fence_after_sync();
            const uint32_t taddr = acc_base + a * N + (static_cast<uint32_t>(q * 32) << 16);
            const uint32_t row = tile * kTileM + q * 32 + lane;
            epilogue(row, row < M, 0, [&](int c0, float (&v)[16]) { tmem_ld16(taddr + c0, v); });
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[a]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

template <class Loader, class Epilogue>
int launch(int device, const Loader& loader, const Epilogue& epilogue, const float* W, int ldw, int transposed, int64_t M,
           int K, int N, cudaStream_t stream, const char* who) {
    LTGNN_REQUIRE(K % 32 == 0 && K > 0 && K <= 256, LTGNN_E_SHAPE, "%s: K=%d must be a multiple of 32, <= 256", who, K);
    LTGNN_REQUIRE(N % 16 == 0 && N > 0 && N <= 192, LTGNN_E_SHAPE, "%s: N=%d must be a multiple of 16, <= 192", who, N);
    LTGNN_REQUIRE(M > 0 && M < (1ll << 31) - kTileM, LTGNN_E_SHAPE, "%s: M=%lld", who, static_cast<long long>(M));
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "%s: device is sm_%d%d, need sm_100", who, di->cc_major,
                  di->cc_minor);
    const size_t smem = smem_bytes(K, N);
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "%s: K=%d N=%d needs %zu B of shared memory",
                  who, K, N, smem);
    const int stages = stages_for(N);
    LTGNN_REQUIRE(stages >= 2, LTGNN_E_SHAPE, "%s: N=%d leaves no tensor memory for the A stages", who, N);
    LTGNN_CUDA_TRY(cudaSetDevice(device));
    auto kern = rowgemm_ts_kernel<Loader, Epilogue>;
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    const int grid = static_cast<int>(tiles < di->sm_count ? tiles : di->sm_count);
    kern<<<grid, kThreads, smem, stream>>>(loader, epilogue, W, ldw, transposed, static_cast<uint32_t>(M), K, N, stages);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

}  // namespace rowgemm_ts
}  // namespace ltgnn
