// rowgemm.cuh -- warp-specialised, pipelined row-tile GEMM skeleton on tcgen05 (3xTF32).
//
//   D[128 x N] = A_tile[128 x K] * B[N x K]^T      per 128-row tile, persistent CTAs, one per SM
//
//   warps 0-15  LOADERS  produce the A tile 32 K-columns at a time: a `Loader` functor returns 4 consecutive fp32
//                        of a row (a plain coalesced global load, a pipe-end gather, an on-the-fly gradient...);
//                        values are split into TF32 hi/lo and stored K-major SWIZZLE_128B into a 4-stage ring,
//                        one group of loader warps per stage
//   warp  16    MMA      one elected lane issues the tcgen05.mma chain into one of two TMEM accumulators and
//                        commits to the "stage free" and "accumulator full" mbarriers
//   warps 17-20 EPILOGUE tcgen05.ld their 32 TMEM lanes (lane = tile row) 16 columns at a time inside an
//                        `Epilogue` functor (bias, activation, masks, reductions, stores), release the accumulator;
//                        each warp owns a transposition patch (patch.cuh) so that its stores can be coalesced
//
// B (the weight, N x K) is split once per CTA and stays resident in shared memory.
// All hand-offs are mbarriers; no __syncthreads after the prologue.
#pragma once

#include "patch.cuh"
#include "umma.cuh"

namespace ltgnn {
namespace rowgemm {

using namespace ltgnn::ptx;
using namespace ltgnn::umma;

constexpr int kTileM = 128;
constexpr int kLoaderWarps = 16;
constexpr int kMmaWarp = kLoaderWarps;
constexpr int kEpiWarp0 = kMmaWarp + 1;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (kLoaderWarps + 1 + kEpiWarps) * 32;
constexpr int kMaxStages = 4;
constexpr int kKG = 32;                              // K columns per ring stage = one 128-byte swizzle atom
constexpr uint32_t kStageBytes = 2u * kTileM * kKG * 4;  // hi block + lo block = 32 KB

// The A ring has 4 stages (2 when the resident B leaves no room).  Each stage is owned by its own group of
// loader warps, so the global loads / gathers of up to 4 stages are in flight at once -- the loaders are
// latency-bound, and this is what keeps the tensor pipe and HBM busy.
__host__ inline size_t smem_bytes(int K, int N, int stages) {
    return 1024 + static_cast<size_t>(stages) * kStageBytes + 2ull * N * K * 4 + kEpiWarps * patch::kPatchBytes;
}
__host__ inline int pick_stages(int K, int N, size_t limit, int prefer = 4) {
    if (prefer == 3 && smem_bytes(K, N, 3) <= limit) return 3;
    if (smem_bytes(K, N, 4) <= limit) return 4;
    if (smem_bytes(K, N, 2) <= limit) return 2;
    return 0;
}
__host__ inline uint32_t tmem_cols_for(int N) {
    uint32_t c = 32;
    while (c < 2u * N) c <<= 1;
    return c;
}

// Where the resident B operand comes from.  A launch may carry several B "variants" (disjoint slices of one
// weight): CTA c works with variant c % nvar for its whole life and walks row tiles c / nvar, c / nvar + grid / nvar, ...
//   transposed = 0:  B_var[n][k] = W[(var * N + n) * ldw + k]      (rows of a torch.nn.Linear weight [N_total, K])
//   transposed = 1:  B_var[n][k] = W[k * ldw + var * N + n]        (columns of a [K, N_total] matrix: A * W)
struct BSpec {
    const float* W;
    int ldw;
    int transposed;
    int nvar;
};

__device__ inline void fill_b(uint8_t* b_hi, uint8_t* b_lo, const BSpec& bs, int var, int K, int N, int tid,
                              int nthreads) {
    if (!bs.transposed) {
        const int k4 = K >> 2;
        for (int i = tid; i < N * k4; i += nthreads) {
            const int r = i / k4, c = i - r * k4;
            float4 hi, lo;
            split4(__ldg(reinterpret_cast<const float4*>(bs.W + static_cast<size_t>(var * N + r) * bs.ldw) + c), hi, lo);
            const uint32_t off = sw128_offset(r, c, N);
            *reinterpret_cast<float4*>(b_hi + off) = hi;
            *reinterpret_cast<float4*>(b_lo + off) = lo;
        }
    } else {
        for (int i = tid; i < N * K; i += nthreads) {
            const int k = i / N, n = i - k * N;  // coalesced along n
            const float w = __ldg(bs.W + static_cast<size_t>(k) * bs.ldw + var * N + n);
            const float hi = tf32_hi(w), lo = w - hi;
            const uint32_t off = sw128_offset(n, k >> 2, N) + (k & 3) * 4;
            *reinterpret_cast<float*>(b_hi + off) = hi;
            *reinterpret_cast<float*>(b_lo + off) = lo;
        }
    }
}

template <class Loader, class Epilogue>
__global__ void __launch_bounds__(kThreads, 1)
rowgemm_kernel(const Loader loader, const Epilogue epilogue, const BSpec bspec, uint32_t M, int K, int N, int n_stages,
               uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n_kg = K / kKG;
    constexpr uint32_t a_half = kTileM * kKG * 4;  // hi block, then lo block
    uint8_t* b_hi = smem + n_stages * kStageBytes;
    uint8_t* b_lo = b_hi + N * K * 4;
    uint8_t* scratch = b_lo + N * K * 4;  // one transposition patch per epilogue warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warps_per_group = n_stages >= 3 ? 4 : 8;  // 3 stages: warps 12-15 idle

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&bar_full[s], warps_per_group);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], kEpiWarps);
        }
        fence_mbar_init();
    }
    const int var = blockIdx.x % bspec.nvar;
    const uint32_t tile0 = blockIdx.x / bspec.nvar, tile_step = gridDim.x / bspec.nvar;
    fill_b(b_hi, b_lo, bspec, var, K, N, tid, kThreads);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t n_tiles = (M + kTileM - 1) / kTileM;

    if (warp < warps_per_group * n_stages) {
        // ------------------------------- loaders -------------------------------
        // group g owns ring stage g and fills it for every (tile, k-group) step `it` with it % n_stages == g
        const int grp = warp / warps_per_group;
        const int gtid = tid - grp * warps_per_group * 32;
        const int gthreads = warps_per_group * 32;
        const int per_thread = (kTileM * 8) / gthreads;  // 16-byte chunks per thread per stage: 8 or 4
        // a thread always owns 16-byte column c of rows r0, r0 + rstep, ...; rstep is a multiple of 8, so the
        // swizzled offset only advances by whole 8-row groups
        const int r0 = gtid >> 3, c = gtid & 7, rstep = gthreads >> 3;
        const uint32_t off0 = sw128_offset(r0, c, kTileM), offstep = static_cast<uint32_t>(rstep >> 3) * 1024u;
        uint8_t* a_hi = smem + grp * kStageBytes;
        uint8_t* a_lo = a_hi + a_half;
        uint32_t it = 0, use = 0;
        for (uint32_t tile = tile0; tile < n_tiles; tile += tile_step) {
            const uint32_t row0 = tile * kTileM + r0;
            for (int kg = 0; kg < n_kg; ++kg, ++it) {
                if (static_cast<int>(it % n_stages) != grp) continue;
                mbar_wait(&bar_empty[grp], (use & 1) ^ 1);
                ++use;
                // every load of this thread is issued before the first is consumed
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < per_thread && row0 + j * rstep < M) v[j] = loader(row0 + j * rstep, kg * 8 + c);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < per_thread) {
                        float4 hi, lo;
                        split4(v[j], hi, lo);
                        *reinterpret_cast<float4*>(a_hi + off0 + j * offstep) = hi;
                        *reinterpret_cast<float4*>(a_lo + off0 + j * offstep) = lo;
                    }
                }
                fence_proxy_async_smem();  // this thread's smem writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[grp]);
            }
        }
    } else if (warp < kLoaderWarps) {
        // spare loader warps (3-stage configuration)
    } else if (warp == kMmaWarp) {
        // ------------------------------- MMA issuer -------------------------------
        // The whole warp walks the loop (warp-uniform control flow keeps addresses and descriptors in uniform
        // registers); one elected lane issues the tcgen05 instructions.
        const uint32_t idesc = idesc_tf32(kTileM, N);
        const uint32_t a_base = smem_u32(smem), bh_base = smem_u32(b_hi), bl_base = smem_u32(b_lo);
        uint32_t it = 0, t = 0;
        for (uint32_t tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
            const uint32_t a = t & 1;
            mbar_wait(&bar_tempty[a], ((t >> 1) & 1) ^ 1);  // epilogue drained this accumulator
            const uint32_t d = tmem_base + a * N;
            for (int kg = 0; kg < n_kg; ++kg, ++it) {
                const uint32_t s = it % n_stages;
                mbar_wait(&bar_full[s], (it / n_stages) & 1);  // operands landed
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = a_base + s * kStageBytes;
                    mma_katom_3x(d, a_hi, a_hi + a_half, bh_base + kg * N * 128, bl_base + kg * N * 128, idesc, kg == 0);
                    commit(&bar_empty[s]);                    // smem stage reusable once these MMAs retire
                    if (kg == n_kg - 1) commit(&bar_tfull[a]);  // accumulator ready for the epilogue
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------- epilogue -------------------------------
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const patch::Patch pt(scratch + (warp - kEpiWarp0) * patch::kPatchBytes, lane);
        uint32_t t = 0;
        for (uint32_t tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
            const int a = t & 1;
            const uint32_t pha = (t >> 1) & 1;
            mbar_wait(&bar_tfull[a], pha);
            fence_after_sync();
            const uint32_t taddr = tmem_base + a * N + (static_cast<uint32_t>(q * 32) << 16);
            const uint32_t row = tile * kTileM + q * 32 + lane;
            // the epilogue pulls 16-column chunks itself (tcgen05.ld is warp-collective: every lane must pull
            // every chunk, valid row or not) and may keep per-row state across chunks
            epilogue(row, M, var, [&](int c0, float* v, int w) { if (w == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v); }, pt, lane);
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[a]);
        }
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after_sync();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// host-side launch helper: validates shapes against the device, sets the smem attribute, launches
template <class Loader, class Epilogue>
int launch(int device, const Loader& loader, const Epilogue& epilogue, const BSpec& bspec, int64_t M, int K, int N,
           cudaStream_t stream, const char* who, int prefer_stages = 4) {
    LTGNN_REQUIRE(K % 32 == 0 && K > 0 && K <= 256, LTGNN_E_SHAPE, "%s: K=%d must be a multiple of 32, <= 256", who, K);
    LTGNN_REQUIRE(N % 16 == 0 && N > 0 && N <= 256, LTGNN_E_SHAPE, "%s: N=%d must be a multiple of 16, <= 256", who, N);
    LTGNN_REQUIRE(bspec.nvar >= 1, LTGNN_E_ARG, "%s: nvar=%d", who, bspec.nvar);
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "%s: device is sm_%d%d, need sm_100", who, di->cc_major,
                  di->cc_minor);
    const int stages = pick_stages(K, N, static_cast<size_t>(di->smem_optin), prefer_stages);
    LTGNN_REQUIRE(stages > 0, LTGNN_E_SHAPE, "%s: K=%d N=%d needs %zu B of shared memory (limit %d)", who, K, N,
                  smem_bytes(K, N, 2), di->smem_optin);
    LTGNN_REQUIRE(M < (1ll << 31) - kTileM, LTGNN_E_SHAPE, "%s: M=%lld rows exceed the 32-bit row index", who,
                  static_cast<long long>(M));
    const size_t smem = smem_bytes(K, N, stages);
    LTGNN_USE_DEVICE(device);
    auto kern = rowgemm_kernel<Loader, Epilogue>;
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    int64_t grid = tiles * bspec.nvar < di->sm_count ? tiles * bspec.nvar : di->sm_count;
    grid -= grid % bspec.nvar;
    LTGNN_REQUIRE(grid > 0, LTGNN_E_SHAPE, "%s: nvar=%d exceeds the SM count", who, bspec.nvar);
    kern<<<static_cast<int>(grid), kThreads, smem, stream>>>(loader, epilogue, bspec, static_cast<uint32_t>(M), K, N,
                                                            stages, tmem_cols_for(N));
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

}  // namespace rowgemm
}  // namespace ltgnn
