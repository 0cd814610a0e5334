// tgrad.cuh -- weight gradients on tensor cores:  D[Mo x No] = sum over rows of  G[row, :Mo]^T  X[row, :No]
//
// The reduction dimension of this GEMM is the ROW index, so both operands are "MN-major" for the tensor
// core.  For 32-bit (TF32) MN-major operands tcgen05 accepts exactly one shared-memory layout,
// SWIZZLE_128B_BASE32B: blocks of [32 columns][rows][128 B] in which the 32-BYTE chunk c of row r sits at
// chunk c ^ (r & 3) (atoms of 4 rows x 128 B).  Descriptor: leading-dimension offset = one 32-column block,
// stride offset = one group of 4 rows (512 B).  One tcgen05.mma (kind::tf32, K = 8) consumes 8 rows;
// 3xTF32 as everywhere.
//
// The accumulator (Mo = 128 lanes x No <= 256 columns) stays in tensor memory for the whole life of the CTA;
// at the end the CTA writes its partial to a workspace and a second kernel adds the partials in a fixed order.
//   warps 0-15  loaders, two groups of 8 warps, one ring stage each (rows are produced by functors, so the
//               "matrices" can be gathers or on-the-fly gradients and never exist in memory)
//   warp  16    MMA issue (elected lane)
//   warps 17-20 final read-out
#pragma once

#include <type_traits>

#include "umma.cuh"

namespace ltgnn {
namespace tgrad {

using namespace ltgnn::ptx;
using namespace ltgnn::umma;

constexpr int kMo = 128;        // accumulator rows (operand-G columns); pad / stack operands up to it
constexpr int kChunk = 32;      // rows per ring stage (4 K-steps)
constexpr int kLoaderWarps = 16;
constexpr int kGroups = 2;
constexpr int kGroupThreads = kLoaderWarps / kGroups * 32;  // 256
constexpr int kMmaWarp = kLoaderWarps;
constexpr int kThreads = (kLoaderWarps + 1 + 4) * 32;
constexpr uint32_t kBlockBytes = kChunk * 128;  // one 32-column block of a 32-row chunk = 4 KB

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = distance between 32-column blocks, SBO = 4 rows = 512 B
__device__ __forceinline__ uint32_t mn_desc_lo(uint32_t saddr) {
    return ((saddr & 0x3FFFFu) >> 4) | ((kBlockBytes >> 4) << 16);
}
constexpr uint32_t kMnDescHi = (32u /*SBO = 512 B*/) | (1u << 14) /*version*/ | (1u << 29) /*SWIZZLE_128B_BASE32B*/;
// byte offset of 16-byte chunk c16 of row r in a `rows`-row tile of this layout
__device__ __forceinline__ uint32_t mn_offset(int r, int c16, int rows) {
    const int blk = c16 >> 3, c = c16 & 7;
    return static_cast<uint32_t>(blk * rows * 128 + r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}
__device__ __forceinline__ void mma_tf32_mn(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kMnDescHi)
        : "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

__host__ inline size_t stage_bytes(int No, int mt) { return 2ull * (kMo * mt + No) * kChunk * 4; }
__host__ inline size_t smem_bytes(int No, int mt) { return 1024 + kGroups * stage_bytes(No, mt); }

// GLoader: float4 operator()(uint32_t row, int c16) for c16 < kMo/4;  XLoader: same for c16 < No/4.
// Both are only called for row < M.  ws: [gridDim.x][kMo][No].
// kMT = 1 or 2 accumulator tiles of 128 rows: operand G has 128 * kMT columns and shares one pass over X.
// kGJ / kXJ: 16-byte chunks a loader thread fetches per row of G / X (chunks q, q + 8, ...).  The default covers
// every operand; a narrow operand (kGJ < 4 kMT: G columns beyond 32 kGJ are zero and are written once at start;
// kXJ = No / 32) needs few registers per stage, and then the loads of the NEXT chunk are issued before the current
// one is stored (kPipe) -- twice the bytes in flight for the skinny, HBM-bound weight gradients.
template <class GLoader, class XLoader, int kMT, int kGJ = 4 * kMT, int kXJ = 8>
__global__ void __launch_bounds__(kThreads, 1)
tgrad_kernel(const GLoader gload, const XLoader xload, float* __restrict__ ws, uint32_t M, int No, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kGroups], bar_empty[kGroups], bar_done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kG = kMo * kMT;                        // operand-G columns
    const uint32_t g_half = kG * kChunk * 4;             // G hi (then G lo)
    const uint32_t x_half = static_cast<uint32_t>(No) * kChunk * 4;
    const uint32_t stage = 2 * (g_half + x_half);

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < kGroups; ++s) {
            mbar_init(&bar_full[s], kLoaderWarps / kGroups);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_done, 1);
        fence_mbar_init();
    }
    if (kGJ < 4 * kMT) {  // operand-G columns the loaders never write: zero in every stage, hi and lo
        for (uint32_t i = tid; i < kGroups * 2 * (g_half / 16); i += kThreads) {
            const uint32_t st = i / (2 * (g_half / 16)), r = i - st * 2 * (g_half / 16);
            *reinterpret_cast<float4*>(smem + st * stage + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    // contiguous range of row chunks per CTA
    const uint32_t n_chunks = (M + kChunk - 1) / kChunk;
    const uint32_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const uint32_t c_begin = blockIdx.x * per;
    const uint32_t c_end = c_begin + per < n_chunks ? c_begin + per : n_chunks;

    if (warp < kLoaderWarps) {
        const int grp = warp / (kLoaderWarps / kGroups);
        const int gtid = tid - grp * kGroupThreads;
        uint8_t* g_hi = smem + grp * stage;
        uint8_t* g_lo = g_hi + g_half;
        uint8_t* x_hi = g_lo + g_half;
        uint8_t* x_lo = x_hi + x_half;
        // 8 threads per row: thread (r, q) owns 16-byte chunks q, q + 8, ... of row r of an operand, so the
        // row-dependent part of a gather (division, end-node lookup) is done once and all loads go out together
        // kRowFast sources (blocked-32 tensors: 32 consecutive rows of one chunk are contiguous) flip the mapping:
        // a warp = 32 rows x one chunk (512 contiguous bytes) at the price of 2-way conflicts on the STS.  The two
        // operands choose independently.
        const int rg = GLoader::kRowFast ? (gtid & 31) : (gtid >> 3), qg = GLoader::kRowFast ? (gtid >> 5) : (gtid & 7);
        const int rx = XLoader::kRowFast ? (gtid & 31) : (gtid >> 3), qx = XLoader::kRowFast ? (gtid >> 5) : (gtid & 7);
        const int x4 = No / 4;
        constexpr bool kPipe = (kGJ + kXJ) <= 6;
        // chunk j of a thread: q, q + 8, ... for the 8-threads-per-row mapping; for kRowFast sources the warp owns
        // PAIRS of neighbouring chunks (2 q, 2 q + 1, 2 q + 16, ...) so that the store below can be conflict-free
        static_assert(!GLoader::kRowFast || kGJ % 2 == 0, "kRowFast operands need an even chunk count");
        static_assert(!XLoader::kRowFast || kXJ % 2 == 0, "kRowFast operands need an even chunk count");
        auto cg = [&](int j) { return GLoader::kRowFast ? 2 * qg + (j & 1) + 16 * (j >> 1) : qg + 8 * j; };
        auto cx = [&](int j) { return XLoader::kRowFast ? 2 * qx + (j & 1) + 16 * (j >> 1) : qx + 8 * j; };
        float4 gv[kGJ], xv[kXJ];
        auto fetch = [&](uint32_t ch, float4 (&gd)[kGJ], float4 (&xd)[kXJ]) {
            const uint32_t row_g = ch * kChunk + rg, row_x = ch * kChunk + rx;
#pragma unroll
            for (int j = 0; j < kGJ; ++j) gd[j] = row_g < M ? gload(row_g, cg(j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kXJ; ++j)
                xd[j] = (row_x < M && cx(j) < x4) ? xload(row_x, cx(j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // store one operand's chunks (hi and lo).  kRowFast: the 32 lanes of a warp hold 32 rows of the SAME chunk,
        // whose swizzled offsets share only 4 bank groups (2-way conflicts on every quarter-warp).  Lanes whose row
        // has bit 2 set therefore store the two chunks of a pair in the opposite order: a quarter-warp then covers
        // both parities x 4 swizzle phases = all 8 bank groups.
        auto store = [&](auto fast, const auto& v, int nj, int r, auto cj, int climit, uint8_t* hi_base, uint8_t* lo_base) {
            if constexpr (decltype(fast)::value) {
                const bool swap = (r >> 2) & 1;
#pragma unroll
                for (int j = 0; j < nj; j += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const bool second = swap != (h == 1);
                        const float4 val = second ? v[j + 1] : v[j];
                        const int c = cj(j) + (second ? 1 : 0);
                        if (c < climit) {
                            float4 hi, lo;
                            split4(val, hi, lo);
                            const uint32_t off = mn_offset(r, c, kChunk);
                            *reinterpret_cast<float4*>(hi_base + off) = hi;
                            *reinterpret_cast<float4*>(lo_base + off) = lo;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < nj; ++j) {
                    if (cj(j) < climit) {
                        float4 hi, lo;
                        split4(v[j], hi, lo);
                        const uint32_t off = mn_offset(r, cj(j), kChunk);
                        *reinterpret_cast<float4*>(hi_base + off) = hi;
                        *reinterpret_cast<float4*>(lo_base + off) = lo;
                    }
                }
            }
        };
        uint32_t use = 0;
        if (kPipe && c_begin + grp < c_end) fetch(c_begin + grp, gv, xv);
        for (uint32_t ch = c_begin + grp; ch < c_end; ch += kGroups, ++use) {
            float4 gc[kGJ], xc[kXJ];
            if (kPipe) {
#pragma unroll
                for (int j = 0; j < kGJ; ++j) gc[j] = gv[j];
#pragma unroll
                for (int j = 0; j < kXJ; ++j) xc[j] = xv[j];
                if (ch + kGroups < c_end) fetch(ch + kGroups, gv, xv);  // in flight while this chunk is stored
            } else {
                fetch(ch, gc, xc);
            }
            mbar_wait(&bar_empty[grp], (use & 1) ^ 1);
            store(std::integral_constant<bool, GLoader::kRowFast>{}, gc, kGJ, rg, cg, kG / 4, g_hi, g_lo);
            store(std::integral_constant<bool, XLoader::kRowFast>{}, xc, kXJ, rx, cx, x4, x_hi, x_lo);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[grp]);
        }
    } else if (warp == kMmaWarp) {
        const uint32_t idesc = idesc_tf32_mn(kMo, No);
        const uint32_t base = smem_u32(smem);
        uint32_t n = 0;
        for (uint32_t ch = c_begin; ch < c_end; ++ch, ++n) {
            const uint32_t s = n % kGroups;
            mbar_wait(&bar_full[s], (n / kGroups) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t gh = mn_desc_lo(base + s * stage), gl = mn_desc_lo(base + s * stage + g_half);
                const uint32_t xh = mn_desc_lo(base + s * stage + 2 * g_half);
                const uint32_t xl = mn_desc_lo(base + s * stage + 2 * g_half + x_half);
#pragma unroll
                for (uint32_t k = 0; k < kChunk / 8; ++k) {  // 8 rows = 1024 B = 64 descriptor units per K-step
#pragma unroll
                    for (uint32_t mt = 0; mt < kMT; ++mt) {  // accumulator tile mt <- operand-G columns [128 mt, +128)
                        const uint32_t go = mt * (4 * kBlockBytes >> 4) + 64 * k, d = tmem_base + mt * No;
                        mma_tf32_mn(d, gl + go, xh + 64 * k, idesc, (n == 0 && k == 0) ? 0u : 1u);
                        mma_tf32_mn(d, gh + go, xl + 64 * k, idesc, 1u);
                        mma_tf32_mn(d, gh + go, xh + 64 * k, idesc, 1u);
                    }
                }
                commit(&bar_empty[s]);
                if (ch + 1 == c_end) commit(&bar_done);
            }
            __syncwarp();
        }
    } else {
        // read-out: lane = accumulator row (operand-G column), columns = operand-X columns
        const int q = warp & 3;
        if (c_begin < c_end) {
            mbar_wait(&bar_done, 0);
            fence_after_sync();
        }
        for (int mt = 0; mt < kMT; ++mt) {
            float* out = ws + ((static_cast<size_t>(blockIdx.x) * kMT + mt) * kMo + q * 32 + lane) * No;
            if (c_begin < c_end) {
                const uint32_t taddr = tmem_base + mt * No + (static_cast<uint32_t>(q * 32) << 16);
                for (int c0 = 0; c0 < No; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<float4*>(out + c0)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            } else {
                for (int c0 = 0; c0 < No; c0 += 4) *reinterpret_cast<float4*>(out + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after_sync();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ws: [grid][128 * kMT][No] partial accumulators; `gather` below adds them up (its row index runs over 128 * kMT)
template <int kMT = 1, int kGJ = 4 * kMT, int kXJ = 8, class GLoader, class XLoader>
int launch(int device, const GLoader& g, const XLoader& x, float* ws, int64_t M, int No, int* grid_out,
           cudaStream_t stream, const char* who) {
    LTGNN_REQUIRE(No <= 32 * kXJ, LTGNN_E_SHAPE, "%s: No=%d exceeds this variant's %d columns", who, No, 32 * kXJ);
    LTGNN_REQUIRE(No % 32 == 0 && No > 0 && No <= 256, LTGNN_E_SHAPE, "%s: No=%d must be a multiple of 32, <= 256", who, No);
    LTGNN_REQUIRE(M > 0 && M < (1ll << 31) - kChunk, LTGNN_E_SHAPE, "%s: M=%lld", who, static_cast<long long>(M));
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "%s: device is sm_%d%d, need sm_100", who, di->cc_major,
                  di->cc_minor);
    const size_t smem = smem_bytes(No, kMT);
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "%s: %zu B of shared memory", who, smem);
    LTGNN_USE_DEVICE(device);
    auto kern = tgrad_kernel<GLoader, XLoader, kMT, kGJ, kXJ>;
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    uint32_t cols = 32;
    while (cols < static_cast<uint32_t>(No) * kMT) cols <<= 1;
    LTGNN_REQUIRE(cols <= 512, LTGNN_E_SHAPE, "%s: %d accumulator columns exceed tensor memory", who, No * kMT);
    const int grid = di->sm_count;
    kern<<<grid, kThreads, smem, stream>>>(g, x, ws, static_cast<uint32_t>(M), No, cols);
    LTGNN_CUDA_TRY(cudaGetLastError());
    *grid_out = grid;
    return LTGNN_OK;
}

// out[r][c] (+)= sum_p ws[p][r][c_src] for a sub-rectangle of the [128][No] accumulator.  The partials of one output
// are summed by kGatherY threads (thread y takes p = y, y + kGatherY, ...: independent loads in flight instead of one
// chain of ~150 dependent-latency loads) and combined in a fixed order: deterministic.
constexpr int kGatherX = 64, kGatherY = 8;
static __global__ void __launch_bounds__(kGatherX * kGatherY)
gather_partials_kernel(const float* __restrict__ ws, int n_parts, int part_rows, int No, int r0, int rows, int c0, int cols,
                       float* __restrict__ out, int ld_out, int accumulate) {
    __shared__ float red[kGatherY][kGatherX];
    const int x = threadIdx.x % kGatherX, y = threadIdx.x / kGatherX;
    const int i = blockIdx.x * kGatherX + x;
    float t = 0.f;
    int r = 0, c = 0;
    if (i < rows * cols) {
        r = i / cols;
        c = i - r * cols;
        const size_t src = static_cast<size_t>(r0 + r) * No + c0 + c, stride = static_cast<size_t>(part_rows) * No;
        for (int p = y; p < n_parts; p += kGatherY) t += ws[p * stride + src];
    }
    red[y][x] = t;
    __syncthreads();
    if (y == 0 && i < rows * cols) {
        float s = accumulate ? out[r * ld_out + c] : 0.f;
#pragma unroll
        for (int k = 0; k < kGatherY; ++k) s += red[k][x];
        out[r * ld_out + c] = s;
    }
}

inline int gather(const float* ws, int n_parts, int No, int r0, int rows, int c0, int cols, float* out, int ld_out,
                  int accumulate, cudaStream_t stream, int part_rows = kMo) {
    const int n = rows * cols;
    gather_partials_kernel<<<(n + kGatherX - 1) / kGatherX, kGatherX * kGatherY, 0, stream>>>(
        ws, n_parts, part_rows, No, r0, rows, c0, cols, out, ld_out, accumulate);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}


}  // namespace tgrad
}  // namespace ltgnn
