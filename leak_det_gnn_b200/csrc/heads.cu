// heads.cu -- the two read-out heads of the detector on top of the last GCN layer.
//
//   pipe head  (models/detector.py:76-88, 204-211): for every class pipe p = (u, v)
//              feat = [h_u, h_v, |h_u - h_v|] -> Linear(3D, H) -> ReLU -> Dropout -> Linear(H, 1)
//              The reference gathers h_u / h_v into (B, P, D) tensors, concatenates a (B, P, 3D) feature
//              tensor (2.4 GB at B = 4096, P = 764) and runs cuBLAS on it.  Here the features are formed
//              on the fly by the loader warps straight from the node states; the hidden layer never leaves
//              the SM in the forward except as the saved post-activation the backward needs.
//   mean pool  (models/detector.py:214-215, PyG global_mean_pool over equal-sized graphs).
//
// Both head GEMMs keep the A operand in tensor memory (building blocks: rowgemm_ts.cuh): a thread owns one pipe
// row (= TMEM lane), splits it into TF32 hi/lo in registers and tcgen05.st's it, so shared memory only holds the
// whole weight W1 (192 KB as hi + lo) and every tile is produced exactly once.  Global memory is always touched
// in 128-byte row segments; the switch to / from the row-per-thread form goes through patch.cuh.
#include "patch.cuh"
#include "rowgemm_ts.cuh"

using namespace ltgnn;

namespace {

// ------------------------------------------------------------------ forward
constexpr int kD = 64;        // node width supported by the fused head
constexpr int kD4 = kD / 4;

// hidden = dropout(relu(acc + b1)) for hidden units [c_begin, c_end) of one pipe row; returns their share of
// the logit, sum of hidden * w2.  The inverted-dropout scale 1 / (1 - p) is folded into W1 and b1 by the kernel
// (relu commutes with a positive scale), so a kept unit costs add, max, compare, select, fma.
struct HeadFwdEpilogue {
    const float* b1;   // [H]
    const float* w2;   // [H]
    float* part;       // [M]
    float* hpost;      // blocked-32 [Mp, H] saved post-activation, or nullptr (inference)
    int H;
    uint32_t drop_thresh16;  // round(p * 2^16), 0 = no dropout
    float keep_scale;        // 1 / (1 - drop_thresh16 / 2^16)
    uint64_t drop_seed;
    template <class Pull>
    __device__ __forceinline__ float operator()(uint32_t row, int c_begin, int c_end, Pull&& pull) const {
        float acc = 0.f;
        const uint32_t thresh_hi = drop_thresh16 << 16;
        for (int col = c_begin; col < c_end; col += 16) {
            float4 bb[4], ww[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // issued before the tcgen05.ld wait
                bb[j] = __ldg(reinterpret_cast<const float4*>(b1 + col) + j);
                ww[j] = __ldg(reinterpret_cast<const float4*>(w2 + col) + j);
            }
            float v[16];
            pull(col, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[4 * j + 0] = fmaxf(fmaf(bb[j].x, keep_scale, v[4 * j + 0]), 0.f);
                v[4 * j + 1] = fmaxf(fmaf(bb[j].y, keep_scale, v[4 * j + 1]), 0.f);
                v[4 * j + 2] = fmaxf(fmaf(bb[j].z, keep_scale, v[4 * j + 2]), 0.f);
                v[4 * j + 3] = fmaxf(fmaf(bb[j].w, keep_scale, v[4 * j + 3]), 0.f);
            }
            if (drop_thresh16) {
                const uint64_t i8 = static_cast<uint64_t>(row) * (H >> 3) + (col >> 3);
                ptx::dropout8_mask(v, i8, drop_seed, thresh_hi);
                ptx::dropout8_mask(v + 8, i8 + 1, drop_seed, thresh_hi);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (hpost)  // blocked-32 layout (rows padded to 128 by the caller): coalesced across the warp's rows
                    ptx::stg_stream(reinterpret_cast<float4*>(hpost) + ptx::b32(row, (col >> 2) + j, H >> 2),
                                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                acc = fmaf(v[4 * j + 0], ww[j].x, acc);
                acc = fmaf(v[4 * j + 1], ww[j].y, acc);
                acc = fmaf(v[4 * j + 2], ww[j].z, acc);
                acc = fmaf(v[4 * j + 3], ww[j].w, acc);
            }
        }
        return acc;
    }
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// ------------------------------------------------------------------ forward kernel
// The generic tensor-memory skeleton (rowgemm_ts.cuh) lets every loader thread fetch its own row, which for
// gathered node states means 32 different 128-byte lines per load instruction: ncu showed the L1 data pipe
// (l1tex lsu wavefronts) at 70 % of peak and the tensor pipe at 29 %.  Here a warp fetches the rows
// cooperatively -- 8 lanes read the 128 B of one row, 4 rows per instruction, every line used in full -- and
// turns them into the row-per-thread form tcgen05.st needs through a 4 KB XOR-swizzled shared-memory patch
// (conflict-free both ways).  Each end-node row is fetched once per tile and reused for the |h_u - h_v| block.
//
//   warps 0-7   LOADERS   group g = warp / 4 owns feature columns [32 g, 32 g + 32) of every tile and emits
//                         three A stages: h_u, h_v, |h_u - h_v|  (K blocks 2 seg + g of W1)
//   warp  8     MMA       elected lane, 3xTF32, A from tensor memory, W1 resident in shared memory (192 KB)
//   warps 9-16  EPILOGUE  tcgen05.ld + bias / ReLU / dropout / w2 dot, double-buffered accumulators; two warps
//                         per TMEM lane quadrant, 64 hidden units each (the Philox chains are latency-bound)
namespace hf {
using namespace ltgnn::ptx;
using namespace ltgnn::umma;
constexpr int kLdWarps = 8, kMmaWarp = 8, kEpWarps = 8, kThreads = (kLdWarps + 1 + kEpWarps) * 32;
constexpr int kK = 192, kN = 128, kStages = 4, kStageCols = 64;
constexpr uint32_t kScrBytes = 4096;

__global__ void __launch_bounds__(kThreads, 1)
pipe_head_fwd_kernel(const float4* __restrict__ x, const int2* __restrict__ ends, uint32_t P, uint32_t N, uint64_t magic,
                     const HeadFwdEpilogue epilogue, const float* __restrict__ W1, uint32_t M) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float part_s[2][128];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_hi = smem;
    uint8_t* b_lo = b_hi + kN * kK * 4;
    uint8_t* scratch = b_lo + kN * kK * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar_full[s], 4);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], kEpWarps);
        }
        fence_mbar_init();
    }
    rowgemm_ts::fill_b(b_hi, b_lo, W1, kK, 0, kK, kN, tid, kThreads, epilogue.keep_scale);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t acc_base = tmem_base, a_base = tmem_base + 2 * kN;
    const uint32_t n_tiles = (M + 127) / 128;
    // each CTA owns a contiguous range of tiles: consecutive tiles share a window, so its node states are fetched
    // from HBM once and the CTA's reads / writes / reductions stay local
    const uint32_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_begin = min(blockIdx.x * per_cta, n_tiles), t_end = min(t_begin + per_cta, n_tiles);

    if (warp < kLdWarps) {
        const int grp = warp >> 2, quad = warp & 3;
        const patch::Patch patch(scratch + warp * kScrBytes, lane);
        const uint32_t lane_base = a_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int sub = lane >> 3, ch = lane & 7;
        uint32_t t = 0;
        for (uint32_t tile = t_begin; tile < t_end; ++tile, ++t) {
            const uint32_t row = tile * 128 + quad * 32 + lane;
            uint32_t iu = 0, iv = 0;  // node rows of the two ends (pad rows read node row 0; their results are unused)
            if (row < M) {
                const uint32_t b = magic ? fastdiv(row, magic) : row;
                const int2 e = __ldg(ends + (row - b * P));
                iu = b * N + e.x;
                iv = b * N + e.y;
            }
            float4 gu[8], gv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t ru = __shfl_sync(0xffffffffu, iu, 4 * k + sub), rv = __shfl_sync(0xffffffffu, iv, 4 * k + sub);
                gu[k] = __ldg(x + static_cast<size_t>(ru) * kD4 + grp * 8 + ch);
                gv[k] = __ldg(x + static_cast<size_t>(rv) * kD4 + grp * 8 + ch);
            }
            // emit one A stage: split 32 values into TF32 hi / lo and store them to this thread's TMEM lane
            auto emit = [&](int seg, auto&& value) {
                const uint32_t stage = 6 * t + 3 * grp + seg, slot = stage & 3;
                mbar_wait_relaxed(&bar_empty[slot], ((stage >> 2) & 1) ^ 1);  // slot consumed by the tensor core
                fence_after_sync();
                const uint32_t st_addr = lane_base + slot * kStageCols;
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                    float hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = value(c + j);
                        hi[j] = tf32_hi(a);
                        lo[j] = a - hi[j];
                    }
                    tmem_st8(st_addr + c, hi);
                    tmem_st8(st_addr + 32 + c, lo);
                }
                rowgemm_ts::tmem_wait_st();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[slot]);
            };
            float u[32], v[32];
            patch::transpose_in(patch, gu, u);
            emit(0, [&](int j) { return u[j]; });   // h_v still in flight
            patch::transpose_in(patch, gv, v);
            emit(1, [&](int j) { return v[j]; });
            emit(2, [&](int j) { return fabsf(u[j] - v[j]); });
        }
    } else if (warp == kMmaWarp) {
        const uint32_t idesc = idesc_tf32(128, kN);
        const uint32_t bh = desc_lo(smem_u32(b_hi)), bl = desc_lo(smem_u32(b_lo));
        constexpr uint32_t kg_units = static_cast<uint32_t>(kN) * 128u >> 4;
        uint32_t t = 0;
        for (uint32_t tile = t_begin; tile < t_end; ++tile, ++t) {
            const uint32_t a = t & 1;
            mbar_wait_relaxed(&bar_tempty[a], ((t >> 1) & 1) ^ 1);
            const uint32_t d = acc_base + a * kN;
#pragma unroll 1
            for (uint32_t s = 0; s < 6; ++s) {
                const uint32_t stage = 6 * t + s, slot = stage & 3;
                const uint32_t kg = s < 3 ? 2 * s : 2 * (s - 3) + 1;  // K block of W1 this stage multiplies
                mbar_wait_relaxed(&bar_full[slot], (stage >> 2) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = a_base + slot * kStageCols, a_lo = a_hi + 32;
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t boff = kg * kg_units + 2 * k;
                        rowgemm_ts::mma_tf32_ts(d, a_lo + 8 * k, bh + boff, idesc, (s == 0 && k == 0) ? 0u : 1u);
                        rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bl + boff, idesc, 1u);
                        rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bh + boff, idesc, 1u);
                    }
                    commit(&bar_empty[slot]);
                    if (s == 5) commit(&bar_tfull[a]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3, half = (warp - kMmaWarp - 1) >> 2;  // TMEM lane quadrant = warp % 4
        uint32_t t = 0;
        for (uint32_t tile = t_begin; tile < t_end; ++tile, ++t) {
            const uint32_t a = t & 1;
            mbar_wait_relaxed(&bar_tfull[a], (t >> 1) & 1);
            fence_after_sync();
            const uint32_t taddr = acc_base + a * kN + (static_cast<uint32_t>(q * 32) << 16);
            const uint32_t row = tile * 128 + q * 32 + lane;
            const float acc = epilogue(row, half * (kN / 2), (half + 1) * (kN / 2),
                                       [&](int c0, float (&v)[16]) { tmem_ld16(taddr + c0, v); });
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[a]);
            // the two warps of a quadrant add their halves of the logit in a fixed order
            if (half == 1) part_s[a][q * 32 + lane] = acc;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (half == 0 && row < M) epilogue.part[row] = acc + part_s[a][q * 32 + lane];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}
}  // namespace hf

// ------------------------------------------------------------------ backward (input gradient)
// A[row, j] = d loss / d pre[row, j] = dlogit[row] * w2[j] * (hpost[row, j] > 0 ? scale : 0)
struct DpreLoader {
    const float4* hpost;   // blocked-32 [Mp, H]
    const float* dlogit;   // [M]
    const float4* w2;      // [H/4]
    float scale;
};

__device__ __forceinline__ float sgn(float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }

// dfeat[row, 0:3D] = dpre[row, :] W1 is scattered back to the two end nodes: +u for the h_u block, +v for the
// h_v block, +-sign(h_u - h_v) for the |.| block (torch: d|x| = sign(x), 0 at 0).  Several pipes share a node,
// so the adds are fp32 reductions in L2 (red.global.add.v4.f32) -- like the reference's index_add_ in autograd.
//
// Same warp roles as the forward kernel.  The epilogue warps (two per TMEM lane quadrant, 32 of the 64 node
// features each) turn the row-per-thread accumulator blocks into 128-byte row segments through the swizzled
// shared-memory patch, so that the node-state reads for the sign and the reductions touch 4 full lines per
// instruction instead of 32 partial ones, and every end node receives one reduction per 128 B.
namespace hb {
using namespace ltgnn::ptx;
using namespace ltgnn::umma;
constexpr int kLdWarps = 4, kMmaWarp = 4, kEpWarps = 8, kThreads = (kLdWarps + 1 + kEpWarps) * 32;  // 13 warps: 128 regs
constexpr int kK = 128, kN = 192, kStages = 2, kStageCols = 64;
constexpr uint32_t kScrBytes = 4096;

__global__ void __launch_bounds__(kThreads, 1)
pipe_head_bwd_dx_kernel(const DpreLoader loader, const float4* __restrict__ x, float4* __restrict__ dx,
                        const int2* __restrict__ ends, uint32_t P, uint32_t N, uint64_t magic,
                        const float* __restrict__ W1, uint32_t M) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_hi = smem;  // B[n = feature column (192)][k = hidden unit (128)] = W1[k][n]
    uint8_t* b_lo = b_hi + kN * kK * 4;
    uint8_t* scratch = b_lo + kN * kK * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar_full[s], 4);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], kEpWarps);
        }
        fence_mbar_init();
    }
    rowgemm_ts::fill_b(b_hi, b_lo, W1, kN, 1, kK, kN, tid, kThreads);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t acc_base = tmem_base, a_base = tmem_base + 2 * kN;
    const uint32_t n_tiles = (M + 127) / 128;
    // each CTA owns a contiguous range of tiles: consecutive tiles share a window, so its node states are fetched
    // from HBM once and the CTA's reads / writes / reductions stay local
    const uint32_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t t_begin = min(blockIdx.x * per_cta, n_tiles), t_end = min(t_begin + per_cta, n_tiles);

    if (warp < kLdWarps) {
        // The four loader warps (thread = pipe row = TMEM lane) emit the four 32-wide K blocks of every tile into the
        // two A slots in turn.  The saved activations come from HBM (~1.5 us away): the 8 loads of the NEXT block
        // are issued before the current one is split and stored.
        const int quad = warp & 3;
        const uint32_t st_base = a_base + (static_cast<uint32_t>(quad * 32) << 16);
        const float4* w2 = loader.w2;
        uint32_t stage = 0;
        uint32_t tile = t_begin;
        int kg = 0;
        float4 raw[8];
        float g = 0.f;
        auto fetch = [&]() {  // pad rows exist in the blocked-32 tensor; their g is 0
            const uint32_t row = tile * 128 + quad * 32 + lane;
#pragma unroll
            for (int j = 0; j < 8; ++j) raw[j] = ldg_stream(loader.hpost + b32(row, kg * 8 + j, kK / 4));
            if (kg == 0) g = row < M ? __ldg(loader.dlogit + row) * loader.scale : 0.f;
            // one register stage cannot cover HBM latency: pull the same block of the next tile into L2 now
            if (tile + 1 < t_end) {
                const uint32_t nrow = row + 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(loader.hpost + b32(nrow, kg * 8 + j, kK / 4)));
            }
        };
        if (tile < t_end) fetch();
        while (tile < t_end) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = __ldg(w2 + kg * 8 + j);
                v[4 * j] = raw[j].x > 0.f ? g * w.x : 0.f;
                v[4 * j + 1] = raw[j].y > 0.f ? g * w.y : 0.f;
                v[4 * j + 2] = raw[j].z > 0.f ? g * w.z : 0.f;
                v[4 * j + 3] = raw[j].w > 0.f ? g * w.w : 0.f;
            }
            if (++kg == 4) {
                kg = 0;
                ++tile;
            }
            if (tile < t_end) fetch();
            const uint32_t slot = stage & 1;
            mbar_wait_relaxed(&bar_empty[slot], ((stage >> 1) & 1) ^ 1);
            ++stage;
            fence_after_sync();
            const uint32_t st_addr = st_base + slot * kStageCols;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                float hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    hi[j] = tf32_hi(v[c + j]);
                    lo[j] = v[c + j] - hi[j];
                }
                tmem_st8(st_addr + c, hi);
                tmem_st8(st_addr + 32 + c, lo);
            }
            rowgemm_ts::tmem_wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[slot]);
        }
    } else if (warp == kMmaWarp) {
        const uint32_t idesc = idesc_tf32(128, kN);
        const uint32_t bh = desc_lo(smem_u32(b_hi)), bl = desc_lo(smem_u32(b_lo));
        constexpr uint32_t kg_units = static_cast<uint32_t>(kN) * 128u >> 4;
        uint32_t t = 0;
        for (uint32_t tile = t_begin; tile < t_end; ++tile, ++t) {
            const uint32_t a = t & 1;
            mbar_wait_relaxed(&bar_tempty[a], ((t >> 1) & 1) ^ 1);
            const uint32_t d = acc_base + a * kN;
#pragma unroll 1
            for (uint32_t kg = 0; kg < 4; ++kg) {
                const uint32_t stage = 4 * t + kg, slot = stage & 1;
                mbar_wait_relaxed(&bar_full[slot], (stage >> 1) & 1);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t a_hi = a_base + slot * kStageCols, a_lo = a_hi + 32;
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t boff = kg * kg_units + 2 * k;
                        rowgemm_ts::mma_tf32_ts(d, a_lo + 8 * k, bh + boff, idesc, (kg == 0 && k == 0) ? 0u : 1u);
                        rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bl + boff, idesc, 1u);
                        rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bh + boff, idesc, 1u);
                    }
                    commit(&bar_empty[slot]);
                    if (kg == 3) commit(&bar_tfull[a]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3, half = (warp - kMmaWarp - 1) >> 2;  // TMEM lane quadrant = warp % 4
        const patch::Patch patch(scratch + (warp - kMmaWarp - 1) * kScrBytes, lane);
        const int sub = lane >> 3, ch = lane & 7;
        uint32_t t = 0;
        for (uint32_t tile = t_begin; tile < t_end; ++tile, ++t) {
            const uint32_t a = t & 1;
            const uint32_t row0 = tile * 128 + q * 32;
            uint32_t iu = 0, iv = 0;
            if (row0 + lane < M) {
                const uint32_t b = magic ? fastdiv(row0 + lane, magic) : row0 + lane;
                const int2 e = __ldg(ends + (row0 + lane - b * P));
                iu = b * N + e.x;
                iv = b * N + e.y;
            }
            // sign(h_u - h_v) of this lane's 8 row segments, packed as two bit masks (bit 4 k + component): the
            // node states do not depend on the accumulator, so their latency is paid before the wait
            uint32_t pos = 0, neg = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t ru = __shfl_sync(0xffffffffu, iu, 4 * k + sub), rv = __shfl_sync(0xffffffffu, iv, 4 * k + sub);
                const float4 p = __ldg(x + ru * kD4 + half * 8 + ch), r = __ldg(x + rv * kD4 + half * 8 + ch);
                const float d[4] = {p.x - r.x, p.y - r.y, p.z - r.z, p.w - r.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pos |= (d[i] > 0.f ? 1u : 0u) << (4 * k + i);
                    neg |= (d[i] < 0.f ? 1u : 0u) << (4 * k + i);
                }
            }
            mbar_wait_relaxed(&bar_tfull[a], (t >> 1) & 1);
            fence_after_sync();
            const uint32_t taddr = acc_base + a * kN + (static_cast<uint32_t>(q * 32) << 16) + 32 * half;
            auto pull32 = [&](uint32_t col, float4 (&g)[8]) {
                float v[32];
                tmem_ld32(taddr + col, v);
                patch::transpose_out(patch, v, g);
            };
            auto signed_c = [&](float c, int bit) { return (pos >> bit) & 1u ? c : ((neg >> bit) & 1u ? -c : 0.f); };
            float4 gc[8], ga[8];
            pull32(2 * kD, gc);  // d / d |x_u - x_v|  ->  +- sign
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                gc[k].x = signed_c(gc[k].x, 4 * k); gc[k].y = signed_c(gc[k].y, 4 * k + 1);
                gc[k].z = signed_c(gc[k].z, 4 * k + 2); gc[k].w = signed_c(gc[k].w, 4 * k + 3);
            }
            pull32(0, ga);       // d / d x_u
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t ru = __shfl_sync(0xffffffffu, iu, 4 * k + sub);
                if (row0 + 4 * k + sub < M)
                    atomicAdd(dx + ru * kD4 + half * 8 + ch,
                              make_float4(ga[k].x + gc[k].x, ga[k].y + gc[k].y, ga[k].z + gc[k].z, ga[k].w + gc[k].w));
            }
            pull32(kD, ga);      // d / d x_v
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[a]);  // accumulator drained
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t rv = __shfl_sync(0xffffffffu, iv, 4 * k + sub);
                if (row0 + 4 * k + sub < M)
                    atomicAdd(dx + rv * kD4 + half * 8 + ch,
                              make_float4(ga[k].x - gc[k].x, ga[k].y - gc[k].y, ga[k].z - gc[k].z, ga[k].w - gc[k].w));
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}
}  // namespace hb

// ------------------------------------------------------------------ mean pool and its adjoint
__global__ void __launch_bounds__(256)
mean_pool_kernel(const float4* __restrict__ x, float4* __restrict__ pooled, int64_t B, int N, int d4) {
    __shared__ float4 red[256];
    const int tid = threadIdx.x;
    const int c = tid % d4, r0 = tid / d4, rstep = 256 / d4;
    const float inv = 1.f / static_cast<float>(N);
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const float4* xb = x + b * N * d4 + c;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < N; r += rstep) {
            const float4 v = ptx::ldg_stream(xb + static_cast<int64_t>(r) * d4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        red[tid] = s;
        __syncthreads();
        if (tid < d4) {
            float4 t = red[tid];
            for (int k = tid + d4; k < 256; k += d4) {
                const float4 o = red[k];
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            }
            pooled[b * d4 + tid] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
        }
        __syncthreads();
    }
}

// dx[b, i, :] = dpooled[b, :] / N   (adjoint of the mean; also the initial value the pipe-head scatter adds onto)
__global__ void __launch_bounds__(256)
pool_bwd_fill_kernel(const float4* __restrict__ dpooled, float4* __restrict__ dx, int64_t total4, int N, int d4) {
    const float inv = 1.f / static_cast<float>(N);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t per_b = static_cast<int64_t>(N) * d4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int64_t b = i / per_b;
        const int c = static_cast<int>(i % d4);
        const float4 g = __ldg(dpooled + b * d4 + c);
        ptx::stg_stream(dx + i, make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv));
    }
}

uint64_t magic_of(int P) { return P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull; }

int head_shape_check(int D, int H, const char* who) {
    LTGNN_REQUIRE(D == 64, LTGNN_E_SHAPE, "%s: node width D=%d not supported by the fused head (64 only)", who, D);
    LTGNN_REQUIRE(H == 128, LTGNN_E_SHAPE, "%s: hidden width H=%d not supported by the fused head (128 only)", who, H);
    return LTGNN_OK;
}

}  // namespace

extern "C" int ltgnn_pipe_head_fwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                                   const int32_t* ends, const float* W1, const float* b1, const float* w2,
                                   float drop_p, uint64_t drop_seed, float* part, float* hpost, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P > 0, LTGNN_E_ARG, "pipe_head_fwd: B=%lld N=%d P=%d", static_cast<long long>(B), N, P);
    int rc = head_shape_check(D, H, "pipe_head_fwd");
    if (rc) return rc;
    LTGNN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LTGNN_E_ARG, "pipe_head_fwd: dropout p=%f", drop_p);
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && ends && W1 && b1 && w2 && part, LTGNN_E_ARG, "pipe_head_fwd: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(W1) && aligned16(b1) && aligned16(w2) && aligned16(hpost), LTGNN_E_ALIGN,
                  "pipe_head_fwd: 16-byte alignment required");
    const int64_t M = B * P;
    LTGNN_REQUIRE(M < (1ll << 31), LTGNN_E_SHAPE, "pipe_head_fwd: B*P too large");
    const uint32_t t16 = drop_p > 0.f ? static_cast<uint32_t>(static_cast<double>(drop_p) * 65536.0 + 0.5) : 0u;
    HeadFwdEpilogue ep{b1, w2, part, hpost, H, t16, 1.f / (1.f - static_cast<float>(t16) / 65536.f), drop_seed};
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "pipe_head_fwd: device is sm_%d%d, need sm_100", di->cc_major,
                  di->cc_minor);
    LTGNN_REQUIRE(B * N < (1ll << 31), LTGNN_E_SHAPE, "pipe_head_fwd: B*N too large");
    const size_t smem = 1024 + 2ull * hf::kN * hf::kK * 4 + hf::kLdWarps * hf::kScrBytes;
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "pipe_head_fwd: %zu B of shared memory", smem);
    LTGNN_USE_DEVICE(device);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(hf::pipe_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const int64_t tiles = (M + 127) / 128;
    const int grid = static_cast<int>(tiles < di->sm_count ? tiles : di->sm_count);
    hf::pipe_head_fwd_kernel<<<grid, hf::kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), static_cast<uint32_t>(P),
        static_cast<uint32_t>(N), magic_of(P), ep, W1, static_cast<uint32_t>(M));
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int ltgnn_pipe_head_bwd_dx(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                                      const int32_t* ends, const float* W1, const float* w2, const float* hpost,
                                      const float* dlogit, float gate_scale, float* dX, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P > 0, LTGNN_E_ARG, "pipe_head_bwd_dx: B=%lld N=%d P=%d", static_cast<long long>(B), N, P);
    int rc = head_shape_check(D, H, "pipe_head_bwd_dx");
    if (rc) return rc;
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && ends && W1 && w2 && hpost && dlogit && dX, LTGNN_E_ARG, "pipe_head_bwd_dx: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(W1) && aligned16(w2) && aligned16(hpost) && aligned16(dX), LTGNN_E_ALIGN,
                  "pipe_head_bwd_dx: 16-byte alignment required");
    const int64_t M = B * P;
    DpreLoader ld{reinterpret_cast<const float4*>(hpost), dlogit, reinterpret_cast<const float4*>(w2), gate_scale};
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "pipe_head_bwd_dx: device is sm_%d%d, need sm_100",
                  di->cc_major, di->cc_minor);
    LTGNN_REQUIRE(M < (1ll << 31) && B * N < (1ll << 28), LTGNN_E_SHAPE, "pipe_head_bwd_dx: B*P or B*N too large");
    const size_t smem = 1024 + 2ull * hb::kN * hb::kK * 4 + hb::kEpWarps * hb::kScrBytes;
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "pipe_head_bwd_dx: %zu B of shared memory", smem);
    LTGNN_USE_DEVICE(device);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(hb::pipe_head_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const int64_t tiles = (M + 127) / 128;
    const int grid = static_cast<int>(tiles < di->sm_count ? tiles : di->sm_count);
    hb::pipe_head_bwd_dx_kernel<<<grid, hb::kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
        ld, reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(dX), reinterpret_cast<const int2*>(ends),
        static_cast<uint32_t>(P), static_cast<uint32_t>(N), magic_of(P), W1, static_cast<uint32_t>(M));
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int ltgnn_mean_pool_fwd(int device, int64_t B, int32_t N, int32_t D, const float* X, float* pooled,
                                   void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && D > 0, LTGNN_E_ARG, "mean_pool_fwd: B=%lld N=%d D=%d", static_cast<long long>(B), N, D);
    LTGNN_REQUIRE(D % 4 == 0 && 256 % (D / 4) == 0, LTGNN_E_SHAPE, "mean_pool_fwd: D=%d must be 4 * a divisor of 256", D);
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && pooled, LTGNN_E_ARG, "mean_pool_fwd: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(pooled), LTGNN_E_ALIGN, "mean_pool_fwd: 16-byte alignment required");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    const int64_t cap = static_cast<int64_t>(di->sm_count) * 8;
    mean_pool_kernel<<<static_cast<int>(B < cap ? B : cap), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(pooled), B, N, D / 4);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int ltgnn_mean_pool_bwd_fill(int device, int64_t B, int32_t N, int32_t D, const float* dpooled, float* dX,
                                        void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && D > 0 && D % 4 == 0, LTGNN_E_ARG, "mean_pool_bwd_fill: B=%lld N=%d D=%d",
                  static_cast<long long>(B), N, D);
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(dpooled && dX, LTGNN_E_ARG, "mean_pool_bwd_fill: null tensor");
    LTGNN_REQUIRE(aligned16(dpooled) && aligned16(dX), LTGNN_E_ALIGN, "mean_pool_bwd_fill: 16-byte alignment required");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    const int64_t total4 = B * N * (D / 4);
    int64_t blocks = (total4 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(di->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    pool_bwd_fill_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(dpooled), reinterpret_cast<float4*>(dX), total4, N, D / 4);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
