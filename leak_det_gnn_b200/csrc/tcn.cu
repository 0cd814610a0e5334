// tcn.cu -- one dilated causal convolution of the frozen TCN predictor, evaluated only where the last time step
// needs it (reference models/predictor.py:17-81, models/utils.py:169-216).
//
// The reference slides the predictor over l_det overlapping windows of every segment and evaluates all 36 time steps
// of all 8 convolutions, although the model reads only the last step (predictor.py:80).  The host side keeps, per
// convolution, the list of (window, position) rows the last step depends on (99 of 288 position evaluations) and for
// every such row the rows of the previous layer its three taps read (-1 = the zero left padding of CausalConv1d).
// That makes every convolution ONE GEMM over gathered rows with a fused epilogue:
//
//     Y[m, :] = [res[res_row[m], :] +] relu(LayerNorm(bias + sum_tap X[src[tap][m], :] W_tap^T))      K = taps * C = 384
//
//   warps 0-7   A LOADERS  gather the rows of one (tile, tap, 32-channel block) = one "fill" with cp.async into a ring of
//                          transposition patches, split fp32 -> TF32 hi/lo, tcgen05.st into an A slot (as linear.cu).  Two
//                          sets of four warps take the even and the odd fills: a warp runs its fills in series (~2 200
//                          clocks each: copy latency, split, tensor-memory store, hand-over), the tensor core needs ~860
//   warp  8     MMA        3xTF32, A from tensor memory; two 128-column accumulators ROTATING over the CTA's row tiles, so
//                          that tile n + 1 is multiplied while the epilogue normalises tile n
//   warp  17    B PRODUCER the weight (393 KB as hi + lo: cannot be resident) is split and laid out ONCE per launch by
//                          tcn_pack_weight_kernel; this warp's elected lane brings k-atom stages (32 K values x 128 outputs,
//                          hi + lo = 32 KB) in with 1-D bulk copies (TMA), up to three ahead of the tensor core
//   warps 9-16  EPILOGUE   bias, LayerNorm over the 128 channels (two-pass mean / variance; the two warps that share a
//                          row exchange partial sums through shared memory), ReLU, residual, coalesced stores via patches
// fp32-faithful like every GEMM of this library (the residual is a small difference of pressures).
#include "patch.cuh"
#include "rowgemm_ts.cuh"

using namespace ltgnn;

namespace {
namespace tc {
using namespace ltgnn::ptx;
using namespace ltgnn::umma;

constexpr int kC = 128;                  // channels: K per tap and N
constexpr int kALd = 8, kMmaWarp = 8, kEp = 8, kBWarp = kALd + 1 + kEp, kThreads = (kALd + 1 + kEp + 1) * 32;  // 18 warps (96 registers, as with 17)
constexpr int kTiles = 2;                // accumulators in tensor memory (and the most row tiles a group may have)
constexpr int kASlots = 4, kSlotCols = 64;   // 2 x 128 accumulator columns + 4 x 64 A columns = the 512 of tensor memory
constexpr int kBStages = 3, kADepth = 2;   // 8 warps x 2 patches = 64 KB of gathered rows in flight per SM
constexpr uint32_t kBStageBytes = 2u * kC * 128;  // one k-atom of the weight: 128 rows x 128 B, hi then lo

struct Params {
    const float4* X;           // [x_rows, 32] float4 rows of the previous layer
    const int32_t* src;        // [taps][M]
    const uint8_t* Wp;         // packed weight: [taps * 4 k-atoms][hi | lo][128 x 128 B, SWIZZLE_128B] (tcn_pack_weight_kernel)
    const float* bias;         // [128]
    const float* gamma;        // [128] or nullptr (no LayerNorm)
    const float* beta;
    const float4* res;         // [res_rows, 32] or nullptr
    const int32_t* res_row;    // [M]
    float4* Y;                 // [M, 32]
    uint32_t M;
    int32_t taps, relu;
    float eps;
    int32_t tiles;             // row tiles per group, 1 .. kTiles (fewer for small M: more CTAs, shorter chains)
};

// W [taps][128 outputs][128 inputs] fp32 -> the B operand stages the main kernel copies in verbatim: k-atom j = tap * 4 + kg
// (32 inputs) is 32 KB: TF32 hi of the 128 x 32 block as a K-major SWIZZLE_128B tile (row n at (n / 8) * 1024 + (n % 8) * 128,
// 16-byte chunk c at (c ^ n % 8) * 16), then lo the same way.  One thread per 16-byte chunk.
__global__ void __launch_bounds__(256)
tcn_pack_weight_kernel(const float* __restrict__ W, uint8_t* __restrict__ Wp, int taps) {
    const int i = blockIdx.x * 256 + threadIdx.x;          // (j, n, c)
    if (i >= taps * 4 * kC * 8) return;
    const int c = i & 7, n = (i >> 3) & (kC - 1), j = i >> 10, tap = j >> 2, kg = j & 3;
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (static_cast<size_t>(tap) * kC + n) * kC + kg * 32) + c);
    float4 hi, lo;
    split4(w, hi, lo);
    uint8_t* stage = Wp + static_cast<size_t>(j) * kBStageBytes;
    const uint32_t off = sw128_offset(n, c, kC);
    *reinterpret_cast<float4*>(stage + off) = hi;
    *reinterpret_cast<float4*>(stage + kC * 128 + off) = lo;
}

__global__ void __launch_bounds__(kThreads, 1)
tcn_conv_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[kASlots], a_empty[kASlots], b_full[kBStages], b_empty[kBStages], acc_full[kTiles],
        acc_empty[kTiles];
    __shared__ uint32_t tmem_base_s;
    __shared__ float stat_s[2][2][128];   // [pass][half][row]: partial sums the two warps of a row exchange
    __shared__ float4 prm_s[3][kC / 4];   // bias | gamma | beta: read as broadcast LDS.128 by the epilogue (a per-element
                                          // __ldg was 320 load instructions per thread and tile and paced the kernel)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_ring = smem;                                             // kBStages x 32 KB
    uint8_t* a_rings = b_ring + kBStages * kBStageBytes;                // kALd x kADepth patches
    uint8_t* ep_patches = a_rings + kALd * kADepth * patch::kPatchBytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < kASlots; ++s) {
            mbar_init(&a_full[s], 4);            // the four warps (lane quadrants) of the set that owns the fill
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(&b_full[s], 1);            // the expect_tx arrival of the lane that issues the bulk copy
            mbar_init(&b_empty[s], 1);
        }
        for (int t = 0; t < kTiles; ++t) {   // per tile: the tensor core refills tile t of the next group while the
            mbar_init(&acc_full[t], 1);      // epilogue is still normalising tiles t + 1, ...
            mbar_init(&acc_empty[t], kEp);
        }
        fence_mbar_init();
    }
    if (tid < kC) {
        reinterpret_cast<float*>(prm_s[0])[tid] = __ldg(p.bias + tid);
        reinterpret_cast<float*>(prm_s[1])[tid] = p.gamma ? __ldg(p.gamma + tid) : 1.f;
        reinterpret_cast<float*>(prm_s[2])[tid] = p.gamma ? __ldg(p.beta + tid) : 0.f;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t acc_base = tmem_base, a_base = tmem_base + kTiles * kC;  // 256 accumulator columns + 4 A slots
    const uint32_t rows_per_group = p.tiles * 128;
    const uint32_t n_groups = (p.M + rows_per_group - 1) / rows_per_group;
    const uint32_t my_groups = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_katoms = p.taps * 4;        // k-atoms of 32 channels per group pass

    if (warp < kALd) {
        // ---------------- A loaders: set 0 (warps 0-3) takes the even fills, set 1 (warps 4-7) the odd ones ----------------
        const int quad = warp & 3;
        const uint32_t set = warp >> 2;
        uint8_t* ring = a_rings + static_cast<size_t>(warp) * kADepth * patch::kPatchBytes;
        const uint32_t lane_base = a_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int sub = lane >> 3, ch = lane & 7;
        const uint32_t fills_per_group = n_katoms * p.tiles, n_fills = my_groups * fills_per_group;
        // The source-row indices of a fill are loaded one iteration BEFORE its cp.async are issued (idx / ok / kg_i).
        // (gi_n, j_n, tile_n) walk this set's fills incrementally -- no divisions.
        int32_t idx[8];
        uint32_t ok = 0, kg_i = 0, have = 0;
        uint32_t gi_n = 0, j_n = 0, tile_n = 0, f_n = 0;
        auto advance = [&]() {
            ++f_n;
            if (++tile_n == static_cast<uint32_t>(p.tiles)) {
                tile_n = 0;
                if (++j_n == static_cast<uint32_t>(n_katoms)) {
                    j_n = 0;
                    ++gi_n;
                }
            }
        };
        if (set) advance();
        auto load_idx = [&]() {
            have = f_n < n_fills;
            if (have) {
                const uint32_t tap = j_n >> 2;
                kg_i = j_n & 3;
                const uint32_t row0 = (blockIdx.x + gi_n * gridDim.x) * rows_per_group + tile_n * 128 + quad * 32 + sub;
                const int32_t* srct = p.src + static_cast<size_t>(tap) * p.M;
                ok = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t row = row0 + 4 * k;
                    idx[k] = __ldg(srct + (row < p.M ? row : p.M - 1));   // unconditional: the value is not used here
                    ok |= (row < p.M ? 1u : 0u) << k;
                }
                advance();
                advance();
            }
        };
        auto issue = [&](int slot_p) {   // the fill whose indices load_idx fetched last
            if (have) {
                const patch::Patch pt(ring + slot_p * patch::kPatchBytes, lane);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int32_t s = (ok >> k) & 1 ? idx[k] : -1;   // -1: left padding / past the end -> zeros
                    const float4* from = p.X + static_cast<size_t>(s < 0 ? 0 : s) * (kC / 4) + kg_i * 8 + ch;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(pt.co(k))), "l"(from),
                                 "r"(s < 0 ? 0 : 16)
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int d = 0; d < kADepth; ++d) {
            load_idx();
            issue(d);
        }
        load_idx();
        uint32_t i = 0;
        for (uint32_t f = set; f < n_fills; f += 2, ++i) {
            const int slot_p = static_cast<int>(i % kADepth);
            asm volatile("cp.async.wait_group %0;" ::"n"(kADepth - 1) : "memory");
            __syncwarp();
            float v[32];
            {
                const patch::Patch pt(ring + slot_p * patch::kPatchBytes, lane);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 t = *pt.row(j);
                    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                }
            }
            __syncwarp();
            issue(slot_p);     // this set's fill kADepth ahead: its indices arrived during the previous iteration
            load_idx();        // the one after that: in flight until the next iteration
            const uint32_t slot = f & (kASlots - 1);
            mbar_wait_relaxed(&a_empty[slot], ((f / kASlots) & 1) ^ 1);
            fence_after_sync();
            const uint32_t st_addr = lane_base + slot * kSlotCols;
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
            if (lane == 0) mbar_arrive(&a_full[slot]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == kMmaWarp) {
        // ---------------- MMA ----------------
        // This warp is ONE serial instruction stream every fill passes through, and the tensor pipe drains whenever it is
        // busy with anything else (a clock64 timeline showed ~900 clocks of index arithmetic, ring refills and modulo
        // operations between two 720-clock bursts of MMA issue): ring positions and parities are carried incrementally,
        // and the weight ring is refilled by a warp of its own.
        const uint32_t idesc = idesc_tf32(128, kC);
        const uint32_t ring_lo = desc_lo(smem_u32(b_ring));
        uint32_t stage = 0, stage_par = 0, slot = 0, slot_par = 0, nt = 0;
        if (p.tiles == 1) {
            // one tile per group (what the host picks): two fills = two k-atoms per hand-over, 24 MMAs back to back --
            // the tensor pipe only runs while this warp is issuing, so the waits, fences and commits are paid half as often
            for (uint32_t gi = 0; gi < my_groups; ++gi, ++nt) {
                const uint32_t acc = nt % kTiles, d = acc_base + acc * kC;
                mbar_wait_relaxed(&acc_empty[acc], ((nt / kTiles) & 1) ^ 1);   // the epilogue drained it
                for (int j = 0; j < n_katoms; j += 2) {   // n_katoms = 4 * taps: even
                    const uint32_t stage1 = stage + 1 == kBStages ? 0 : stage + 1, par1 = stage1 ? stage_par : stage_par ^ 1;
                    const uint32_t slot1 = slot + 1;      // slot is even here: the pair shares the parity
                    mbar_wait_relaxed(&b_full[stage], stage_par);
                    mbar_wait_relaxed(&b_full[stage1], par1);
                    mbar_wait_relaxed(&a_full[slot], slot_par);
                    mbar_wait_relaxed(&a_full[slot1], slot_par);
                    fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t h = 0; h < 2; ++h) {
                            const uint32_t st = h ? stage1 : stage, sl = h ? slot1 : slot;
                            const uint32_t bh = ring_lo + st * (kBStageBytes >> 4), bl = bh + ((kC * 128) >> 4);
                            const uint32_t a_hi = a_base + sl * kSlotCols, a_lo = a_hi + 32;
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                rowgemm_ts::mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, idesc, (j == 0 && h == 0 && k == 0) ? 0u : 1u);
                                rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, idesc, 1u);
                                rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, idesc, 1u);
                            }
                            commit(&a_empty[sl]);
                            commit(&b_empty[st]);
                        }
                        if (j == n_katoms - 2) commit(&acc_full[acc]);
                    }
                    __syncwarp();
                    slot = (slot + 2) & (kASlots - 1);
                    slot_par ^= (slot == 0);
                    stage = stage1 + 1 == kBStages ? 0 : stage1 + 1;
                    stage_par = stage ? par1 : par1 ^ 1;
                }
            }
        } else {
        for (uint32_t gi = 0; gi < my_groups; ++gi) {
            for (int j = 0; j < n_katoms; ++j) {
                mbar_wait_relaxed(&b_full[stage], stage_par);
                const uint32_t bh = ring_lo + stage * (kBStageBytes >> 4), bl = bh + ((kC * 128) >> 4);
                for (int tile = 0; tile < p.tiles; ++tile) {
                    const uint32_t ntile = nt + tile, acc = ntile % kTiles;   // accumulators rotate over the CTA's tiles
                    if (j == 0) mbar_wait_relaxed(&acc_empty[acc], ((ntile / kTiles) & 1) ^ 1);   // the epilogue drained it
                    mbar_wait_relaxed(&a_full[slot], slot_par);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t d = acc_base + acc * kC;
                        const uint32_t a_hi = a_base + slot * kSlotCols, a_lo = a_hi + 32;
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k) {
                            rowgemm_ts::mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, idesc, (j == 0 && k == 0) ? 0u : 1u);
                            rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, idesc, 1u);
                            rowgemm_ts::mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, idesc, 1u);
                        }
                        commit(&a_empty[slot]);
                        if (tile == p.tiles - 1) commit(&b_empty[stage]);
                        if (j == n_katoms - 1) commit(&acc_full[acc]);
                    }
                    __syncwarp();
                    slot = (slot + 1) & (kASlots - 1);
                    slot_par ^= (slot == 0);
                }
                if (++stage == kBStages) {
                    stage = 0;
                    stage_par ^= 1;
                }
            }
            nt += p.tiles;
        }
        }
    } else if (warp == kBWarp) {
        // ---------------- weight ring producer: k-atom stages by 1-D bulk copies (TMA), up to kBStages ahead ----------------
        const uint32_t total_katoms = my_groups * n_katoms;
        uint32_t stage = 0, par = 1, j = 0;   // parity 1 passes on a fresh barrier: the first pass over the ring does not wait
        for (uint32_t ka = 0; ka < total_katoms; ++ka) {
            mbar_wait_relaxed(&b_empty[stage], par);
            if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[stage], kBStageBytes);
                bulk_load(b_ring + stage * kBStageBytes, p.Wp + static_cast<size_t>(j) * kBStageBytes, kBStageBytes, &b_full[stage]);
            }
            __syncwarp();
            if (++j == static_cast<uint32_t>(n_katoms)) j = 0;
            if (++stage == kBStages) {
                stage = 0;
                par ^= 1;
            }
        }
    } else {
        // ---------------- epilogue: thread = (row of the tile, 64-channel half) ----------------
        const int ew = warp - kMmaWarp - 1;          // 0 .. 7
        const int q = warp & 3, half = ew >> 2;      // tensor-memory lane quadrant = warp % 4
        const int rt = q * 32 + lane;                // row within the tile
        const patch::Patch pt(ep_patches + ew * patch::kPatchBytes, lane);
        const int sub = lane >> 3, ch = lane & 7;
        const float inv_c = 1.f / static_cast<float>(kC);
        for (uint32_t gi = 0; gi < my_groups; ++gi) {
            const uint32_t g_row0 = (blockIdx.x + gi * gridDim.x) * rows_per_group;
            for (int tile = 0; tile < p.tiles; ++tile) {
                const uint32_t nt = gi * p.tiles + tile, acc = nt % kTiles;
                mbar_wait(&acc_full[acc], (nt / kTiles) & 1);
                fence_after_sync();
                const uint32_t taddr = acc_base + acc * kC + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
                const uint32_t row0 = g_row0 + tile * 128 + q * 32;   // first row of this warp's 32
                float mean = 0.f, rstd = 1.f;
                if (p.gamma) {
                    // pass 1: mean of bias + acc over the row's 128 channels
                    float s = 0.f;
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = prm_s[0][(half * 64 + c0 + j) >> 2];
                            s += (v[j] + b4.x) + (v[j + 1] + b4.y) + (v[j + 2] + b4.z) + (v[j + 3] + b4.w);
                        }
                    }
                    stat_s[0][half][rt] = s;
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                    mean = (stat_s[0][0][rt] + stat_s[0][1][rt]) * inv_c;
                    // pass 2: variance around that mean
                    float s2 = 0.f;
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = prm_s[0][(half * 64 + c0 + j) >> 2];
                            const float d0 = v[j] + b4.x - mean, d1 = v[j + 1] + b4.y - mean;
                            const float d2 = v[j + 2] + b4.z - mean, d3 = v[j + 3] + b4.w - mean;
                            s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
                        }
                    }
                    stat_s[1][half][rt] = s2;
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                    rstd = 1.f / sqrtf((stat_s[1][0][rt] + stat_s[1][1][rt]) * inv_c + p.eps);
                }
                // pass 3: normalise, activation, residual, coalesced stores (4 rows x 128 B per instruction)
                int32_t my_res = -1;
                if (p.res && row0 + lane < p.M) my_res = __ldg(p.res_row + row0 + lane);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    if (c0 == 32) {  // last read of this tile's accumulator: hand it back before the stores
                        fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int c4 = (half * 64 + c0 + j) >> 2;
                        const float4 b4 = prm_s[0][c4], g4 = prm_s[1][c4], e4 = prm_s[2][c4];
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w},
                                    ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float y = v[j + i] + bb[i];
                            if (p.gamma) y = fmaf((y - mean) * rstd, gg[i], ee[i]);
                            v[j + i] = p.relu ? fmaxf(y, 0.f) : y;
                        }
                    }
                    // residual rows: all eight loads of this chunk are in flight before the first is used (taken one by
                    // one inside the store loop they were 16 serial L2 round trips per tile: the residual convolutions ran
                    // 50 % longer per row than the others)
                    float4 a[8];
                    const int c4 = (half * 64 + c0) / 4 + ch;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int32_t rr = __shfl_sync(0xffffffffu, my_res, 4 * k + sub);
                        a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rr >= 0) a[k] = __ldg(p.res + static_cast<size_t>(rr) * (kC / 4) + c4);
                    }
                    float4 g[8];
                    patch::transpose_out(pt, v, g);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t r = row0 + 4 * k + sub;
                        if (r < p.M) {
                            g[k].x += a[k].x; g[k].y += a[k].y; g[k].z += a[k].z; g[k].w += a[k].w;
                            p.Y[static_cast<size_t>(r) * (kC / 4) + c4] = g[k];
                        }
                    }
                }
                if (p.gamma) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // stat_s reusable for the next tile
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

}  // namespace tc
}  // namespace

extern "C" int64_t ltgnn_tcn_ws_floats(int32_t taps) {
    return taps >= 1 && taps <= 8 ? static_cast<int64_t>(taps) * 4 * (tc::kBStageBytes / 4) : -1;
}

extern "C" int ltgnn_tcn_conv(int device, int64_t M, int32_t C, int32_t taps, const float* X, const int32_t* src,
                              const float* W, const float* bias, const float* gamma, const float* beta, float eps, int relu,
                              const float* res, const int32_t* res_row, float* Y, float* ws, void* stream_) {
    LTGNN_REQUIRE(M >= 0 && M < (1ll << 31) - 512, LTGNN_E_ARG, "tcn_conv: M=%lld", static_cast<long long>(M));
    LTGNN_REQUIRE(C == tc::kC, LTGNN_E_SHAPE, "tcn_conv: C=%d (the kernel is built for %d channels)", C, tc::kC);
    LTGNN_REQUIRE(taps >= 1 && taps <= 8, LTGNN_E_SHAPE, "tcn_conv: taps=%d", taps);
    LTGNN_REQUIRE((gamma == nullptr) == (beta == nullptr), LTGNN_E_ARG, "tcn_conv: gamma and beta come together");
    LTGNN_REQUIRE((res == nullptr) == (res_row == nullptr), LTGNN_E_ARG, "tcn_conv: res and res_row come together");
    if (M == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && src && W && bias && Y && ws, LTGNN_E_ARG, "tcn_conv: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(W) && aligned16(Y) && aligned16(res) && aligned16(ws), LTGNN_E_ALIGN,
                  "tcn_conv: 16-byte alignment");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "tcn_conv: device is sm_%d%d, need sm_100", di->cc_major, di->cc_minor);
    const size_t smem = 1024 + tc::kBStages * tc::kBStageBytes +
                        static_cast<size_t>(tc::kALd * tc::kADepth + tc::kEp) * patch::kPatchBytes;
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "tcn_conv: %zu B of shared memory", smem);
    LTGNN_USE_DEVICE(device);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(tc::tcn_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // One row tile per group: with the accumulators rotating over a CTA's tiles, the tensor core then works on tile n + 1
    // (and n + 2) while the epilogue normalises tile n.  Larger groups reuse every weight k-atom for 2 or 3 tiles, but the
    // B loaders are idle most of the time anyway and the epilogue of a group is then exposed (measured: never faster).
    const int64_t n_tiles = (M + 127) / 128;
    int tiles = 1;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* wp = reinterpret_cast<uint8_t*>(ws);
    tc::tcn_pack_weight_kernel<<<(taps * 4 * tc::kC * 8 + 255) / 256, 256, 0, stream>>>(W, wp, taps);
    LTGNN_CUDA_TRY(cudaGetLastError());
    tc::Params p{reinterpret_cast<const float4*>(X), src, wp, bias, gamma, beta, reinterpret_cast<const float4*>(res), res_row,
                 reinterpret_cast<float4*>(Y), static_cast<uint32_t>(M), taps, relu, eps, tiles};
    const int64_t groups = (n_tiles + tiles - 1) / tiles;
    const int grid = static_cast<int>(groups < di->sm_count ? groups : di->sm_count);
    tc::tcn_conv_kernel<<<grid, tc::kThreads, smem, stream>>>(p);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
