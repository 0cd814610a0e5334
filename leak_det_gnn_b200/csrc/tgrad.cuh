// tgrad.cuh -- weight gradients on tensor cores:  D[Mo x No] = sum over rows of  G[row, :Mo]^T  X[row, :No]
//
// The reduction dimension of this GEMM is the ROW index, so both operands are "MN-major" for the tensor
// core.  For 32-bit (TF32) MN-major operands tcgen05 accepts exactly one shared-memory layout,
// SWIZZLE_128B_BASE32B: blocks of [32 columns][rows][128 B] in which the 32-BYTE chunk c of row r sits at
// chunk c ^ (r & 3) (atoms of 4 rows x 128 B).  Descriptor: leading-dimension offset = one 32-column block,
// stride offset = one group of 4 rows (512 B).  One tcgen05.mma (kind::tf32, K = 8) consumes 8 rows;
// 3xTF32 as everywhere.
//
// Accumulation.  The tensor core adds into its fp32 accumulator with TRUNCATION: a bias of ~2e-8 |acc| per MMA that
// grows linearly with the number of accumulations.  Measured on this kernel with one accumulator for the CTA's whole
// row range: 4e-6 after 570 rows per CTA, 1.2e-4 after 18 000 (the B = 4096 bench size); and on the ill-conditioned
// sums a cross-entropy gradient produces (sum of |terms| ~ 100 x |sum|) 1e-4 already at B = 128, 30 x the error of an
// fp32 FMA loop.  So the tensor core only ever accumulates the hi*hi products of ONE 32-row chunk (4 MMAs) into a
// `main` temporary (64-column blocks, 2-4 of them in rotation); the read-out warps add it to the running total with
// round-to-nearest fp32 adds on the CUDA cores.  The total also lives in tensor memory (tcgen05.ld / add / tcgen05.st),
// so a flush moves no bytes outside the SM.  The cross products lo*hi + hi*lo are 2^-11 of the sum, their truncation
// bias is 2^-11 of an already small number, and they accumulate in a persistent `cross` accumulator for the CTA's
// life.  Error after the change: 2-4e-7 on random operands, 6-9e-6 on the ill-conditioned ones (torch fp32: 2-3e-6).
// At the end the CTA writes total + cross to a workspace and a second kernel adds the CTA partials in a fixed order.
//   warps 0-15  loaders, two groups of 8 warps, one ring stage each (rows are produced by functors, so the
//               "matrices" can be gathers or on-the-fly gradients and never exist in memory)
//   warp  16    MMA issue (elected lane)
//   warps 17-20 read-out: one warp per tensor-memory lane quadrant
#pragma once

#include <type_traits>

#include "umma.cuh"

namespace ltgnn {
namespace tgrad {

using namespace ltgnn::ptx;
using namespace ltgnn::umma;

constexpr int kMo = 128;        // accumulator rows (operand-G columns); pad / stack operands up to it
constexpr int kChunk = 32;      // rows per ring stage (4 K-steps)
constexpr int kReadWarps = 4;    // one per tensor-memory lane quadrant (more warps would cap the loaders below 80 registers)
constexpr int kMaxBuf = 4;       // `main` temporaries in rotation
constexpr int kLoaderWarps = 16;
constexpr int kGroups = 2;
constexpr int kGroupThreads = kLoaderWarps / kGroups * 32;  // 256
constexpr int kMmaWarp = kLoaderWarps;
constexpr int kThreads = (kLoaderWarps + 1 + kReadWarps) * 32;
constexpr int kThreadsReg = (kLoaderWarps + 1 + 8) * 32;  // register-total form: two read-out warps per quadrant
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

// Loaders come in two styles.  Plain: `float4 operator()(row, c)` returns the operand chunk.  Raw: `using Raw = ...;
// Raw raw(row, c); float aux(row); float4 convert(const Raw&, float aux, int c[, Side&])` -- `raw` / `aux` only LOAD (no
// arithmetic on the loaded values), `convert` turns them into the operand chunk when the chunk is stored, one pipeline
// stage later.  With few loads per thread the difference is everything: a fetch that computes on what it loads waits
// for memory right there, and the prefetch of the next chunk no longer overlaps anything.
// A raw-style G loader may also carry a per-thread side accumulation (`using Side = ...`, updated in `convert`,
// written out by `finish(Side*, cols, n, cta, group, lane)`).
template <class T, class = void>
struct has_side : std::false_type {};
template <class T>
struct has_side<T, std::void_t<typename T::Side>> : std::true_type {};
template <class T, bool = has_side<T>::value>
struct side_of { struct type {}; };
template <class T>
struct side_of<T, true> { using type = typename T::Side; };
template <class T, class = void>
struct has_raw : std::false_type {};
template <class T>
struct has_raw<T, std::void_t<typename T::Raw>> : std::true_type {};
template <class T, bool = has_raw<T>::value>
struct raw_of { using type = float4; };
template <class T>
struct raw_of<T, true> { using type = typename T::Raw; };

// Optional loader traits.  XLoader::kSlices = S: S consecutive CTAs share one row range and each produces its own No
// columns of the result (the loaders look at blockIdx.x % S themselves) -- a wide result split so that every CTA can keep
// its totals in registers.  GLoader::kExact: the G values are exactly representable in TF32 (0 / 1 masks), so the lo
// copy and its MMAs are skipped.
template <class T, class = void>
struct slices_of { static constexpr int value = 1; };
template <class T>
struct slices_of<T, std::void_t<decltype(T::kSlices)>> { static constexpr int value = T::kSlices; };
template <class T, class = void>
struct exact_of { static constexpr bool value = false; };
template <class T>
struct exact_of<T, std::void_t<decltype(T::kExact)>> { static constexpr bool value = T::kExact; };

// GLoader::Fill: a hand-written loader loop for one (G, X) pair.  `Fill::run(g, x, ring, gtid, grp, c_begin, c_end, M)`
// is called by every thread of loader group `grp` (gtid = thread index inside the group) and owns
// the whole protocol for the group's chunks c_begin + grp, c_begin + grp + kGroups, ...: the group's use-th chunk goes to
// stage grp + kGroups * (use % (kStages / kGroups)): wait for its `empty` (parity ((use / (kStages / kGroups)) & 1) ^ 1),
// write the stage in the mn_offset layout, fence_proxy_async_smem, one arrive per warp on its `full`.  The
// generic loop below costs ~430 instructions per thread and chunk for the pipe head (spills under the 72-register cap,
// per-chunk index arithmetic) and its issue slots paced that kernel; a bespoke loop needs a third of that.
template <class T, class = void>
struct fill_of { using type = void; };
template <class T>
struct fill_of<T, std::void_t<typename T::Fill>> { using type = typename T::Fill; };

// GLoader::kStages (with a Fill): ring stages, a multiple of kGroups; group g fills stages g, g + kGroups, ... in turn.
// With one stage per group the loaders and the MMA warp wait for each other half of the time (measured on the pipe
// head: loaders 45 % on `empty`, MMA warp 45 % on `full`); an exact-G stage has no lo copy, so four stages fit.
constexpr int kMaxStages = 4;
template <class T, class = void>
struct stages_of { static constexpr int value = kGroups; };
template <class T>
struct stages_of<T, std::void_t<decltype(T::kStages)>> { static constexpr int value = T::kStages; };
struct Ring {   // what a Fill needs to know about the ring
    uint8_t* base;                  // stage 0 (operand G, hi)
    uint32_t stage, x_hi, x_lo;     // bytes per stage; offsets of the X copies inside a stage
    uint64_t *full, *empty;         // [stages]
};

// chunks per flush of `main` in the register-total form (kSeg = -1, -2)
__host__ __device__ constexpr uint32_t flush_chunks(int seg) { return seg < 0 ? static_cast<uint32_t>(-(seg + 1)) + 1u : 1u; }

// float4 index of (row, columns 4 c4 .. 4 c4 + 3) inside one [128 x No] partial of the workspace: row-fast, so that a
// read-out warp (lane = row) touches 512 contiguous bytes per instruction
__host__ __device__ inline size_t ws_f4(int row, int c4) { return static_cast<size_t>(c4) * kMo + row; }

__host__ inline size_t stage_bytes(int No, int mt, bool exact_g = false) {
    return ((exact_g ? 1ull : 2ull) * kMo * mt + 2ull * No) * kChunk * 4;
}
__host__ inline size_t smem_bytes(int No, int mt, bool exact_g = false, int stages = kGroups) {
    return 1024 + stages * stage_bytes(No, mt, exact_g);
}

// GLoader: float4 operator()(uint32_t row, int c16) for c16 < kMo/4;  XLoader: same for c16 < No/4.
// Both are only called for row < M.  ws: [gridDim.x][kMo][No].
// kMT = 1 or 2 accumulator tiles of 128 rows: operand G has 128 * kMT columns and shares one pass over X.
// kGJ / kXJ: 16-byte chunks a loader thread fetches per row of G / X (chunks q, q + 8, ...).  The default covers
// every operand; a narrow operand (kGJ < 4 kMT: G columns beyond 32 kGJ are zero and are written once at start;
// kXJ = No / 32) needs few registers per stage, and then the loads of the NEXT chunk are issued before the current
// one is stored (kPipe) -- twice the bytes in flight for the skinny, HBM-bound weight gradients.
// kSeg = 0: the accurate form described in the header (per-chunk flush).  kSeg > 0: the cheap form for sums that are not
// ill-conditioned (the GRU parameter gradients, held to 5e-5): all three products go to one of two full-width
// accumulators for kSeg chunks, then the read-out warps add it to the CTA's partial in the (L2-resident) workspace.
// kSeg = -1: the accurate form for narrow results (No <= 64, kMT = 1; the GCN layers): the totals stay in REGISTERS
// of eight read-out warps (32 columns each), so a flush is one tensor-memory read of `main` -- tensor memory moves
// 64 B / clock / SM, and the load / add / store of a total that lives there costs three times that.
// kSeg = -2, -4: the same with `main` flushed every second / fourth chunk (64 / 128 rows = 8 / 16 truncating accumulations
// instead of 4: the bias grows to ~4e-7 / 9e-7 of the sum, still inside the 1e-5 bar) -- for kernels whose pace is the
// tensor-memory read.  With an exact G the cross products are flushed too (width 2 No per temporary).
template <class GLoader, class XLoader, int kMT, int kGJ = 4 * kMT, int kXJ = 8, int kSeg = 0>
__global__ void __launch_bounds__(kSeg < 0 ? kThreadsReg : kThreads, 1)
tgrad_kernel(const GLoader gload, const XLoader xload, float* __restrict__ ws, uint32_t M, int No, uint32_t tmem_cols,
             int nb, uint32_t nbuf_log2) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tmp_full[kMaxBuf], bar_tmp_empty[kMaxBuf];
    constexpr bool kExactG = exact_of<GLoader>::value;
    constexpr int kStages = stages_of<GLoader>::value;
    static_assert(kStages % kGroups == 0 && kStages <= kMaxStages, "ring stages: a multiple of the loader groups");
    static_assert(kStages == kGroups || kSeg < 0, "only the register-total form walks a deeper ring");
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kG = kMo * kMT;                        // operand-G columns
    const uint32_t g_half = kG * kChunk * 4;             // G hi (then G lo, unless G is exact)
    const uint32_t g_all = (kExactG ? 1 : 2) * g_half;
    const uint32_t x_half = static_cast<uint32_t>(No) * kChunk * 4;
    const uint32_t stage = g_all + 2 * x_half;

    if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar_full[s], kLoaderWarps / kGroups);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < kMaxBuf; ++a) {
            mbar_init(&bar_tmp_full[a], 1);
            mbar_init(&bar_tmp_empty[a], kSeg < 0 ? 8 : kReadWarps);
        }
        fence_mbar_init();
    }
    if (kGJ < 4 * kMT) {  // operand-G columns the loaders never write: zero in every stage, hi and lo
        for (uint32_t i = tid; i < kStages * (g_all / 16); i += blockDim.x) {
            const uint32_t st = i / (g_all / 16), r = i - st * (g_all / 16);
            *reinterpret_cast<float4*>(smem + st * stage + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t nbuf_mask = (1u << nbuf_log2) - 1;  // 2 or 4 temporaries in rotation
    // tensor memory: total [kMT][No] | cross [kMT][No] | main [nbuf][nb]   (nb = 64 or 32 columns per block)
    // (register-total form: cross [No] | main [nbuf][nb])
    // (register-total form with an exact G: [nbuf][main No | cross No] -- one MMA of width 2 No per K step, see below)
    const uint32_t t_cross = kSeg < 0 ? tmem_base : tmem_base + kMT * No;
    const uint32_t t_tmp = kSeg < 0 ? (kExactG ? tmem_base : tmem_base + No) : tmem_base + 2 * kMT * No;

    // contiguous range of row chunks per CTA
    constexpr uint32_t kSl = slices_of<XLoader>::value;
    static_assert(!kExactG || kSeg < 0, "exact-G operands are only handled by the register-total form");
    const uint32_t n_chunks = (M + kChunk - 1) / kChunk;
    const uint32_t n_ranges = gridDim.x / kSl;
    const uint32_t per = (n_chunks + n_ranges - 1) / n_ranges;
    const uint32_t c_begin = (blockIdx.x / kSl) * per;
    const uint32_t c_end = c_begin + per < n_chunks ? c_begin + per : n_chunks;

    using Fill = typename fill_of<GLoader>::type;
    if (warp < kLoaderWarps) {
        const int grp = warp / (kLoaderWarps / kGroups);
        if constexpr (!std::is_void<Fill>::value) {
            const Ring ring{smem, stage, g_all, g_all + x_half, bar_full, bar_empty};
            Fill::run(gload, xload, ring, tid - grp * kGroupThreads, grp, c_begin, c_end, M);
        } else {
            const int gtid = tid - grp * kGroupThreads;
            uint8_t* g_hi = smem + grp * stage;
            uint8_t* g_lo = g_hi + g_half;
            uint8_t* x_hi = g_hi + g_all;
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
            using GRaw = typename raw_of<GLoader>::type;
            using XRaw = typename raw_of<XLoader>::type;
            static_assert(!has_side<GLoader>::value || has_raw<GLoader>::value, "side accumulation needs a raw-style loader");
            GRaw gv[kGJ];
            XRaw xv[kXJ];
            float gaux_v = 0.f, xaux_v = 0.f;
            typename side_of<GLoader>::type side[kGJ] = {};
            // raw-style loaders: rows past M are fetched from row M - 1 and zeroed at convert time -- a select on the loaded
            // value would be a use of it, and the thread would wait for memory inside the fetch
            auto fetch = [&](uint32_t ch, GRaw (&gd)[kGJ], XRaw (&xd)[kXJ], float& gaux, float& xaux) {
                const uint32_t row_g = ch * kChunk + rg, row_x = ch * kChunk + rx;
                const uint32_t row_gc = row_g < M ? row_g : M - 1, row_xc = row_x < M ? row_x : M - 1;
    #pragma unroll
                for (int j = 0; j < kGJ; ++j) {
                    if constexpr (has_raw<GLoader>::value) gd[j] = gload.raw(row_gc, cg(j));
                    else gd[j] = row_g < M ? gload(row_g, cg(j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if constexpr (has_raw<GLoader>::value) gaux = gload.aux(row_gc);
    #pragma unroll
                for (int j = 0; j < kXJ; ++j) {
                    if constexpr (has_raw<XLoader>::value) xd[j] = xload.raw(row_xc, cx(j) < x4 ? cx(j) : 0);
                    else xd[j] = (row_x < M && cx(j) < x4) ? xload(row_x, cx(j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if constexpr (has_raw<XLoader>::value) xaux = xload.aux(row_xc);
            };
            // store one operand's chunks (hi and lo).  kRowFast: the 32 lanes of a warp hold 32 rows of the SAME chunk,
            // whose swizzled offsets share only 4 bank groups (2-way conflicts on every quarter-warp).  Lanes whose row
            // has bit 2 set therefore store the two chunks of a pair in the opposite order: a quarter-warp then covers
            // both parities x 4 swizzle phases = all 8 bank groups.
            auto store = [&](auto fast, auto with_lo, const auto& v, int nj, int r, auto cj, int climit, uint8_t* hi_base,
                             uint8_t* lo_base) {
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
                                if constexpr (decltype(with_lo)::value) *reinterpret_cast<float4*>(lo_base + off) = lo;
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
                            if constexpr (decltype(with_lo)::value) *reinterpret_cast<float4*>(lo_base + off) = lo;
                        }
                    }
                }
            };
            uint32_t use = 0;
            if (kPipe && c_begin + grp < c_end) fetch(c_begin + grp, gv, xv, gaux_v, xaux_v);
            for (uint32_t ch = c_begin + grp; ch < c_end; ch += kGroups, ++use) {
                GRaw gc[kGJ];
                XRaw xc[kXJ];
                float gaux = 0.f, xaux = 0.f;
                if (kPipe) {
    #pragma unroll
                    for (int j = 0; j < kGJ; ++j) gc[j] = gv[j];
    #pragma unroll
                    for (int j = 0; j < kXJ; ++j) xc[j] = xv[j];
                    gaux = gaux_v;
                    xaux = xaux_v;
                    if (ch + kGroups < c_end) fetch(ch + kGroups, gv, xv, gaux_v, xaux_v);  // in flight while this chunk is stored
                } else {
                    fetch(ch, gc, xc, gaux, xaux);
                }
                float4 gval[kGJ], xval[kXJ];
                const bool g_ok = ch * kChunk + rg < M, x_ok = ch * kChunk + rx < M;
                if constexpr (has_raw<GLoader>::value) gaux = g_ok ? gaux : 0.f;   // (a zero aux keeps the side sums clean)
    #pragma unroll
                for (int j = 0; j < kGJ; ++j) {
                    if constexpr (has_side<GLoader>::value) gval[j] = gload.convert(gc[j], gaux, cg(j), side[j]);
                    else if constexpr (has_raw<GLoader>::value) gval[j] = gload.convert(gc[j], gaux, cg(j));
                    else gval[j] = gc[j];
                    if (has_raw<GLoader>::value && !g_ok) gval[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
    #pragma unroll
                for (int j = 0; j < kXJ; ++j) {
                    if constexpr (has_raw<XLoader>::value) xval[j] = xload.convert(xc[j], xaux, cx(j));
                    else xval[j] = xc[j];
                    if (has_raw<XLoader>::value && !(x_ok && cx(j) < x4)) xval[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&bar_empty[grp], (use & 1) ^ 1);
                store(std::integral_constant<bool, GLoader::kRowFast>{}, std::integral_constant<bool, !kExactG>{}, gval, kGJ, rg, cg,
                      kG / 4, g_hi, g_lo);
                store(std::integral_constant<bool, XLoader::kRowFast>{}, std::true_type{}, xval, kXJ, rx, cx, x4, x_hi, x_lo);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[grp]);
            }
            if constexpr (has_side<GLoader>::value) {
                static_assert(GLoader::kRowFast, "side accumulation reduces over the lanes = rows of the row-fast mapping");
                int cols[kGJ];
    #pragma unroll
                for (int j = 0; j < kGJ; ++j) cols[j] = cg(j);
                gload.finish(side, cols, kGJ, blockIdx.x, grp, lane);
            }
        }
    } else if (warp == kMmaWarp && kSeg < 0) {
        constexpr uint32_t kFl = flush_chunks(kSeg);   // chunks per flush of `main` (1 or 2)
        // register-total form: one block per chunk.  This warp is a single serial instruction stream that every chunk
        // passes through (an earlier, general version of this loop -- runtime block widths, a division for the rotation
        // index -- ran 190 instructions per chunk and set the kernel's pace), so everything loop-invariant is hoisted.
        const uint32_t idesc = idesc_tf32_mn(kMo, No);
        const uint32_t base = smem_u32(smem);
        const uint32_t gh0 = mn_desc_lo(base), gl0 = mn_desc_lo(base + g_half), xh0 = mn_desc_lo(base + g_all);
        const uint32_t xl0 = mn_desc_lo(base + g_all + x_half), st16 = stage >> 4;   // descriptors advance by stage / 16
        const uint32_t n_local = c_begin < c_end ? c_end - c_begin : 0;
        for (uint32_t n = 0; n < n_local; ++n) {
            const uint32_t s = n % kStages, f = n / kFl, in_f = n % kFl, a = f & nbuf_mask;
            if (in_f == 0) mbar_wait(&bar_tmp_empty[a], ((f >> nbuf_log2) & 1) ^ 1);  // the read-out drained this temporary
            mbar_wait(&bar_full[s], (n / kStages) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t so = s * st16, g_h = gh0 + so, g_l = gl0 + so, x_h = xh0 + so, x_l = xl0 + so;
                const uint32_t d_main = t_tmp + a * nb;
                if constexpr (kExactG) {
                    // G x [X hi | X lo]: the two copies of X are adjacent in the stage, so they are ONE operand of width
                    // 2 No -- G is fetched from shared memory once per K step instead of twice (an SS-form MMA this narrow
                    // is paced by its operand fetch), and the warp issues half as many instructions.  The cross products
                    // land in the upper No columns of the temporary and are flushed with the main ones.
                    const uint32_t idesc2 = idesc_tf32_mn(kMo, 2 * No);
#pragma unroll
                    for (uint32_t k = 0; k < kChunk / 8; ++k)
                        mma_tf32_mn(d_main, g_h + 64 * k, x_h + 64 * k, idesc2, (k || in_f) ? 1u : 0u);
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < kChunk / 8; ++k) {
                        mma_tf32_mn(t_cross, g_l + 64 * k, x_h + 64 * k, idesc, (n == 0 && k == 0) ? 0u : 1u);
                        mma_tf32_mn(t_cross, g_h + 64 * k, x_l + 64 * k, idesc, 1u);
                    }
#pragma unroll
                    for (uint32_t k = 0; k < kChunk / 8; ++k)
                        mma_tf32_mn(d_main, g_h + 64 * k, x_h + 64 * k, idesc, (k || in_f) ? 1u : 0u);
                }
                if (in_f == kFl - 1 || n + 1 == n_local) commit(&bar_tmp_full[a]);
                commit(&bar_empty[s]);
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp && kSeg > 0) {
        const uint32_t idesc = idesc_tf32_mn(kMo, No);
        const uint32_t base = smem_u32(smem);
        uint32_t n = 0;
        for (uint32_t ch = c_begin; ch < c_end; ++ch, ++n) {
            const uint32_t s = n % kGroups;
            const uint32_t seg = n / (kSeg > 0 ? kSeg : 1), in_seg = n % (kSeg > 0 ? kSeg : 1), a = seg & 1;
            if (in_seg == 0) mbar_wait(&bar_tmp_empty[a], ((seg >> 1) & 1) ^ 1);  // the read-out drained this accumulator
            mbar_wait(&bar_full[s], (n / kGroups) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t gh = mn_desc_lo(base + s * stage), gl = mn_desc_lo(base + s * stage + g_half);
                const uint32_t xh = mn_desc_lo(base + s * stage + g_all);
                const uint32_t xl = mn_desc_lo(base + s * stage + g_all + x_half);
#pragma unroll
                for (uint32_t k = 0; k < kChunk / 8; ++k) {  // 8 rows = 1024 B = 64 descriptor units per K-step
#pragma unroll
                    for (uint32_t mt = 0; mt < kMT; ++mt) {  // accumulator tile mt <- operand-G columns [128 mt, +128)
                        const uint32_t go = mt * (4 * kBlockBytes >> 4) + 64 * k, d = tmem_base + (a * kMT + mt) * No;
                        mma_tf32_mn(d, gl + go, xh + 64 * k, idesc, (in_seg == 0 && k == 0) ? 0u : 1u);
                        mma_tf32_mn(d, gh + go, xl + 64 * k, idesc, 1u);
                        mma_tf32_mn(d, gh + go, xh + 64 * k, idesc, 1u);
                    }
                }
                commit(&bar_empty[s]);
                if (in_seg + 1 == static_cast<uint32_t>(kSeg > 0 ? kSeg : 1) || ch + 1 == c_end) commit(&bar_tmp_full[a]);
            }
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        const uint32_t base = smem_u32(smem);
        uint32_t n = 0, unit = 0;
        for (uint32_t ch = c_begin; ch < c_end; ++ch, ++n) {
            const uint32_t s = n % kGroups;
            mbar_wait(&bar_full[s], (n / kGroups) & 1);
            fence_after_sync();
            const uint32_t gh = mn_desc_lo(base + s * stage), gl = mn_desc_lo(base + s * stage + g_half);
            const uint32_t xh = mn_desc_lo(base + s * stage + g_all);
            const uint32_t xl = mn_desc_lo(base + s * stage + g_all + x_half);
            const uint32_t idesc_full = idesc_tf32_mn(kMo, No);
#pragma unroll 1
            for (uint32_t mt = 0; mt < kMT; ++mt) {  // accumulator tile mt <- operand-G columns [128 mt, +128)
                if (elect_one()) {  // cross products, full width, into the persistent accumulator
#pragma unroll
                    for (uint32_t k = 0; k < kChunk / 8; ++k) {  // 8 rows = 1024 B = 64 descriptor units per K-step
                        const uint32_t go = mt * (4 * kBlockBytes >> 4) + 64 * k;
                        mma_tf32_mn(t_cross + mt * No, gl + go, xh + 64 * k, idesc_full, (n == 0 && k == 0) ? 0u : 1u);
                        mma_tf32_mn(t_cross + mt * No, gh + go, xl + 64 * k, idesc_full, 1u);
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int c0 = 0; c0 < No; c0 += nb, ++unit) {
                    const uint32_t a = unit & nbuf_mask;
                    mbar_wait(&bar_tmp_empty[a], ((unit >> nbuf_log2) & 1) ^ 1);  // the read-out drained this temporary
                    fence_after_sync();
                    if (elect_one()) {
                        const int w = No - c0 < nb ? No - c0 : nb;
                        const uint32_t idesc = idesc_tf32_mn(kMo, w);
                        const uint32_t xo = static_cast<uint32_t>(c0 >> 5) * (kBlockBytes >> 4);
#pragma unroll
                        for (uint32_t k = 0; k < kChunk / 8; ++k)
                            mma_tf32_mn(t_tmp + a * nb, gh + mt * (4 * kBlockBytes >> 4) + 64 * k, xh + xo + 64 * k, idesc,
                                        k == 0 ? 0u : 1u);
                        commit(&bar_tmp_full[a]);
                        if (mt + 1 == kMT && c0 + nb >= No) commit(&bar_empty[s]);  // last unit of the chunk: stage free
                    }
                    __syncwarp();
                }
            }
        }
    } else if (kSeg < 0) {
        // read-out, register-total form: warp (q, half) owns columns [32 half, +32) of lane quadrant q
        const int q = warp & 3, half = (warp - kMmaWarp - 1) >> 2, row = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        constexpr uint32_t kFl = flush_chunks(kSeg);
        const uint32_t n_chunks_local = c_begin < c_end ? c_end - c_begin : 0;
        const uint32_t n_local = (n_chunks_local + kFl - 1) / kFl;   // flushes
        float tot[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) tot[j] = 0.f;
        if (32 * half < No) {
            for (uint32_t n = 0; n < n_local; ++n) {
                const uint32_t a = n & nbuf_mask;
                mbar_wait(&bar_tmp_full[a], (n >> nbuf_log2) & 1);
                fence_after_sync();
                float m0[16], m1[16];
                tmem_ld16_nowait(t_tmp + lane_off + a * nb + 32 * half, m0);
                tmem_ld16_nowait(t_tmp + lane_off + a * nb + 32 * half + 16, m1);
                if constexpr (kExactG) {   // the flushed cross products sit No columns further; read in two rounds (registers)
                    tmem_wait_ld();
                    tmem_pin16(m0);
                    tmem_pin16(m1);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        tot[j] += m0[j];
                        tot[16 + j] += m1[j];
                    }
                    tmem_ld16_nowait(t_tmp + lane_off + a * nb + No + 32 * half, m0);
                    tmem_ld16_nowait(t_tmp + lane_off + a * nb + No + 32 * half + 16, m1);
                }
                tmem_wait_ld();
                tmem_pin16(m0);
                tmem_pin16(m1);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tmp_empty[a]);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    tot[j] += m0[j];
                    tot[16 + j] += m1[j];
                }
            }
            if (n_local && !kExactG) {  // the last `main` commit also covers every cross MMA issued before it
                float c0v[16], c1v[16];
                tmem_ld16(t_cross + lane_off + 32 * half, c0v);
                tmem_ld16(t_cross + lane_off + 32 * half + 16, c1v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    tot[j] += c0v[j];
                    tot[16 + j] += c1v[j];
                }
            }
            float4* part = reinterpret_cast<float4*>(ws) + static_cast<size_t>(blockIdx.x) * kMo * (No / 4);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                part[ws_f4(row, 8 * half + j)] = make_float4(tot[4 * j], tot[4 * j + 1], tot[4 * j + 2], tot[4 * j + 3]);
        } else {  // No = 32: the second warp of the quadrant only keeps the barrier counts right
            for (uint32_t n = 0; n < n_local; ++n) {
                const uint32_t a = n & nbuf_mask;
                mbar_wait(&bar_tmp_full[a], (n >> nbuf_log2) & 1);
                if (lane == 0) mbar_arrive(&bar_tmp_empty[a]);
            }
        }
    } else if (kSeg > 0) {
        // read-out, cheap form: segment 0 overwrites the CTA's partial, later segments add to it (each thread only ever
        // touches its own entries; row-fast layout: a warp moves 512 contiguous bytes per instruction)
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t n_local = c_begin < c_end ? c_end - c_begin : 0;
        const uint32_t n_seg = (n_local + (kSeg > 0 ? kSeg : 1) - 1) / (kSeg > 0 ? kSeg : 1);
        float4* part = reinterpret_cast<float4*>(ws) + static_cast<size_t>(blockIdx.x) * kMT * kMo * (No / 4);
        for (uint32_t seg = 0; seg < n_seg; ++seg) {
            const uint32_t a = seg & 1;
            mbar_wait(&bar_tmp_full[a], (seg >> 1) & 1);
            fence_after_sync();
            for (int mt = 0; mt < kMT; ++mt) {
                float4* out = part + static_cast<size_t>(mt) * kMo * (No / 4);
                const uint32_t taddr = tmem_base + (a * kMT + mt) * No + (static_cast<uint32_t>(q * 32) << 16);
                for (int c0 = 0; c0 < No; c0 += 16) {
                    float4 prev[4];
                    if (seg) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) prev[j] = out[ws_f4(row, (c0 >> 2) + j)];
                    }
                    float v[16];
                    tmem_ld16(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (seg) {
                            t.x += prev[j].x; t.y += prev[j].y; t.z += prev[j].z; t.w += prev[j].w;
                        }
                        out[ws_f4(row, (c0 >> 2) + j)] = t;
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmp_empty[a]);
        }
        if (n_seg == 0) {
            for (int i = 0; i < kMT * (No / 4); ++i) part[static_cast<size_t>(i) * kMo + row] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        // read-out: lane = accumulator row (operand-G column).  total += main per (chunk, tile, block), in fp32 on the CUDA
        // cores, 16 columns per round (main and total loads in flight together, one wait).
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t n_local = c_begin < c_end ? c_end - c_begin : 0;
        uint32_t unit = 0;
        for (uint32_t n = 0; n < n_local; ++n) {
            for (int mt = 0; mt < kMT; ++mt) {
                for (int c0 = 0; c0 < No; c0 += nb, ++unit) {
                    const uint32_t a = unit & nbuf_mask;
                    mbar_wait(&bar_tmp_full[a], (unit >> nbuf_log2) & 1);
                    fence_after_sync();
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");  // my stores of the previous chunk
                    const int w = No - c0 < nb ? No - c0 : nb;
                    const uint32_t t_total = tmem_base + lane_off + mt * No + c0, t_main = t_tmp + lane_off + a * nb;
                    for (int p = 0; p < w; p += 16) {
                        float m0[16], s0[16];
                        tmem_ld16_nowait(t_main + p, m0);
                        if (n) tmem_ld16_nowait(t_total + p, s0);
                        tmem_wait_ld();
                        tmem_pin16(m0);
                        if (n) {
                            tmem_pin16(s0);
#pragma unroll
                            for (int j = 0; j < 16; ++j) m0[j] += s0[j];
                        }
                        tmem_st16(t_total + p, m0);
                    }
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tmp_empty[a]);  // the loads of `main` have completed (wait::ld)
                }
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // the last `main` commit also covers every cross MMA issued before it
        float4* part = reinterpret_cast<float4*>(ws) + static_cast<size_t>(blockIdx.x) * kMT * kMo * (No / 4);
        for (int mt = 0; mt < kMT; ++mt) {
            float4* out = part + static_cast<size_t>(mt) * kMo * (No / 4);
            for (int c0 = 0; c0 < No; c0 += 8) {
                float v[8], vc[8];
                if (n_local) {
                    tmem_ld8x2(tmem_base + lane_off + mt * No + c0, t_cross + lane_off + mt * No + c0, v, vc);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += vc[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
                out[ws_f4(row, c0 >> 2)] = make_float4(v[0], v[1], v[2], v[3]);
                out[ws_f4(row, (c0 >> 2) + 1)] = make_float4(v[4], v[5], v[6], v[7]);
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
template <int kMT = 1, int kGJ = 4 * kMT, int kXJ = 8, int kSeg = 0, class GLoader, class XLoader>
int launch(int device, const GLoader& g, const XLoader& x, float* ws, int64_t M, int No, int* grid_out,
           cudaStream_t stream, const char* who) {
    LTGNN_REQUIRE(No <= 32 * kXJ, LTGNN_E_SHAPE, "%s: No=%d exceeds this variant's %d columns", who, No, 32 * kXJ);
    LTGNN_REQUIRE(No % 32 == 0 && No > 0 && No <= 256, LTGNN_E_SHAPE, "%s: No=%d must be a multiple of 32, <= 256", who, No);
    LTGNN_REQUIRE(M > 0 && M < (1ll << 31) - kChunk, LTGNN_E_SHAPE, "%s: M=%lld", who, static_cast<long long>(M));
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED, "%s: device is sm_%d%d, need sm_100", who, di->cc_major,
                  di->cc_minor);
    const size_t smem = smem_bytes(No, kMT, exact_of<GLoader>::value, stages_of<GLoader>::value);
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "%s: %zu B of shared memory", who, smem);
    LTGNN_USE_DEVICE(device);
    auto kern = tgrad_kernel<GLoader, XLoader, kMT, kGJ, kXJ, kSeg>;
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    uint32_t cols = 32;
    // tensor memory: total + cross accumulators (No * kMT columns each) + 2..4 `main` temporaries of 64 (or 32) columns
    // (kSeg > 0: two full-width accumulators; kSeg < 0: cross + 4 temporaries as wide as the result)
    static_assert(kSeg >= 0 || kMT == 1, "the register-total form handles one accumulator tile");
    LTGNN_REQUIRE(kSeg >= 0 || No <= 64, LTGNN_E_SHAPE, "%s: the register-total form needs No <= 64, got %d", who, No);
    const int left = 512 - 2 * No * kMT;
    constexpr bool kExact = exact_of<GLoader>::value;
    const int nb = kSeg < 0 ? (kExact ? 2 * No : No) : (left >= 2 * 64 ? 64 : 32);
    int nbuf = kSeg > 0 ? 2 : (kSeg < 0 ? kMaxBuf : left / nb);
    nbuf = nbuf >= 4 ? 4 : nbuf;  // rotation counts are powers of two
    if (nbuf == 3) nbuf = 2;
    LTGNN_REQUIRE(nbuf >= 2 && left >= 0, LTGNN_E_SHAPE, "%s: 2 x %d accumulator columns leave no room in tensor memory", who,
                  No * kMT);
    const uint32_t need = kSeg > 0 ? 2 * No * kMT : (kSeg < 0 ? (kExact ? 0 : No) + nbuf * nb : 2 * No * kMT + nbuf * nb);
    while (cols < need) cols <<= 1;
    constexpr int kSl = slices_of<XLoader>::value;
    const int grid = di->sm_count / kSl * kSl;
    kern<<<grid, kSeg < 0 ? kThreadsReg : kThreads, smem, stream>>>(g, x, ws, static_cast<uint32_t>(M), No, cols, nb,
                                                                  nbuf == 4 ? 2u : 1u);
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
                       float* __restrict__ out, int ld_out, int accumulate, int p0, int p_step,
                       const float* __restrict__ row_scale) {
    __shared__ float red[kGatherY][kGatherX];
    const int x = threadIdx.x % kGatherX, y = threadIdx.x / kGatherX;
    const int i = blockIdx.x * kGatherX + x;
    float t = 0.f;
    int r = 0, c = 0;
    if (i < rows * cols) {
        r = i / cols;
        c = i - r * cols;
        // partial layout: [part_rows / 128][No / 4][128][4] (ws_f4)
        const int rr = r0 + r, cc = c0 + c;
        const size_t src = static_cast<size_t>(rr >> 7) * kMo * No + (ws_f4(rr & 127, cc >> 2) << 2) + (cc & 3);
        const size_t stride = static_cast<size_t>(part_rows) * No;
        for (int p = p0 + y * p_step; p < n_parts; p += kGatherY * p_step) t += ws[p * stride + src];
    }
    red[y][x] = t;
    __syncthreads();
    if (y == 0 && i < rows * cols) {
        float s = accumulate ? out[r * ld_out + c] : 0.f;
        float u = 0.f;
#pragma unroll
        for (int k = 0; k < kGatherY; ++k) u += red[k][x];
        if (row_scale) u *= row_scale[r];
        out[r * ld_out + c] = s + u;
    }
}

// parts p0, p0 + p_step, ... < n_parts are summed (sliced results: p_step = number of slices); row_scale[r] multiplies row r
inline int gather(const float* ws, int n_parts, int No, int r0, int rows, int c0, int cols, float* out, int ld_out,
                  int accumulate, cudaStream_t stream, int part_rows = kMo, int p0 = 0, int p_step = 1,
                  const float* row_scale = nullptr) {
    const int n = rows * cols;
    gather_partials_kernel<<<(n + kGatherX - 1) / kGatherX, kGatherX * kGatherY, 0, stream>>>(
        ws, n_parts, part_rows, No, r0, rows, c0, cols, out, ld_out, accumulate, p0, p_step, row_scale);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}


}  // namespace tgrad
}  // namespace ltgnn
