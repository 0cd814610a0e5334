// tgrad.cu -- tensor-core weight gradients (see tgrad.cuh) for
//   * GCNConv.lin:   dW[Do, Di] = G^T X over the B*N node rows      (autograd of models/detector.py:199)
//   * EdgeHead.mlp.0: dW1[H, 3D] = dpre^T feat and db1 = column sums of dpre over the B*P pipe rows
//                     (autograd of models/detector.py:79-87), with dpre and feat formed on the fly from the
//                     saved hidden activations / the node states -- neither (B*P, H) nor (B*P, 3D) exists.
#include "functors.cuh"
#include "tgrad.cuh"

using namespace ltgnn;
using namespace ltgnn::functors;

namespace {

// ---- pipe head.  d pre[row, j] = g[row] * w2[j] * live[row, j] with g = dlogit * scale is rank one up to the 0/1 gate,
// so with  A[j, :] = sum_rows live[row, j] * g[row] * feat[row, :]  and  a[j] = sum_rows live[row, j] * g[row]:
//      dW1[j, :] = w2[j] * A[j, :]        db1[j] = w2[j] * a[j]
//      dw2[j]    = sum_rows dlogit * hidden[row, j] = sum_c W1[j, c] * A[j, c] + b1[j] * a[j]
// (hidden = live * scale * (W1 feat + b1)): the saved hidden activations are not needed at all -- the forward keeps 1 bit
// per hidden unit instead of 4 bytes.  The GEMM operands: G = live (0 / 1: exact in TF32, so no lo copy and one MMA
// less per K step), X = F' = g * feat.  The 192 result columns are split over three CTAs per row range -- h_u, h_v and
// |h_u - h_v| -- so that each keeps its 128 x 64 totals in registers (tgrad.cuh, register-total form; a 128 x 192 total in
// tensor memory cost 3 x 96 KB of tensor-memory traffic per 32 rows and ran at 1.7 ms).  Both loaders are raw-style: with
// 6 loads per thread the kernel is latency bound unless the next chunk's loads are in flight while this one is stored
// (a first version that computed inside the fetch ran at 5.4 ms).
constexpr int kHeadSlices = 3;
constexpr int kSideParts = tgrad::kGroups * tgrad::kLoaderWarps / tgrad::kGroups;  // partials of a[] per CTA: one per loader warp
struct HeadFeatSlice {
    static constexpr int kSlices = kHeadSlices;
    const float4* x;   // node states [B*N, 16]
    const int2* ends;
    uint32_t P, N;
    uint64_t magic;    // fastdiv by P (0: P = 1)
};
struct HeadFill;
struct HeadLive {
    static constexpr bool kExact = true;
    static constexpr int kStages = 4;   // 32 KB each (no lo copy of G)
    using Fill = HeadFill;   // both operands are produced by the hand-written loop below
    const uint4* hmask;     // [Mp]: hidden unit j of a row <-> bit 31 - j % 32 of word j / 32 (heads.cu)
    const float* dlogit;    // [M]
    float scale;
    float* side_part;       // [kSideParts * gridDim.x][128]: one partial of a[] per (CTA, loader warp)
};

// The loader loop of the pipe-head weight gradient (tgrad.cuh, `Fill`).  A group = 256 threads = 32 rows x 8 threads;
// thread (r, q) owns, in row r, the 16-byte chunks q, q + 8, q + 16, q + 24 of G (hidden units 4 q + 32 blk ..+ 3: one
// nibble of gate word blk) and the chunks q, q + 8 of this CTA's 64-column feature slice -- the same swizzled offset
// inside every 32-column block, computed once.  Three-deep software pipeline: the pipe ends of chunk n + 2, the node
// rows / gate words / dlogit of chunk n + 1 and the stores of chunk n are in flight together; (window, pipe) advance
// incrementally (no division per chunk).  kSlice is a template parameter: slice 0 (h_u) also sums a[], slice 2
// (|h_u - h_v|) loads two node rows -- no CTA pays registers for both.
struct HeadFill {
    struct Rows {           // what one thread loads for one chunk
        uint4 w;
        float d;
        float4 a[2], b[2];
    };
    template <int kSlice>
    static __device__ __forceinline__ void loop(const HeadLive& g, const HeadFeatSlice& x, const tgrad::Ring& ring, int gtid,
                                                int grp, uint32_t c_begin, uint32_t c_end, uint32_t M) {
        using namespace ltgnn::ptx;
        using namespace ltgnn::umma;
        constexpr uint32_t kStep = tgrad::kGroups * tgrad::kChunk;   // rows between two chunks of this group
        const int r = gtid >> 3, q = gtid & 7, lane = gtid & 31;
        const uint32_t off = static_cast<uint32_t>(r * 128 + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4)));
        const uint32_t sh = 28 - 4 * q;
        float4 side[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) side[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t last_b = x.magic ? fastdiv(M - 1, x.magic) : M - 1, last_p = M - 1 - last_b * x.P;

        // (b, p) of this thread's row in the chunk being fetched; rows past M read row M - 1 and are zeroed at store time
        uint32_t row = (c_begin + grp) * tgrad::kChunk + r;
        uint32_t b = x.magic ? fastdiv(row < M ? row : M - 1, x.magic) : (row < M ? row : M - 1);
        uint32_t p = (row < M ? row : M - 1) - b * x.P;
        auto advance = [&]() {          // to the same row of the group's next chunk
            row += kStep;
            p += kStep;
            while (p >= x.P) {
                p -= x.P;
                ++b;
            }
            if (row >= M) {
                b = last_b;
                p = last_p;
            }
        };
        auto load_ends = [&]() { return __ldg(x.ends + p); };
        auto load_rows = [&](Rows& t, const int2 e, uint32_t rw, uint32_t bb) {
            const uint32_t rc = rw < M ? rw : M - 1;
            t.w = __ldg(g.hmask + rc);
            t.d = __ldg(g.dlogit + rc);
            const float4* xb = x.x + static_cast<int64_t>(bb) * x.N * 16 + q;
            const float4* pa = xb + (kSlice == 1 ? e.y : e.x) * 16;
            t.a[0] = __ldg(pa);
            t.a[1] = __ldg(pa + 8);
            if (kSlice == 2) {
                const float4* pb = xb + e.y * 16;
                t.b[0] = __ldg(pb);
                t.b[1] = __ldg(pb + 8);
            }
        };

        const uint32_t first = c_begin + grp;
        if (first >= c_end) {
            finish<kSlice>(g, side, grp, gtid, q, lane);
            return;
        }
        Rows nxt;
        uint32_t row_n = row, b_n = b;      // row / window of the chunk held in `nxt`
        load_rows(nxt, load_ends(), row, b);
        advance();
        int2 e_n = load_ends();             // pipe ends of the chunk after `nxt` (garbage-free: p is always valid)
        constexpr uint32_t kPer = HeadLive::kStages / tgrad::kGroups;   // stages this group rotates through
        uint32_t use = 0;
        for (uint32_t ch = first; ch < c_end; ch += tgrad::kGroups, ++use) {
            const Rows cur = nxt;
            const bool ok = row_n < M;
            if (ch + tgrad::kGroups < c_end) {
                row_n = row;
                b_n = b;
                load_rows(nxt, e_n, row, b);
                advance();
                e_n = load_ends();
            }
            const float gs = ok ? cur.d * g.scale : 0.f;
            // G: four nibbles -> 0 / 1 floats (a row past M stores zeros: its gate word belongs to row M - 1)
            float4 live[4];
            const uint32_t wd[4] = {cur.w.x, cur.w.y, cur.w.z, cur.w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t nib = ok ? wd[k] >> sh : 0u;
                live[k] = make_float4(nib & 8u ? 1.f : 0.f, nib & 4u ? 1.f : 0.f, nib & 2u ? 1.f : 0.f, nib & 1u ? 1.f : 0.f);
                if (kSlice == 0) {
                    side[k].x += nib & 8u ? gs : 0.f;
                    side[k].y += nib & 4u ? gs : 0.f;
                    side[k].z += nib & 2u ? gs : 0.f;
                    side[k].w += nib & 1u ? gs : 0.f;
                }
            }
            // X: g * feature chunk, split into TF32 hi / lo
            float4 hi[2], lo[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float4 v = cur.a[j];
                if (kSlice == 2) v = make_float4(fabsf(v.x - cur.b[j].x), fabsf(v.y - cur.b[j].y), fabsf(v.z - cur.b[j].z), fabsf(v.w - cur.b[j].w));
                split4(make_float4(gs * v.x, gs * v.y, gs * v.z, gs * v.w), hi[j], lo[j]);
            }
            const uint32_t st = grp + tgrad::kGroups * (use % kPer);
            uint8_t* g_hi = ring.base + st * ring.stage + off;
            mbar_wait(ring.empty + st, ((use / kPer) & 1) ^ 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(g_hi + k * tgrad::kBlockBytes) = live[k];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                *reinterpret_cast<float4*>(g_hi + ring.x_hi + j * tgrad::kBlockBytes) = hi[j];
                *reinterpret_cast<float4*>(g_hi + ring.x_lo + j * tgrad::kBlockBytes) = lo[j];
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(ring.full + st);
        }
        finish<kSlice>(g, side, grp, gtid, q, lane);
    }
    // a[]: the four lanes of a warp with the same q hold four rows' shares of the same 16 hidden units
    template <int kSlice>
    static __device__ __forceinline__ void finish(const HeadLive& g, float4 (&side)[4], int grp, int gtid, int q, int lane) {
        float4* out = reinterpret_cast<float4*>(g.side_part + (static_cast<size_t>(blockIdx.x) * kSideParts +
                                                               grp * (tgrad::kLoaderWarps / tgrad::kGroups) + (gtid >> 5)) * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 t = side[k];
            if (kSlice == 0) {
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                    t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                    t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
                }
            }
            if (lane < 8) out[8 * k + q] = t;      // hidden units 32 k + 4 q .. + 3 (zeros from the other slices' CTAs)
        }
    }
    static __device__ __forceinline__ void run(const HeadLive& g, const HeadFeatSlice& x, const tgrad::Ring& ring, int gtid,
                                               int grp, uint32_t c_begin, uint32_t c_end, uint32_t M) {
        const int slice = blockIdx.x % kHeadSlices;
        if (slice == 0) loop<0>(g, x, ring, gtid, grp, c_begin, c_end, M);
        else if (slice == 1) loop<1>(g, x, ring, gtid, grp, c_begin, c_end, M);
        else loop<2>(g, x, ring, gtid, grp, c_begin, c_end, M);
    }
};

// dW1[j, :] = w2[j] * A[j, :], db1[j] = w2[j] * a[j], dw2[j] = W1[j, :] . A[j, :] + b1[j] * a[j]; one warp per hidden unit j
__global__ void __launch_bounds__(128)
head_param_epilogue_kernel(const float* __restrict__ A, const float* __restrict__ a, const float* __restrict__ W1,
                           const float* __restrict__ b1, const float* __restrict__ w2, float* __restrict__ dW1,
                           float* __restrict__ db1, float* __restrict__ dw2, int H, int F) {
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= H) return;
    const float wj = __ldg(w2 + j);
    float dot = 0.f;
    for (int c = lane; c < F; c += 32) {
        const float v = A[j * F + c];
        dW1[j * F + c] = wj * v;
        dot = fmaf(__ldg(W1 + j * F + c), v, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        const float aj = a[j];
        db1[j] = wj * aj;
        dw2[j] = fmaf(__ldg(b1 + j), aj, dot);
    }
}

}  // namespace

extern "C" int64_t ltgnn_pipe_head_ws_floats(int device) {
    const DeviceInfo* di = device_info(device);  // tgrad partials [sm][128][64] + partials of a[] [16 sm][128] + A [128][192] + a [128]
    return di ? static_cast<int64_t>(di->sm_count) * (tgrad::kMo * 64 + kSideParts * 128) + 128 * 192 + 128 : -1;
}

extern "C" int64_t ltgnn_tgrad_ws_floats(int device, int32_t No) {
    const DeviceInfo* di = device_info(device);
    return di ? static_cast<int64_t>(di->sm_count) * tgrad::kMo * No : -1;
}

// dW[Do, Di] = G[M, Do]^T X[M, Di];  (Do, Di) in {(64, 64), (128, 128), (128, 64), (64, 128), ...}: Do in {64, 128},
// Di a multiple of 32 <= 256.  ws: ltgnn_tgrad_ws_floats(device, Di) floats.
extern "C" int ltgnn_wgrad_tc(int device, int64_t M, int32_t Do, int32_t Di, const float* G, const float* X, float* dW,
                              int accumulate, float* ws, void* stream_) {
    LTGNN_REQUIRE(M >= 0, LTGNN_E_ARG, "wgrad_tc: M=%lld", static_cast<long long>(M));
    LTGNN_REQUIRE(Do == 64 || Do == 128, LTGNN_E_SHAPE, "wgrad_tc: Do=%d must be 64 or 128", Do);
    LTGNN_REQUIRE(Di % 32 == 0 && Di > 0 && Di <= 192, LTGNN_E_SHAPE, "wgrad_tc: Di=%d must be a multiple of 32, <= 192", Di);
    LTGNN_REQUIRE(G && X && dW && ws, LTGNN_E_ARG, "wgrad_tc: null tensor");
    LTGNN_REQUIRE(aligned16(G) && aligned16(X), LTGNN_E_ALIGN, "wgrad_tc: G/X must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (M == 0) {
        if (!accumulate) LTGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * Do * Di, stream));
        return LTGNN_OK;
    }
    // Do = 64 fills only half of the 128 accumulator rows: the other half is fed with zeros
    StackedRows g{reinterpret_cast<const float4*>(G), nullptr, Do / 4, 0};
    StackedRows x{reinterpret_cast<const float4*>(X), nullptr, Di / 4, 0};
    int grid = 0;
    // the GCN shape (64 x 64) takes the narrow variant: half of operand G is never stored, loads are pipelined
    int rc = (Do == 64 && Di == 64) ? tgrad::launch<1, 2, 2, -1>(device, g, x, ws, M, Di, &grid, stream, "wgrad_tc")
                                    : tgrad::launch(device, g, x, ws, M, Di, &grid, stream, "wgrad_tc");
    if (rc) return rc;
    return tgrad::gather(ws, grid, Di, 0, Do, 0, Di, dW, Di, accumulate, stream);
}

// Pipe-head parameter gradients: dW1 [128, 192], db1 [128], dw2 [128] from the 1-bit gate of the hidden layer, dlogit and
// the node states (see the comment above the loaders).  ws: ltgnn_pipe_head_ws_floats(device) floats.
extern "C" int ltgnn_pipe_head_bwd_w(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                                     const int32_t* ends, const float* W1, const float* b1, const float* w2,
                                     const uint32_t* hmask, const float* dlogit, float gate_scale, float* dW1, float* db1,
                                     float* dw2, float* ws, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P > 0, LTGNN_E_ARG, "pipe_head_bwd_w: B=%lld N=%d P=%d", static_cast<long long>(B), N, P);
    LTGNN_REQUIRE(D == 64 && H == 128, LTGNN_E_SHAPE, "pipe_head_bwd_w: D=%d H=%d (64 / 128 only)", D, H);
    LTGNN_REQUIRE(X && ends && W1 && b1 && w2 && hmask && dlogit && dW1 && db1 && dw2 && ws, LTGNN_E_ARG,
                  "pipe_head_bwd_w: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(hmask) && aligned16(ws), LTGNN_E_ALIGN, "pipe_head_bwd_w: alignment");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (B == 0) {
        LTGNN_CUDA_TRY(cudaMemsetAsync(dW1, 0, sizeof(float) * H * 3 * D, stream));
        LTGNN_CUDA_TRY(cudaMemsetAsync(db1, 0, sizeof(float) * H, stream));
        LTGNN_CUDA_TRY(cudaMemsetAsync(dw2, 0, sizeof(float) * H, stream));
        return LTGNN_OK;
    }
    const int64_t M = B * P;
    const DeviceInfo* di = device_info(device);
    LTGNN_REQUIRE(di, LTGNN_E_CUDA, "pipe_head_bwd_w: device %d", device);
    const int No = D;  // one 64-column slice of the 192 feature columns per CTA
    float* part = ws + static_cast<size_t>(di->sm_count) * tgrad::kMo * No;  // [16 sm][128] partials of a[]
    float* A = part + static_cast<size_t>(di->sm_count) * kSideParts * 128;  // [128][192]
    float* a = A + 128 * 192;                                                // [128]
    HeadLive g{reinterpret_cast<const uint4*>(hmask), dlogit, gate_scale, part};
    HeadFeatSlice x{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends), static_cast<uint32_t>(P),
                    static_cast<uint32_t>(N), P >= 2 ? (~0ull / static_cast<uint64_t>(P)) + 1 : 0ull};
    int grid = 0;
    int rc = tgrad::launch<1, 4, 2, -4>(device, g, x, ws, M, No, &grid, stream, "pipe_head_bwd_w");
    if (rc) return rc;
    for (int sl = 0; sl < kHeadSlices; ++sl) {  // A[:, 64 sl : 64 sl + 64] = sum over the row ranges of slice sl
        rc = tgrad::gather(ws, grid, No, 0, H, 0, No, A + sl * No, 3 * D, 0, stream, tgrad::kMo, sl, kHeadSlices);
        if (rc) return rc;
    }
    rc = reduce_parts(part, 128, a, kSideParts * grid, H, 0, stream);
    if (rc) return rc;
    head_param_epilogue_kernel<<<(H + 3) / 4, 128, 0, stream>>>(A, a, W1, b1, w2, dW1, db1, dw2, H, 3 * D);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
