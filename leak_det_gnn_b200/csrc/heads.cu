// heads.cu -- the two read-out heads of the detector on top of the last GCN layer.
//
//   pipe head  (models/detector.py:76-88, 204-211): for every class pipe p = (u, v)
//              feat = [h_u, h_v, |h_u - h_v|] -> Linear(3D, H) -> ReLU -> Dropout -> Linear(H, 1)
//              The reference gathers h_u / h_v into (B, P, D) tensors, concatenates a (B, P, 3D) feature
//              tensor (2.4 GB at B = 4096, P = 764) and runs cuBLAS on it.  Here the features are formed
//              on the fly by the loader warps of the tcgen05 row-GEMM (rowgemm.cuh) straight from the
//              L2-resident node states; the hidden layer never leaves the SM in the forward except as the
//              saved post-activation the backward needs.
//   mean pool  (models/detector.py:214-215, PyG global_mean_pool over equal-sized graphs).
//
// Both GEMMs run on the tensor-memory-operand skeleton (rowgemm_ts.cuh): the loaders keep one pipe row per
// thread, split it into TF32 hi/lo in registers and tcgen05.st it into TMEM, so shared memory only holds the
// whole weight W1 (192 KB as hi + lo) and every tile is produced exactly once.
#include "rowgemm_ts.cuh"

using namespace ltgnn;

namespace {

// ------------------------------------------------------------------ forward
constexpr int kD = 64;        // node width supported by the fused head
constexpr int kD4 = kD / 4;

// one row = one class pipe of one window; K values [32 kg, 32 kg + 32) of [x_u | x_v | |x_u - x_v|]
struct PipeFeatLoader {
    const float4* x;   // node states [B*N, kD4]
    const int2* ends;  // [P] (u, v)
    uint32_t P, N;
    uint64_t magic;    // fastdiv constant of P (P >= 2), 0 when P == 1
    __device__ __forceinline__ void operator()(uint32_t row, int kg, float (&v)[32]) const {
        const uint32_t b = magic ? ptx::fastdiv(row, magic) : row;
        const int2 e = __ldg(ends + (row - b * P));
        const int seg = kg >> 1, half = kg & 1;  // uniform across the CTA for a ring stage
        const float4* xb = x + static_cast<int64_t>(b) * N * kD4 + half * 8;
        const float4* pu = xb + e.x * kD4;
        const float4* pv = xb + e.y * kD4;
        if (seg < 2) {
            const float4* src = seg == 0 ? pu : pv;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(src + j);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 a = __ldg(pu + j), c = __ldg(pv + j);
                v[4 * j] = fabsf(a.x - c.x); v[4 * j + 1] = fabsf(a.y - c.y);
                v[4 * j + 2] = fabsf(a.z - c.z); v[4 * j + 3] = fabsf(a.w - c.w);
            }
        }
    }
};

// hidden = dropout(relu(acc + b1)); part[row] = sum over the hidden units of hidden * w2
struct HeadFwdEpilogue {
    const float* b1;   // [H]
    const float* w2;   // [H]
    float* part;       // [nvar][M]
    float* hpost;      // [M][H] saved post-activation, or nullptr (inference)
    int64_t M;
    int H, nh;         // hidden units in total / per variant
    uint32_t drop_thresh16;  // round(p * 2^16), 0 = no dropout
    float keep_scale;        // 1 / (1 - drop_thresh16 / 2^16)
    uint64_t drop_seed;
    template <class Pull>
    __device__ __forceinline__ void operator()(uint32_t row, bool valid, int var, Pull&& pull) const {
        float acc = 0.f;
        for (int c0 = 0; c0 < nh; c0 += 16) {
            float v[16];
            pull(c0, v);
            const int col = var * nh + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + col) + j);
                v[4 * j + 0] = fmaxf(v[4 * j + 0] + bb.x, 0.f);
                v[4 * j + 1] = fmaxf(v[4 * j + 1] + bb.y, 0.f);
                v[4 * j + 2] = fmaxf(v[4 * j + 2] + bb.z, 0.f);
                v[4 * j + 3] = fmaxf(v[4 * j + 3] + bb.w, 0.f);
            }
            if (drop_thresh16) {
                const uint64_t i8 = static_cast<uint64_t>(row) * (H >> 3) + (col >> 3);
                ptx::dropout8(v, i8, drop_seed, drop_thresh16, keep_scale);
                ptx::dropout8(v + 8, i8 + 1, drop_seed, drop_thresh16, keep_scale);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (hpost)  // blocked-32 layout (rows padded to 128 by the caller): coalesced across the warp's rows
                    ptx::stg_stream(reinterpret_cast<float4*>(hpost) + ptx::b32(row, (col >> 2) + j, H >> 2),
                                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w2 + col) + j);
                acc = fmaf(v[4 * j + 0], ww.x, acc);
                acc = fmaf(v[4 * j + 1], ww.y, acc);
                acc = fmaf(v[4 * j + 2], ww.z, acc);
                acc = fmaf(v[4 * j + 3], ww.w, acc);
            }
        }
        if (valid) part[static_cast<int64_t>(var) * M + row] = acc;
    }
};

// ------------------------------------------------------------------ backward (input gradient)
// A[row, j] = d loss / d pre[row, j] = dlogit[row] * w2[j] * (hpost[row, j] > 0 ? scale : 0)
struct DpreLoader {
    const float4* hpost;   // blocked-32 [Mp, H]
    const float* dlogit;   // [M]
    const float4* w2;      // [H/4]
    float scale;
    int h4;
    __device__ __forceinline__ void operator()(uint32_t row, int kg, float (&v)[32]) const {
        const float g = __ldg(dlogit + row) * scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = ptx::ldg_stream(hpost + ptx::b32(row, kg * 8 + j, h4));  // blocked-32 layout
            const float4 w = __ldg(w2 + kg * 8 + j);
            v[4 * j] = t.x > 0.f ? g * w.x : 0.f;
            v[4 * j + 1] = t.y > 0.f ? g * w.y : 0.f;
            v[4 * j + 2] = t.z > 0.f ? g * w.z : 0.f;
            v[4 * j + 3] = t.w > 0.f ? g * w.w : 0.f;
        }
    }
};

__device__ __forceinline__ float sgn(float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }

// dfeat[row, 0:3D] is scattered back to the two end nodes: +u for the h_u block, +v for the h_v block,
// +-sign(h_u - h_v) for the |.| block (torch: d|x| = sign(x), 0 at 0).  Several pipes share a node, so the
// adds are fp32 reductions in L2 (red.global.add.v4.f32) -- like the reference's index_add_ in autograd.
struct HeadBwdEpilogue {
    float* dx;          // [B*N, D], pre-filled with the mean-pool gradient
    const float* x;     // node states (for the sign of h_u - h_v)
    const int2* ends;
    uint32_t P, N;
    int D, ncols;       // ncols = feature-gradient columns per variant
    uint64_t magic;
    template <class Pull>
    __device__ __forceinline__ void operator()(uint32_t row, bool valid, int var, Pull&& pull) const {
        uint32_t b = 0;
        int2 e = make_int2(0, 0);
        if (valid) {
            b = magic ? ptx::fastdiv(row, magic) : row;
            e = __ldg(ends + (row - b * P));
        }
        const int64_t ru = (static_cast<int64_t>(b) * N + e.x) * D, rv = (static_cast<int64_t>(b) * N + e.y) * D;
        // columns [0, D) = d/d x_u, [D, 2D) = d/d x_v, [2D, 3D) = d/d |x_u - x_v|: pull the three blocks of a
        // 16-column chunk together so every end node receives ONE reduction per float4 (32 per row, not 64)
        for (int c0 = 0; c0 < D; c0 += 16) {
            float a[16], b[16], c[16];
            pull(c0, a);
            pull(D + c0, b);
            pull(2 * D + c0, c);
            if (!valid) continue;
            const float4* hu = reinterpret_cast<const float4*>(x + ru + c0);
            const float4* hv = reinterpret_cast<const float4*>(x + rv + c0);
            float4* du = reinterpret_cast<float4*>(dx + ru + c0);
            float4* dv = reinterpret_cast<float4*>(dx + rv + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 p = __ldg(hu + j), q = __ldg(hv + j);
                const float s0 = sgn(p.x - q.x) * c[4 * j], s1 = sgn(p.y - q.y) * c[4 * j + 1];
                const float s2 = sgn(p.z - q.z) * c[4 * j + 2], s3 = sgn(p.w - q.w) * c[4 * j + 3];
                atomicAdd(du + j, make_float4(a[4 * j] + s0, a[4 * j + 1] + s1, a[4 * j + 2] + s2, a[4 * j + 3] + s3));
                atomicAdd(dv + j, make_float4(b[4 * j] - s0, b[4 * j + 1] - s1, b[4 * j + 2] - s2, b[4 * j + 3] - s3));
            }
        }
    }
};

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
    PipeFeatLoader ld{reinterpret_cast<const float4*>(X), reinterpret_cast<const int2*>(ends),
                      static_cast<uint32_t>(P), static_cast<uint32_t>(N), magic_of(P)};
    const uint32_t t16 = drop_p > 0.f ? static_cast<uint32_t>(static_cast<double>(drop_p) * 65536.0 + 0.5) : 0u;
    HeadFwdEpilogue ep{b1, w2, part, hpost, M, H, H, t16, 1.f / (1.f - static_cast<float>(t16) / 65536.f), drop_seed};
    return rowgemm_ts::launch(device, ld, ep, W1, 3 * D, 0, M, 3 * D, H, static_cast<cudaStream_t>(stream_),
                              "pipe_head_fwd");
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
    DpreLoader ld{reinterpret_cast<const float4*>(hpost), dlogit, reinterpret_cast<const float4*>(w2), gate_scale, H / 4};
    HeadBwdEpilogue ep{dX, X, reinterpret_cast<const int2*>(ends), static_cast<uint32_t>(P), static_cast<uint32_t>(N),
                       D, 3 * D, magic_of(P)};
    return rowgemm_ts::launch(device, ld, ep, W1, 3 * D, 1, M, H, 3 * D, static_cast<cudaStream_t>(stream_),
                              "pipe_head_bwd_dx");
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
    LTGNN_CUDA_TRY(cudaSetDevice(device));
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
    LTGNN_CUDA_TRY(cudaSetDevice(device));
    const int64_t total4 = B * N * (D / 4);
    int64_t blocks = (total4 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(di->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    pool_bwd_fill_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(dpooled), reinterpret_cast<float4*>(dX), total4, N, D / 4);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
