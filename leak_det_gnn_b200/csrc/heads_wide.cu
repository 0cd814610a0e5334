// heads_wide.cu -- the pipe head for node widths the fused head kernels (heads.cu: D = 64, W1 resident) do not take.
//
// Reference: models/detector.py:76-88,204-211 -- feat = [h_u, h_v, |h_u - h_v|] -> Linear(3D, H) -> ReLU -> Dropout ->
// Linear(H, 1).  At D = 128 (BASELINE configs[4]: 100k nodes, 2 000 class pipes, 16 windows) W1 is 393 KB as TF32 hi + lo
// and cannot be resident, but the head is small there (32 000 pipe rows): the features are materialised once,
// F [3][M][D] (48 MB), the 3D -> H product runs on the tensor cores as a three-tap gathered-row GEMM (tcn.cu: tap t reads
// row t * M + m, weight slice W1[:, t D : (t + 1) D]), and the small streaming pieces around it live here:
//
//   pipe_feat_fwd   F[0][m] = x[b, u_p],  F[1][m] = x[b, v_p],  F[2][m] = |x_u - x_v|            (m = b P + p)
//   head_out_fwd    hd = dropout(h) in place (h = relu(pre) from the GEMM);  part[m] = sum_j hd[m, j] w2[j]
//   head_out_bwd    gq[m, j] = hd[m, j] > 0 ? dlogit[m] scale : 0;  dh = gq * w2;  cs[j] = sum_m gq[m, j] (two-stage
//                   fixed-order reduction, fp64 partials: deterministic)
//   head_wide_finish  dW1, db1, dw2 from T = gq^T F (three ltgnn_wgrad_tc products) and cs
//   pipe_feat_bwd   dx[b, n] = dpooled[b] / N + sum over the pipe ends incident to n of
//                   (end u: dF0[m] + s dF2[m];  end v: dF1[m] - s dF2[m]),  s = sign(x_u - x_v)  -- a gather in the
//                   fixed (pipe, end) order of ops.pipe_incidence: dx is written once, no atomics
// All HBM streaming, 16-byte accesses, grid-stride.
#include "common.cuh"

using namespace ltgnn;

namespace {
namespace hw {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
pipe_feat_fwd_kernel(const float4* __restrict__ X, const int32_t* __restrict__ ends, float4* __restrict__ F, int64_t total,
                     int32_t n, int32_t pcount, int32_t d4) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t m = i / d4;
        const int c = static_cast<int>(i - m * d4);
        const int64_t b = m / pcount;
        const int p = static_cast<int>(m - b * pcount);
        const int2 e = __ldg(reinterpret_cast<const int2*>(ends) + p);
        const float4 xu = __ldg(X + (b * n + e.x) * d4 + c), xv = __ldg(X + (b * n + e.y) * d4 + c);
        ptx::stg_stream(F + i, xu);
        ptx::stg_stream(F + total + i, xv);
        ptx::stg_stream(F + 2 * total + i,
                        make_float4(fabsf(xu.x - xv.x), fabsf(xu.y - xv.y), fabsf(xu.z - xv.z), fabsf(xu.w - xv.w)));
    }
}

// one warp per pipe row; lanes own the float4 columns lane, lane + 32, ...; fixed-order butterfly -> deterministic
__global__ void __launch_bounds__(kThreads)
head_out_fwd_kernel(float4* __restrict__ h, const float4* __restrict__ w2, float* __restrict__ part, int64_t M, int32_t h4,
                    uint32_t thresh16, float keep_scale, uint64_t seed_arg, const uint64_t* seed_src) {
    const uint64_t seed = ptx::launch_seed(seed_arg, seed_src);
    const int lane = threadIdx.x & 31;
    const int64_t warps = static_cast<int64_t>(gridDim.x) * (kThreads / 32);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); m < M; m += warps) {
        float acc = 0.f;
        for (int c = lane; c < h4; c += 32) {
            float4 v = h[m * h4 + c];
            if (thresh16) {
                ptx::dropout4h(v, static_cast<uint64_t>(m * h4 + c), seed, thresh16, keep_scale);
                h[m * h4 + c] = v;
            }
            const float4 w = __ldg(w2 + c);
            acc = fmaf(v.x, w.x, acc); acc = fmaf(v.y, w.y, acc); acc = fmaf(v.z, w.z, acc); acc = fmaf(v.w, w.w, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) part[m] = acc;
    }
}

// thread = (row phase, float4 column); a CTA walks its contiguous share of the rows in a fixed order and leaves one partial
// row cs (H) in ws; ltgnn::reduce_parts adds the partial rows in CTA order.  gq = gate * dlogit * scale is the GEMM operand
// of the weight gradient (T = gq^T F), dh = gq * w2 that of the feature gradient.
__global__ void __launch_bounds__(kThreads)
head_out_bwd_kernel(const float4* __restrict__ hd, const float4* __restrict__ w2, const float* __restrict__ dlogit,
                    float scale, float4* __restrict__ dh, float4* __restrict__ gq, float* __restrict__ ws, int64_t M,
                    int32_t h4, int64_t rows_per_cta) {
    __shared__ double red[kThreads][4];
    const int c = threadIdx.x % h4, phase = threadIdx.x / h4, phases = kThreads / h4;  // h4 divides kThreads (host check)
    const int64_t r_begin = blockIdx.x * rows_per_cta, r_end = min(M, r_begin + rows_per_cta);
    const float4 w = __ldg(w2 + c);
    double cs[4] = {0., 0., 0., 0.};   // a cross-entropy gradient summed over all pipe rows: terms ~100x the result
    for (int64_t m = r_begin + phase; m < r_end; m += phases) {
        const float4 v = ptx::ldg_stream(hd + m * h4 + c);
        const float gs = __ldg(dlogit + m) * scale;
        const float4 g = make_float4(v.x > 0.f ? gs : 0.f, v.y > 0.f ? gs : 0.f, v.z > 0.f ? gs : 0.f, v.w > 0.f ? gs : 0.f);
        ptx::stg_stream(gq + m * h4 + c, g);
        ptx::stg_stream(dh + m * h4 + c, make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w));
        cs[0] += g.x; cs[1] += g.y; cs[2] += g.z; cs[3] += g.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = cs[j];
    __syncthreads();
    if (phase == 0) {
        for (int q = 1; q < phases; ++q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) cs[j] += red[q * h4 + c][j];
        }
        reinterpret_cast<float4*>(ws + static_cast<size_t>(blockIdx.x) * 4 * h4)[c] =
            make_float4(static_cast<float>(cs[0]), static_cast<float>(cs[1]), static_cast<float>(cs[2]), static_cast<float>(cs[3]));
    }
}

// The parameter gradients of the head from T[t][j][k] = sum_m gq[m, j] F[t][m, k] (three tensor-core weight-gradient GEMMs)
// and cs[j] = sum_m gq[m, j]:   dW1[j, t D + k] = w2[j] T[t][j][k];   db1[j] = w2[j] cs[j];
// dw2[j] = sum_m dlogit[m] hidden[m, j] = b1[j] cs[j] + sum_{t,k} W1[j, t D + k] T[t][j][k]  (hidden is linear in W1, b1 under
// the gate) -- the hidden activations of the forward GEMM are not summed: their tensor-core rounding bias does not cancel in
// this ill-conditioned sum.  One CTA per hidden unit j, fixed-order tree: deterministic.
__global__ void __launch_bounds__(128)
head_wide_finish_kernel(const float* __restrict__ T, const float* __restrict__ W1, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ cs, float* __restrict__ dW1,
                        float* __restrict__ db1, float* __restrict__ dw2, int H, int D) {
    __shared__ double red[128];
    const int j = blockIdx.x, tid = threadIdx.x;
    const float wj = __ldg(w2 + j);
    double acc = 0.;
    for (int i = tid; i < 3 * D; i += 128) {
        const int t = i / D, k = i - t * D;
        const float tv = __ldg(T + (static_cast<size_t>(t) * H + j) * D + k);
        dW1[static_cast<size_t>(j) * 3 * D + i] = wj * tv;
        acc += static_cast<double>(__ldg(W1 + static_cast<size_t>(j) * 3 * D + i)) * tv;
    }
    red[tid] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        const float c = __ldg(cs + j);
        db1[j] = wj * c;
        dw2[j] = static_cast<float>(red[0] + static_cast<double>(__ldg(b1 + j)) * c);
    }
}

__device__ __forceinline__ float sgn(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(kThreads)
pipe_feat_bwd_kernel(const float4* __restrict__ X, const int32_t* __restrict__ ends, const int32_t* __restrict__ inc_ptr,
                     const int32_t* __restrict__ inc, const float4* __restrict__ dF, const float4* __restrict__ dpooled,
                     float4* __restrict__ dX, int64_t total, int64_t f_stride, int32_t n, int32_t pcount, int32_t d4) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const float inv_n = 1.f / static_cast<float>(n);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / d4;
        const int c = static_cast<int>(i - row * d4);
        const int64_t b = row / n;
        const int node = static_cast<int>(row - b * n);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dpooled) {
            const float4 g = __ldg(dpooled + b * d4 + c);
            acc = make_float4(g.x * inv_n, g.y * inv_n, g.z * inv_n, g.w * inv_n);
        }
        const int k_end = __ldg(inc_ptr + node + 1);
        for (int k = __ldg(inc_ptr + node); k < k_end; ++k) {
            const int e = __ldg(inc + k), p = e >> 1, is_v = e & 1;
            const int2 uv = __ldg(reinterpret_cast<const int2*>(ends) + p);
            const int64_t m = b * pcount + p;
            const float4 xu = __ldg(X + (b * n + uv.x) * d4 + c), xv = __ldg(X + (b * n + uv.y) * d4 + c);
            const float4 gd = __ldg(dF + 2 * f_stride + m * d4 + c);
            const float4 ge = __ldg(dF + (is_v ? f_stride : 0) + m * d4 + c);
            const float o = is_v ? -1.f : 1.f;
            acc.x += ge.x + o * sgn(xu.x - xv.x) * gd.x;
            acc.y += ge.y + o * sgn(xu.y - xv.y) * gd.y;
            acc.z += ge.z + o * sgn(xu.z - xv.z) * gd.z;
            acc.w += ge.w + o * sgn(xu.w - xv.w) * gd.w;
        }
        ptx::stg_stream(dX + i, acc);
    }
}

int grid_for(int64_t work_items, int sm_count) {
    const int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = static_cast<int64_t>(sm_count) * 8;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace hw
}  // namespace

extern "C" int ltgnn_pipe_feat_fwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, const float* X,
                                   const int32_t* ends, float* F, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P >= 0 && D > 0 && D % 4 == 0, LTGNN_E_ARG, "pipe_feat_fwd: B=%lld N=%d P=%d D=%d",
                  static_cast<long long>(B), N, P, D);
    if (B == 0 || P == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && ends && F, LTGNN_E_ARG, "pipe_feat_fwd: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(F) && (reinterpret_cast<uintptr_t>(ends) & 7u) == 0, LTGNN_E_ALIGN,
                  "pipe_feat_fwd: alignment");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    const int64_t total = B * P * (D / 4);
    hw::pipe_feat_fwd_kernel<<<hw::grid_for(total, di->sm_count), hw::kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(X), ends, reinterpret_cast<float4*>(F), total, N, P, D / 4);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int ltgnn_head_out_fwd(int device, int64_t M, int32_t H, float* h, const float* w2, float drop_p,
                                  uint64_t drop_seed, float* part, void* stream_) {
    LTGNN_REQUIRE(M >= 0 && H > 0 && H % 4 == 0, LTGNN_E_ARG, "head_out_fwd: M=%lld H=%d", static_cast<long long>(M), H);
    LTGNN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LTGNN_E_ARG, "head_out_fwd: dropout p=%f not in [0,1)", drop_p);
    if (M == 0) return LTGNN_OK;
    LTGNN_REQUIRE(h && w2 && part, LTGNN_E_ARG, "head_out_fwd: null tensor");
    LTGNN_REQUIRE(aligned16(h) && aligned16(w2), LTGNN_E_ALIGN, "head_out_fwd: alignment");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    const uint32_t t16 = drop_p > 0.f ? static_cast<uint32_t>(static_cast<double>(drop_p) * 65536.0 + 0.5) : 0u;
    const float keep = 1.f / (1.f - static_cast<float>(t16) / 65536.f);
    hw::head_out_fwd_kernel<<<hw::grid_for(M * 32, di->sm_count), hw::kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<float4*>(h), reinterpret_cast<const float4*>(w2), part, M, H / 4, t16, keep, drop_seed, seed_source());
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int64_t ltgnn_head_out_ws_floats(int device, int32_t H) {
    const DeviceInfo* di = device_info(device);
    return di ? static_cast<int64_t>(di->sm_count) * 4 * H : -1;
}

extern "C" int ltgnn_head_out_bwd(int device, int64_t M, int32_t H, const float* hd, const float* w2, const float* dlogit,
                                  float scale, float* dh, float* gq, float* cs, float* ws, void* stream_) {
    LTGNN_REQUIRE(M >= 0 && H > 0 && H % 4 == 0 && hw::kThreads % (H / 4) == 0, LTGNN_E_SHAPE,
                  "head_out_bwd: M=%lld H=%d (H / 4 must divide %d)", static_cast<long long>(M), H, hw::kThreads);
    LTGNN_REQUIRE(hd && w2 && dlogit && dh && gq && cs && ws, LTGNN_E_ARG, "head_out_bwd: null tensor");
    LTGNN_REQUIRE(aligned16(hd) && aligned16(w2) && aligned16(dh) && aligned16(gq) && aligned16(ws), LTGNN_E_ALIGN,
                  "head_out_bwd: alignment");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int64_t ctas = static_cast<int64_t>(di->sm_count) * 4;
    int64_t rows_per_cta = M > 0 ? (M + ctas - 1) / ctas : 1;
    if (rows_per_cta < 256) rows_per_cta = 256;   // few, long fp64 partial sums rather than many short ones added in fp32
    ctas = M > 0 ? (M + rows_per_cta - 1) / rows_per_cta : 0;
    if (ctas > 0) {
        hw::head_out_bwd_kernel<<<static_cast<int>(ctas), hw::kThreads, 0, stream>>>(
            reinterpret_cast<const float4*>(hd), reinterpret_cast<const float4*>(w2), dlogit, scale,
            reinterpret_cast<float4*>(dh), reinterpret_cast<float4*>(gq), ws, M, H / 4, rows_per_cta);
        LTGNN_CUDA_TRY(cudaGetLastError());
    }
    return reduce_parts(ws, H, cs, static_cast<int>(ctas), H, 0, stream);   // zero parts -> zeros
}

extern "C" int ltgnn_head_wide_finish(int device, int32_t H, int32_t D, const float* T, const float* W1, const float* b1,
                                      const float* w2, const float* cs, float* dW1, float* db1, float* dw2, void* stream_) {
    LTGNN_REQUIRE(H > 0 && D > 0, LTGNN_E_ARG, "head_wide_finish: H=%d D=%d", H, D);
    LTGNN_REQUIRE(T && W1 && b1 && w2 && cs && dW1 && db1 && dw2, LTGNN_E_ARG, "head_wide_finish: null tensor");
    LTGNN_USE_DEVICE(device);
    hw::head_wide_finish_kernel<<<H, 128, 0, static_cast<cudaStream_t>(stream_)>>>(T, W1, b1, w2, cs, dW1, db1, dw2, H, D);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int ltgnn_pipe_feat_bwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, const float* X,
                                   const int32_t* ends, const int32_t* inc_ptr, const int32_t* inc, const float* dF,
                                   const float* dpooled, float* dX, void* stream_) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && P >= 0 && D > 0 && D % 4 == 0, LTGNN_E_ARG, "pipe_feat_bwd: B=%lld N=%d P=%d D=%d",
                  static_cast<long long>(B), N, P, D);
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(X && ends && inc_ptr && inc && dF && dX, LTGNN_E_ARG, "pipe_feat_bwd: null tensor");
    LTGNN_REQUIRE(aligned16(X) && aligned16(dF) && aligned16(dX) && aligned16(dpooled) &&
                      (reinterpret_cast<uintptr_t>(ends) & 7u) == 0,
                  LTGNN_E_ALIGN, "pipe_feat_bwd: alignment");
    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_USE_DEVICE(device);
    const int64_t total = B * N * (D / 4);
    hw::pipe_feat_bwd_kernel<<<hw::grid_for(total, di->sm_count), hw::kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const float4*>(X), ends, inc_ptr, inc, reinterpret_cast<const float4*>(dF),
        reinterpret_cast<const float4*>(dpooled), reinterpret_cast<float4*>(dX), total, B * P * (D / 4), N, P, D / 4);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}
