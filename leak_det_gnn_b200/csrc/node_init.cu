// node_init.cu -- node-feature initialisation of the detector and its backward.
//
// Reference (models/detector.py:178-190): h0 = zeros(B, N, d_s); h0[:, sensor_idx] = h_s; mask = sensor
// indicator; x0 = dropout(relu(Linear_{d_s+1 -> D}(cat[h0, mask]))).  That materialises three full
// (B, N, .) tensors and runs a (B*N) x 65 x 64 GEMM in which 632 of 661 rows are all-zero inputs.
// Here: a non-sensor row is the batch-independent constant relu(bias); a sensor row is
// relu(W[:, :d_s] h_s + W[:, d_s] + bias).  HBM traffic = one write of x0 (forward), one read of
// dx0 and x0 (backward); the 29-row GEMMs run on CUDA cores out of shared memory (0.1 % of the bytes).
#include "common.cuh"

using namespace ltgnn;

namespace {

constexpr int kThreads = 256;

struct InitParams {
    const float* hs;        // [B, S, ds]
    const float* W;         // [D, ds + 1]  (torch Linear layout)
    const float* bias;      // [D]
    const int32_t* slot;    // [N]: sensor slot of node i, or -1
    int64_t B;
    int32_t N, S, ds, D;
    uint32_t drop_thresh;
    float keep_scale;
    uint64_t drop_seed;
};

// ---------------------------------------- forward ----------------------------------------
__global__ void __launch_bounds__(kThreads)
node_init_fwd_kernel(const InitParams p, float* __restrict__ X0) {
    extern __shared__ __align__(16) float sm[];
    const int ds1 = p.ds + 1, D = p.D, d4 = D >> 2;
    float* Wt = sm;                   // [ds+1][D]  transposed weight
    float* base = Wt + ds1 * D;       // [D] relu(bias)
    float* cst = base + D;            // [D] W[:, ds] + bias
    float* hs = cst + D;              // [S][ds]
    float* sens = hs + p.S * p.ds;    // [S][D]
    const int tid = threadIdx.x;

    for (int i = tid; i < D * ds1; i += kThreads) {
        const int j = i / ds1, k = i - j * ds1;
        Wt[k * D + j] = __ldg(p.W + i);
    }
    __syncthreads();
    for (int j = tid; j < D; j += kThreads) {
        const float b = __ldg(p.bias + j);
        base[j] = fmaxf(b, 0.f);
        cst[j] = Wt[p.ds * D + j] + b;
    }

    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();  // previous window's sens/hs fully consumed
        const float* hsb = p.hs + b * p.S * p.ds;
        for (int i = tid; i < p.S * p.ds; i += kThreads) hs[i] = __ldg(hsb + i);
        __syncthreads();
        for (int i = tid; i < p.S * D; i += kThreads) {
            const int s = i / D, j = i - s * D;
            float acc = 0.f;
            const float* h = hs + s * p.ds;
#pragma unroll 8
            for (int k = 0; k < p.ds; ++k) acc = fmaf(h[k], Wt[k * D + j], acc);
            sens[i] = fmaxf(acc + cst[j], 0.f);
        }
        __syncthreads();
        float4* out = reinterpret_cast<float4*>(X0) + b * p.N * d4;
        for (int i = tid; i < p.N * d4; i += kThreads) {
            const int r = i / d4, c = i - r * d4;
            const int sl = __ldg(p.slot + r);
            float4 v = sl < 0 ? *reinterpret_cast<const float4*>(base + c * 4)
                              : *reinterpret_cast<const float4*>(sens + sl * D + c * 4);
            if (p.drop_thresh)
                ptx::dropout4(v, static_cast<uint64_t>(b * p.N * d4 + i), p.drop_seed, p.drop_thresh, p.keep_scale);
            ptx::stg_stream(out + i, v);
        }
    }
}

// ---------------------------------------- backward ----------------------------------------
// per CTA partial layout in ws: [D*(ds+1) dW | D db]
__global__ void __launch_bounds__(kThreads)
node_init_bwd_kernel(const InitParams p, const float* __restrict__ dX0, const float* __restrict__ X0, float gate_scale,
                     float* __restrict__ dhs, float* __restrict__ ws) {
    extern __shared__ __align__(16) float sm[];
    const int ds = p.ds, ds1 = ds + 1, D = p.D, d4 = D >> 2;
    float* Wo = sm;                 // [D][ds+1] as stored
    float* hs = Wo + D * ds1;       // [S][ds]
    float* dz = hs + p.S * ds;      // [S][D] gated gradient of the sensor rows
    float* red = dz + p.S * D;      // [kThreads][4] scratch for the db reduction
    const int tid = threadIdx.x;

    for (int i = tid; i < D * ds1; i += kThreads) Wo[i] = __ldg(p.W + i);

    constexpr int kMaxAcc = 40;  // ceil(D*(ds+1)/256): 17 for 64x65, 33 for 128x65
    float wacc[kMaxAcc];
#pragma unroll
    for (int m = 0; m < kMaxAcc; ++m) wacc[m] = 0.f;
    const int n_w = D * ds1;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);  // thread's column group is tid % d4 (kThreads % d4 == 0)

    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        const float* hsb = p.hs + b * p.S * ds;
        for (int i = tid; i < p.S * ds; i += kThreads) hs[i] = __ldg(hsb + i);
        const float4* g4 = reinterpret_cast<const float4*>(dX0) + b * p.N * d4;
        const float4* x4 = reinterpret_cast<const float4*>(X0) + b * p.N * d4;
        for (int i = tid; i < p.N * d4; i += kThreads) {
            const int r = i / d4, c = i - r * d4;
            float4 g = ptx::ldg_stream(g4 + i);
            const float4 x = ptx::ldg_stream(x4 + i);
            g.x = x.x > 0.f ? g.x * gate_scale : 0.f;
            g.y = x.y > 0.f ? g.y * gate_scale : 0.f;
            g.z = x.z > 0.f ? g.z * gate_scale : 0.f;
            g.w = x.w > 0.f ? g.w * gate_scale : 0.f;
            csum.x += g.x; csum.y += g.y; csum.z += g.z; csum.w += g.w;
            const int sl = __ldg(p.slot + r);
            if (sl >= 0) *reinterpret_cast<float4*>(dz + sl * D + c * 4) = g;
        }
        __syncthreads();
        // d h_s[b, s, k] = sum_j dz[s, j] W[j, k]
        float* out = dhs + b * p.S * ds;
        for (int i = tid; i < p.S * ds; i += kThreads) {
            const int s = i / ds, k = i - s * ds;
            float acc = 0.f;
            const float* z = dz + s * D;
#pragma unroll 8
            for (int j = 0; j < D; ++j) acc = fmaf(z[j], Wo[j * ds1 + k], acc);
            out[i] = acc;
        }
        // dW[j, k] += sum_s dz[s, j] * (k < ds ? hs[s, k] : 1)
#pragma unroll
        for (int m = 0; m < kMaxAcc; ++m) {
            const int e = tid + m * kThreads;
            if (e < n_w) {
                const int j = e / ds1, k = e - j * ds1;
                float acc = wacc[m];
                if (k < ds) {
                    for (int s = 0; s < p.S; ++s) acc = fmaf(dz[s * D + j], hs[s * ds + k], acc);
                } else {
                    for (int s = 0; s < p.S; ++s) acc += dz[s * D + j];
                }
                wacc[m] = acc;
            }
        }
    }

    float* part = ws + static_cast<size_t>(blockIdx.x) * (n_w + D);
#pragma unroll
    for (int m = 0; m < kMaxAcc; ++m) {
        const int e = tid + m * kThreads;
        if (e < n_w) part[e] = wacc[m];
    }
    __syncthreads();
    *reinterpret_cast<float4*>(red + tid * 4) = csum;
    __syncthreads();
    if (tid < d4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = tid; k < kThreads; k += d4) {
            const float4 o = *reinterpret_cast<const float4*>(red + k * 4);
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *reinterpret_cast<float4*>(part + n_w + tid * 4) = t;
    }
}

int check_common(int device, int64_t B, int32_t N, int32_t S, int32_t ds, int32_t D, const DeviceInfo** di,
                 const char* who) {
    LTGNN_REQUIRE(B >= 0 && N > 0 && S >= 0 && ds > 0 && D > 0, LTGNN_E_ARG, "%s: B=%lld N=%d S=%d ds=%d D=%d", who,
                  static_cast<long long>(B), N, S, ds, D);
    LTGNN_REQUIRE(D % 4 == 0 && kThreads % (D / 4) == 0, LTGNN_E_SHAPE, "%s: D=%d must be 4 * a divisor of %d", who, D,
                  kThreads);
    LTGNN_REQUIRE(D * (ds + 1) <= 40 * kThreads, LTGNN_E_SHAPE, "%s: D*(ds+1)=%d too large", who, D * (ds + 1));
    *di = device_info(device);
    if (!*di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE((*di)->cc_major == 10, LTGNN_E_UNSUPPORTED, "%s: device is sm_%d%d, need sm_100", who,
                  (*di)->cc_major, (*di)->cc_minor);
    return LTGNN_OK;
}

uint32_t thresh_of(float p) { return p > 0.f ? static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0) : 0u; }

}  // namespace

extern "C" int ltgnn_node_init_fwd(int device, int64_t B, int32_t N, int32_t S, int32_t ds, int32_t D, const float* hs,
                                   const int32_t* slot, const float* W, const float* bias, float drop_p,
                                   uint64_t drop_seed, float* X0, void* stream_) {
    const DeviceInfo* di;
    int rc = check_common(device, B, N, S, ds, D, &di, "node_init_fwd");
    if (rc) return rc;
    LTGNN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LTGNN_E_ARG, "node_init_fwd: dropout p=%f", drop_p);
    if (B == 0) return LTGNN_OK;
    LTGNN_REQUIRE(hs && slot && W && bias && X0, LTGNN_E_ARG, "node_init_fwd: null tensor");
    LTGNN_REQUIRE(aligned16(X0), LTGNN_E_ALIGN, "node_init_fwd: X0 must be 16-byte aligned");
    LTGNN_CUDA_TRY(cudaSetDevice(device));
    InitParams p{hs, W, bias, slot, B, N, S, ds, D, thresh_of(drop_p), 1.f / (1.f - drop_p), drop_seed};
    const size_t smem = sizeof(float) * (static_cast<size_t>(ds + 1) * D + 2 * D + static_cast<size_t>(S) * ds +
                                         static_cast<size_t>(S) * D);
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "node_init_fwd: %zu B of shared memory", smem);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(node_init_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const int64_t cap = static_cast<int64_t>(di->sm_count) * 4;
    const int grid = static_cast<int>(B < cap ? B : cap);
    node_init_fwd_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(p, X0);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

extern "C" int64_t ltgnn_node_init_ws_floats(int device, int32_t ds, int32_t D) {
    const DeviceInfo* di = device_info(device);
    return di ? static_cast<int64_t>(di->sm_count) * 2 * (static_cast<int64_t>(D) * (ds + 1) + D) : -1;
}

extern "C" int ltgnn_node_init_bwd(int device, int64_t B, int32_t N, int32_t S, int32_t ds, int32_t D, const float* hs,
                                   const int32_t* slot, const float* W, const float* dX0, const float* X0,
                                   float gate_scale, float* dhs, float* dW, float* dbias, float* ws, void* stream_) {
    const DeviceInfo* di;
    int rc = check_common(device, B, N, S, ds, D, &di, "node_init_bwd");
    if (rc) return rc;
    LTGNN_REQUIRE(hs && slot && W && dX0 && X0 && dhs && dW && dbias && ws, LTGNN_E_ARG, "node_init_bwd: null tensor");
    LTGNN_REQUIRE(aligned16(dX0) && aligned16(X0), LTGNN_E_ALIGN, "node_init_bwd: dX0/X0 must be 16-byte aligned");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    LTGNN_CUDA_TRY(cudaSetDevice(device));
    const int n_w = D * (ds + 1);
    if (B == 0) {
        LTGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * n_w, stream));
        LTGNN_CUDA_TRY(cudaMemsetAsync(dbias, 0, sizeof(float) * D, stream));
        return LTGNN_OK;
    }
    InitParams p{hs, W, nullptr, slot, B, N, S, ds, D, 0u, 1.f, 0};
    const size_t smem = sizeof(float) * (static_cast<size_t>(n_w) + static_cast<size_t>(S) * ds +
                                         static_cast<size_t>(S) * D + 4 * kThreads);
    LTGNN_REQUIRE(smem <= static_cast<size_t>(di->smem_optin), LTGNN_E_SHAPE, "node_init_bwd: %zu B of shared memory", smem);
    LTGNN_CUDA_TRY(cudaFuncSetAttribute(node_init_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const int64_t cap = static_cast<int64_t>(di->sm_count) * 2;
    const int grid = static_cast<int>(B < cap ? B : cap);
    node_init_bwd_kernel<<<grid, kThreads, smem, stream>>>(p, dX0, X0, gate_scale, dhs, ws);
    LTGNN_CUDA_TRY(cudaGetLastError());
    // ws rows are [dW | db]
    rc = reduce_parts(ws, n_w + D, dW, grid, n_w, 0, stream);
    if (rc) return rc;
    return reduce_parts(ws + n_w, n_w + D, dbias, grid, D, 0, stream);
}
