// graph.cu -- error plumbing and the immutable device-side graph handle.
//
// Replaces the per-forward index work of the reference: _batchify_edge_index
// (models/detector.py:105-114,195-196) and PyG's gcn_norm inside each GCNConv.forward
// (models/detector.py:199).  The handle is created once per model per device.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ltgnn {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

PFN_tmapEncodeTiled tmap_encode_fn() {
    static PFN_tmapEncodeTiled fn = []() -> PFN_tmapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }();
    return fn;
}

const DeviceInfo* device_info(int device) {
    constexpr int kMax = 64;
    static DeviceInfo table[kMax];
    static std::mutex mu;
    if (device < 0 || device >= kMax) {
        fail(LTGNN_E_ARG, "device ordinal %d out of range", device);
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(mu);
    DeviceInfo& d = table[device];
    if (!d.ok) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || device >= ndev) {
            fail(LTGNN_E_CUDA, "no usable CUDA device %d (%s; %d devices visible)", device,
                 e == cudaSuccess ? "ordinal out of range" : cudaGetErrorString(e), ndev);
            return nullptr;
        }
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) {
            fail(LTGNN_E_CUDA, "cudaGetDeviceProperties(%d) -> %s", device, cudaGetErrorString(e));
            return nullptr;
        }
        d.cc_major = prop.major;
        d.cc_minor = prop.minor;
        d.sm_count = prop.multiProcessorCount;
        d.smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
        d.smem_per_sm = static_cast<int>(prop.sharedMemPerMultiprocessor);
        // the few stream-ordered temporaries (cudaMallocAsync: partial sums of the large-graph paths) come from the
        // device's default pool; by default it hands memory back to the driver at every synchronisation and the next
        // allocation takes milliseconds -- keep it cached
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        d.ok = true;
    }
    return &d;
}

namespace {
// one warp per output element: lane l adds parts l, l + 32, ... in order, then a fixed shuffle tree
__global__ void reduce_parts_kernel(const float* __restrict__ parts, int64_t stride, float* __restrict__ out,
                                    int n_parts, int n, int accumulate) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    float t = 0.f;
    for (int c = lane; c < n_parts; c += 32) t += parts[c * stride + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[i] = accumulate ? out[i] + t : t;
}
}  // namespace

int reduce_parts(const float* parts, int64_t stride, float* out, int n_parts, int n, int accumulate,
                 cudaStream_t stream) {
    if (n <= 0) return LTGNN_OK;
    reduce_parts_kernel<<<(n + 3) / 4, 128, 0, stream>>>(parts, stride, out, n_parts, n, accumulate);
    LTGNN_CUDA_TRY(cudaGetLastError());
    return LTGNN_OK;
}

}  // namespace ltgnn

using namespace ltgnn;

extern "C" int ltgnn_version(void) { return LTGNN_VERSION; }

static thread_local const uint64_t* g_seed_src = nullptr;
namespace ltgnn {
const uint64_t* seed_source() { return g_seed_src; }
}  // namespace ltgnn
extern "C" void ltgnn_seed_source(const uint64_t* dev_word) { g_seed_src = dev_word; }

extern "C" size_t ltgnn_last_error(char* buf, size_t cap) {
    size_t n = strlen(g_err);
    if (buf && cap) {
        size_t m = n < cap - 1 ? n : cap - 1;
        memcpy(buf, g_err, m);
        buf[m] = 0;
    }
    return n;
}

static int check_csr(int32_t n, int32_t nnz, const int32_t* rowptr, const int32_t* col, const char* what,
                     int32_t* max_len) {
    LTGNN_REQUIRE(rowptr[0] == 0 && rowptr[n] == nnz, LTGNN_E_ARG, "%s: rowptr[0]=%d rowptr[N]=%d, expected 0 and %d",
                  what, rowptr[0], rowptr[n], nnz);
    int32_t ml = 0;
    for (int32_t i = 0; i < n; ++i) {
        int32_t len = rowptr[i + 1] - rowptr[i];
        LTGNN_REQUIRE(len >= 0, LTGNN_E_ARG, "%s: rowptr not monotone at row %d", what, i);
        if (len > ml) ml = len;
    }
    for (int32_t k = 0; k < nnz; ++k)
        LTGNN_REQUIRE(col[k] >= 0 && col[k] < n, LTGNN_E_ARG, "%s: col[%d]=%d out of range [0,%d)", what, k, col[k], n);
    *max_len = ml;
    return LTGNN_OK;
}

extern "C" int ltgnn_graph_create(int device, int32_t n_nodes, int32_t nnz, const int32_t* rowptr, const int32_t* col,
                                  const float* val, const int32_t* t_rowptr, const int32_t* t_col, const float* t_val,
                                  ltgnn_graph_t* out) {
    LTGNN_REQUIRE(out != nullptr, LTGNN_E_ARG, "graph_create: out is null");
    *out = nullptr;
    LTGNN_REQUIRE(rowptr && col && val && t_rowptr && t_col && t_val, LTGNN_E_ARG, "graph_create: null array");
    LTGNN_REQUIRE(n_nodes > 0 && nnz >= 0, LTGNN_E_ARG, "graph_create: n_nodes=%d nnz=%d", n_nodes, nnz);

    const DeviceInfo* di = device_info(device);
    if (!di) return LTGNN_E_CUDA;
    LTGNN_REQUIRE(di->cc_major == 10, LTGNN_E_UNSUPPORTED,
                  "device %d is sm_%d%d; libltgnn carries sm_100a code only (no fallback path)", device, di->cc_major,
                  di->cc_minor);

    ltgnn_graph g;
    g.device = device;
    g.sm_count = di->sm_count;
    g.smem_optin = di->smem_optin;
    g.n = n_nodes;
    g.nnz = nnz;
    int rc = check_csr(n_nodes, nnz, rowptr, col, "A_hat", &g.max_row_len[0]);
    if (rc) return rc;
    rc = check_csr(n_nodes, nnz, t_rowptr, t_col, "A_hat^T", &g.max_row_len[1]);
    if (rc) return rc;

    LTGNN_USE_DEVICE(device);
    const int32_t* rp[2] = {rowptr, t_rowptr};
    const int32_t* cc[2] = {col, t_col};
    const float* vv[2] = {val, t_val};
    std::vector<int2> packed(static_cast<size_t>(nnz) + 1);
    for (int t = 0; t < 2; ++t) {
        for (int32_t k = 0; k < nnz; ++k) {
            int bits;
            memcpy(&bits, &vv[t][k], 4);
            packed[k] = make_int2(cc[t][k], bits);
        }
        cudaError_t e = cudaMalloc(&g.rowptr[t], sizeof(int32_t) * (static_cast<size_t>(n_nodes) + 1));
        if (e == cudaSuccess) e = cudaMalloc(&g.colval[t], sizeof(int2) * (static_cast<size_t>(nnz) + 1));
        if (e == cudaSuccess)
            e = cudaMemcpy(g.rowptr[t], rp[t], sizeof(int32_t) * (static_cast<size_t>(n_nodes) + 1), cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy(g.colval[t], packed.data(), sizeof(int2) * static_cast<size_t>(nnz), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {  // release whatever was allocated before the failure
            for (int u = 0; u < 2; ++u) {
                cudaFree(g.rowptr[u]);
                cudaFree(g.colval[u]);
            }
            return fail(LTGNN_E_CUDA, "graph_create: %s", cudaGetErrorString(e));
        }
    }
    *out = new ltgnn_graph(g);
    return LTGNN_OK;
}

extern "C" int ltgnn_graph_destroy(ltgnn_graph_t g) {
    if (!g) return LTGNN_OK;
    DeviceGuard guard;
    guard.enter(g->device);
    for (int t = 0; t < 2; ++t) {
        cudaFree(g->rowptr[t]);
        cudaFree(g->colval[t]);
    }
    delete g;
    return LTGNN_OK;
}

extern "C" int ltgnn_graph_info(ltgnn_graph_t g, int32_t* n_nodes, int32_t* nnz, int32_t* device, int32_t* sm_count) {
    LTGNN_REQUIRE(g != nullptr, LTGNN_E_ARG, "graph_info: null handle");
    if (n_nodes) *n_nodes = g->n;
    if (nnz) *nnz = g->nnz;
    if (device) *device = g->device;
    if (sm_count) *sm_count = g->sm_count;
    return LTGNN_OK;
}
