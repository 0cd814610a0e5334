/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into libltgnn.so).
 *
 * Scalar C restatement of the aggregation step of torch_geometric's GCNConv.propagate
 * as the reference invokes it (models/detector.py:199): for every edge e in list order,
 * out[col[e]] += norm[e] * x[row[e]]  -- one fp32 multiply, one fp32 add, no fusion
 * (compile with -ffp-contract=off), the same two roundings ATen's mul + index_add_ do.
 * Expressed over the CSR the product builds (row = target, entries in edge order, self
 * loop last), which visits each target's messages in exactly that order.
 *
 * Y[b, i, :] = sum_k val[k] * X[b, col[k], :],  k in [rowptr[i], rowptr[i+1])
 */
#include <stdint.h>
#include <stddef.h>

void ltgnn_oracle_spmm_f32(int64_t B, int32_t N, int32_t D, const int32_t* rowptr, const int32_t* col,
                           const float* val, const float* X, float* Y) {
    for (int64_t b = 0; b < B; ++b) {
        const float* xb = X + (size_t)b * N * D;
        float* yb = Y + (size_t)b * N * D;
        for (int32_t i = 0; i < N; ++i) {
            float* y = yb + (size_t)i * D;
            for (int32_t d = 0; d < D; ++d) y[d] = 0.0f;
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                const float w = val[k];
                const float* x = xb + (size_t)col[k] * D;
                for (int32_t d = 0; d < D; ++d) {
                    float m = w * x[d];
                    y[d] = y[d] + m;
                }
            }
        }
    }
}

/* Edge-list form (COO, PyG order): the literal scatter-add.  Used to check that the CSR
 * form above is the same computation. */
void ltgnn_oracle_scatter_add_f32(int64_t n_edges, int32_t n_nodes, int32_t D, const int64_t* row,
                                  const int64_t* col, const float* norm, const float* X, float* Y) {
    for (size_t i = 0; i < (size_t)n_nodes * D; ++i) Y[i] = 0.0f;
    for (int64_t e = 0; e < n_edges; ++e) {
        const float w = norm[e];
        const float* x = X + (size_t)row[e] * D;
        float* y = Y + (size_t)col[e] * D;
        for (int32_t d = 0; d < D; ++d) {
            float m = w * x[d];
            y[d] = y[d] + m;
        }
    }
}
