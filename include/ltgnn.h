/* ltgnn.h -- C ABI of libltgnn.so: sm_100a kernels for the pipe-graph message-passing path
 * of Mateng0228/Leak-det-gnn's leak detector.
 *
 * The reference has NO native code and no FFI; the boundary it offers is the Python
 * operator API of torch_geometric (GCNConv / global_mean_pool) called from
 * models/detector.py:170-218.  Each entry point below names the reference lines it
 * replaces.  INTEGRATION.md shows the ctypes binding and the 3-line reference patch.
 *
 * Conventions
 *   - plain C types only; every function returns an ltgnn_status (0 = ok, <0 = error);
 *     ltgnn_last_error() returns the thread's last message.  No exception crosses the ABI.
 *   - the caller owns every tensor: device pointers to fp32, contiguous, row-major,
 *     16-byte aligned.  The library never allocates activations; it owns only the device
 *     copy of the graph held by a handle (immutable after create -> shareable by threads).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous on that stream.  Safe to call from any host thread (torch's autograd
 *     worker calls the *_bwd functions).
 *   - activations are [B, N, D]: B independent windows over the SAME N-node graph
 *     (the reference's B-times replicated disjoint union, detector.py:105-114, kept dense).
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     LTGNN_E_CUDA.
 */
#ifndef LTGNN_H_
#define LTGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTGNN_VERSION 100 /* major*10000 + minor*100 + patch : 0.1.0 */

typedef enum ltgnn_status {
    LTGNN_OK = 0,
    LTGNN_E_ARG = -1,         /* null pointer / negative size / bad flag                 */
    LTGNN_E_SHAPE = -2,       /* shape not supported by the kernels (see each function)  */
    LTGNN_E_ALIGN = -3,       /* pointer not 16-byte aligned                             */
    LTGNN_E_CUDA = -4,        /* CUDA runtime / driver error (message has the detail)    */
    LTGNN_E_UNSUPPORTED = -5  /* device is not sm_100                                    */
} ltgnn_status;

typedef struct ltgnn_graph* ltgnn_graph_t;

int ltgnn_version(void);
/* Copies the calling thread's last error message (NUL terminated) into buf; returns its length. */
size_t ltgnn_last_error(char* buf, size_t cap);

/* Device-resident dropout seed, for CUDA-graph capture of a TRAINING step (the reference draws fresh dropout masks on
 * every call of train_detector.py:310; a captured launch would otherwise replay the drop_seed it was captured with).
 * After ltgnn_seed_source(w), every dropout-bearing launch made FROM THE CALLING THREAD (ltgnn_node_init_fwd,
 * ltgnn_gcn_layer_fwd, ltgnn_spmm_fused, ltgnn_pipe_head_fwd) keys its random stream with drop_seed + *w, where the
 * 64-bit device word *w is read when the kernel runs: bump it between replays (on the stream) and every replay draws
 * new masks.  ltgnn_seed_source(NULL) (the default) restores drop_seed alone.  The word must outlive the launches. */
void ltgnn_seed_source(const uint64_t* dev_word);

/* ---- graph handle -------------------------------------------------------------------
 * Replaces the per-forward work of detector.py:195-196 (_batchify_edge_index) and of
 * torch_geometric gcn_norm inside every GCNConv.forward (detector.py:199): the normalised
 * adjacency A_hat of ONE graph is uploaded once, as CSR (row = message target) and as the
 * CSR of A_hat^T ("CSC transpose", row = message source) used by every backward.
 * Host arrays: rowptr int32[N+1], col int32[nnz], val fp32[nnz]; same for t_*.
 * Row entries are consumed in array order (the summation order is part of the contract).
 */
int ltgnn_graph_create(int device, int32_t n_nodes, int32_t nnz,
                       const int32_t* rowptr, const int32_t* col, const float* val,
                       const int32_t* t_rowptr, const int32_t* t_col, const float* t_val,
                       ltgnn_graph_t* out);
int ltgnn_graph_destroy(ltgnn_graph_t g);
int ltgnn_graph_info(ltgnn_graph_t g, int32_t* n_nodes, int32_t* nnz, int32_t* device, int32_t* sm_count);

/* ---- aggregation (SpMM) ---------------------------------------------------------------
 * Y[b,i,:] = sum_k val[k] * X[b, col[k], :]   over row i of A_hat (transpose=0)
 *                                              or of A_hat^T        (transpose=1).
 * transpose=0 replaces GCNConv.propagate (gather x_j, multiply by norm, scatter_add at the
 * target; detector.py:199) for all B windows; transpose=1 is its adjoint, i.e. the
 * backward of propagate wrt x, as a deterministic gather (no atomics).
 * Each entry is one fp32 multiply then one fp32 add in row order -- bit-identical to a
 * sequential scatter-add over PyG's edge list.
 * D must be a multiple of 4.  X and Y must not alias.
 * algo: LTGNN_SPMM_AUTO picks; the others force a kernel (testing / benchmarking).
 */
enum {
    LTGNN_SPMM_AUTO = 0,
    LTGNN_SPMM_STAGED = 1, /* whole-graph feature slices staged in shared memory by TMA    */
    LTGNN_SPMM_GATHER = 2  /* neighbour rows gathered straight from L2 (any graph size)    */
};
int ltgnn_spmm(ltgnn_graph_t g, int transpose, int64_t B, int32_t D, const float* X, float* Y, int algo,
               void* stream);

/* Fused variant of the STAGED kernel (D % 32 == 0 and the graph slice must fit shared memory,
 * otherwise LTGNN_E_SHAPE and the caller composes the unfused pieces):
 *   input gate   (gate != NULL): X is multiplied by (gate > 0 ? gate_scale : 0) BEFORE aggregation --
 *                the backward of an upstream ReLU(+inverted dropout) whose output is `gate`
 *                (detector.py:200-201); colsum[D] (optional) receives the column sums of the gated X
 *                over all B*N rows, i.e. d loss / d bias of that layer.  `ws` must hold
 *                ltgnn_spmm_ws_floats(g) floats.  Deterministic (fixed reduction order).
 *   output epilogue: Y = dropout(relu(A X + bias)); bias NULL / relu 0 / drop_p 0 switch parts off
 *                (GCNConv's `out + bias`, F.relu, nn.Dropout: detector.py:199-201).  Dropout is
 *                inverted dropout from a counter-based Philox stream keyed by drop_seed:
 *                statistically equivalent to torch's, not bit-identical.
 *   1-bit gates: live_out (optional, with the output epilogue) receives one bit per element of Y,
 *                word [b][D/32][N]: bit 8 c + q of word (b, s, i) = (Y[b, i, 32 s + 4 q + c] > 0), c < 4, q < 8
 *                (the order in which a warp vote delivers them).  live_in (instead of
 *                `gate`) feeds such a tensor to the input gate: the backward then streams 1/32 of the
 *                bytes of the float gate.  Graphs up to 1024 nodes.
 */
int64_t ltgnn_spmm_ws_floats(ltgnn_graph_t g);
int ltgnn_spmm_fused(ltgnn_graph_t g, int transpose, int64_t B, int32_t D, const float* X, float* Y,
                     const float* bias, int relu, float drop_p, uint64_t drop_seed, const float* gate,
                     float gate_scale, float* colsum, float* ws, uint32_t* live_out, const uint32_t* live_in,
                     void* stream);

/* ---- dense row-wise layer on tensor cores --------------------------------------------------
 * Y[M,N] = gate( act( X[M,K] op(W) + bias[N] ) )
 *   op(W) = W^T with W [N,K] row-major (w_transposed = 0, torch.nn.Linear layout)
 *         = W   with W [K,N] row-major (w_transposed = 1; the input-gradient GEMM dX = dY W)
 *   bias may be NULL; relu = 0/1;
 *   gate (NULL or [M,N]): Y *= gate > 0 ? gate_scale : 0 -- the backward of an upstream
 *   ReLU(+inverted dropout) whose OUTPUT is `gate` (detector.py:189-190,200-201).
 * tcgen05 3xTF32 with fp32 accumulation in tensor memory: fp32-faithful (the reference runs these in
 * full fp32: GCNConv.lin detector.py:199; NoLeakHead detector.py:94-99).
 * K multiple of 32 (<= 256), N multiple of 16 (<= 256), operands must fit shared memory.
 */
int ltgnn_linear(int device, int64_t M, int32_t K, int32_t N, const float* X, const float* W, int w_transposed,
                 const float* bias, int relu, const float* gate, float gate_scale, float* Y, void* stream);

/* ---- weight gradient of a row-wise linear map ---------------------------------------------
 * dW[Do,Di] (+)= sum over the M rows of G[row,:]^T X[row,:]   (accumulate = 0/1)
 * Replaces autograd's grad_weight of GCNConv.lin (detector.py:199) over the B*N node rows.
 * (Do/8)*(Di/8) must divide 256 (64x64, 128x128, 64x128, ...).  ws: ltgnn_wgrad_ws_floats() floats.
 * Deterministic: per-CTA partials are added in a fixed order.
 */
int64_t ltgnn_wgrad_ws_floats(int device, int32_t Do, int32_t Di);
int ltgnn_wgrad(int device, int64_t M, int32_t Do, int32_t Di, const float* G, const float* X, float* dW,
                int accumulate, float* ws, void* stream);

/* ---- node-feature initialisation (detector.py:178-190) and its backward ---------------------
 * fwd: X0[b,i,:] = dropout(relu(W[:, :ds] h + W[:, ds] m + bias)), (h, m) = (hs[b, slot[i], :], 1) for a
 *      sensor node (slot[i] >= 0) and (0, 0) otherwise.  hs [B,S,ds]; W [D, ds+1] (sensor_to_node.weight);
 *      slot int32[N] on the device.  Never builds the zero-padded (B,N,ds+1) tensor.
 * bwd: gate = (X0 > 0) * gate_scale applied to dX0; dhs [B,S,ds], dW [D, ds+1], dbias [D].
 *      ws: ltgnn_node_init_ws_floats() floats, 16-byte aligned.  Deterministic.
 *      live_out (fwd, optional, D % 32 == 0) / live_in (bwd, instead of X0): the gate as 1 bit per element,
 *      word [b][D/32][N] as in ltgnn_spmm_fused.
 */
int ltgnn_node_init_fwd(int device, int64_t B, int32_t N, int32_t S, int32_t ds, int32_t D, const float* hs,
                        const int32_t* slot, const float* W, const float* bias, float drop_p, uint64_t drop_seed,
                        float* X0, uint32_t* live_out, void* stream);
int64_t ltgnn_node_init_ws_floats(int device, int64_t B, int32_t S, int32_t ds, int32_t D);
int ltgnn_node_init_bwd(int device, int64_t B, int32_t N, int32_t S, int32_t ds, int32_t D, const float* hs,
                        const int32_t* slot, const float* W, const float* dX0, const float* X0,
                        const uint32_t* live_in, float gate_scale, float* dhs, float* dW, float* dbias, float* ws,
                        void* stream);

/* ---- one GCN layer forward as ONE kernel (detector.py:198-201: conv -> relu -> dropout) ---------
 * Y = dropout(relu(A_hat (X W^T) + bias)) with the dense product on tcgen05 and never written to HBM: the window's
 * X W^T slice goes from tensor memory to shared memory and is aggregated there.  X [B,N,K], W [D,K] (torch Linear
 * layout, GCNConv.lin.weight), bias [D] or null, Y [B,N,D]; dropout / live_out as in ltgnn_spmm_fused.  Bit-identical to
 * ltgnn_linear followed by ltgnn_spmm_fused.  ltgnn_gcn_layer_supported: 1 if the kernel takes (graph, K, D) -- K a
 * multiple of 32 up to 128, D a multiple of 32, at most 896 nodes, the slice of the graph fits shared memory.
 */
int ltgnn_gcn_layer_supported(ltgnn_graph_t g, int32_t K, int32_t D);
int ltgnn_gcn_layer_fwd(ltgnn_graph_t g, int64_t B, int32_t K, int32_t D, const float* X, const float* W,
                        const float* bias, int relu, float drop_p, uint64_t drop_seed, float* Y, uint32_t* live_out,
                        void* stream);

/* ---- read-out heads (detector.py:76-102, 204-216) ---------------------------------------------
 * pipe_head_fwd: for every window b and class pipe p with end nodes ends[p] = (u, v):
 *      hidden = dropout(relu(W1 [x_u, x_v, |x_u - x_v|] + b1)),   W1 [H, 3D] (edge_head.mlp.0.weight)
 *      part[b*P + p] = sum over the H hidden units of hidden * w2
 *   so pipe_logit = part + b2 (edge_head.mlp.3).  hpost (optional; logical [Mp, H], Mp = B*P rounded up to 128,
 *   stored blocked-32 as [Mp/32][H/4][32][4]) receives `hidden` for the backward.  X [B,N,D] node states; ends int32 [P,2] on the device.  D = 64, H = 128.
 *   hmask (optional; uint32 [Mp, H/32]) receives the 1-bit form of hidden > 0: bit 31 - j % 32 of word j / 32 of a row.
 *   hsign (optional; uint32 [Mp, 4]) receives sign(x_u - x_v) as two bits per feature: words 2 g / 2 g + 1 of a row hold
 *   the `> 0` / `< 0` bits of features 32 g .. 32 g + 31 (feature j at bit 31 - j % 32).  The backward needs both.
 * pipe_head_bwd_dx: dX[b, i, :] = dpooled[b, :] / N (the mean-pool gradient; dpooled may be null) + the input gradient
 *   of the pipe head given dlogit [B*P], summed over the pipe ends at node i -- the entries of the incidence lists
 *   inc_ptr int32 [N+1], inc int32 [2P] (entry = pipe << 1 | end, grouped by node) -- in a fixed order (the even and
 *   the odd positions of a node's list as two partial sums): a gather, bit-reproducible, dX is written exactly once.  The node states are not read: hmask and hsign carry everything the forward knew.
 *   ws: ltgnn_pipe_head_dx_ws_floats(device, N, P) floats.
 * mean_pool_fwd / mean_pool_bwd_fill: pooled[b,:] = mean_i X[b,i,:];  dX[b,i,:] = dpooled[b,:] / N.
 */
int ltgnn_pipe_head_fwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                        const int32_t* ends, const float* W1, const float* b1, const float* w2, float drop_p,
                        uint64_t drop_seed, float* part, float* hpost, uint32_t* hmask, uint32_t* hsign, void* stream);
int64_t ltgnn_pipe_head_dx_ws_floats(int device, int32_t N, int32_t P);
int ltgnn_pipe_head_bwd_dx(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const int32_t* inc_ptr,
                           const int32_t* inc, const float* W1, const float* w2, const uint32_t* hmask,
                           const uint32_t* hsign, const float* dlogit, float gate_scale, const float* dpooled, float* ws,
                           float* dX, void* stream);
/* ---- the pipe head for node widths the fused kernels above do not take (csrc/heads_wide.cu) -------------------------
 * Same reference lines (detector.py:76-88,204-211) at D = 128 (BASELINE configs[4]): the features are materialised,
 * F [3][B*P][D] = h_u | h_v | |h_u - h_v|; Linear(3D, H) + ReLU runs as the three-tap gathered-row GEMM ltgnn_tcn_conv
 * (tap t reads row t*B*P + m of F against W1[:, tD:(t+1)D]); these four streaming kernels do the rest.
 *   ltgnn_pipe_feat_fwd   X [B,N,D], ends int32 [P,2] -> F
 *   ltgnn_head_out_fwd    h [M,H] = relu(pre) -> (in place) hd = dropout(h);  part[m] = sum_j hd[m,j] w2[j]
 *   ltgnn_head_out_bwd    gq[m,j] = hd[m,j] > 0 ? dlogit[m] scale : 0;  dh = gq * w2;  cs[j] = sum_m gq[m,j];
 *                         ws: ltgnn_head_out_ws_floats(device, H) floats
 *   ltgnn_head_wide_finish  T [3][H][D] = gq^T F[t] (three ltgnn_wgrad_tc products) -> dW1 [H,3D] = w2[j] T,
 *                         db1 = w2 * cs,  dw2[j] = b1[j] cs[j] + sum W1[j,:] . T[:,j,:]  (the hidden layer is linear in
 *                         W1, b1 under the gate, so dw2 needs no pass over the hidden activations)
 *   ltgnn_pipe_feat_bwd   dX[b,n] = dpooled[b]/N (if given) + sum over the pipe ends at n, in the order of inc
 *                         (inc_ptr int32 [N+1], inc int32 [2P]: pipe << 1 | end), of  end u: dF0[m] + s dF2[m],
 *                         end v: dF1[m] - s dF2[m],  s = sign(x_u - x_v).  dX written once; deterministic.
 */
int ltgnn_pipe_feat_fwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, const float* X, const int32_t* ends,
                        float* F, void* stream);
int ltgnn_head_out_fwd(int device, int64_t M, int32_t H, float* h, const float* w2, float drop_p, uint64_t drop_seed,
                       float* part, void* stream);
int64_t ltgnn_head_out_ws_floats(int device, int32_t H);
int ltgnn_head_out_bwd(int device, int64_t M, int32_t H, const float* hd, const float* w2, const float* dlogit, float scale,
                       float* dh, float* gq, float* cs, float* ws, void* stream);
int ltgnn_head_wide_finish(int device, int32_t H, int32_t D, const float* T, const float* W1, const float* b1,
                           const float* w2, const float* cs, float* dW1, float* db1, float* dw2, void* stream);
int ltgnn_pipe_feat_bwd(int device, int64_t B, int32_t N, int32_t P, int32_t D, const float* X, const int32_t* ends,
                        const int32_t* inc_ptr, const int32_t* inc, const float* dF, const float* dpooled, float* dX,
                        void* stream);

int ltgnn_mean_pool_fwd(int device, int64_t B, int32_t N, int32_t D, const float* X, float* pooled, void* stream);
int ltgnn_mean_pool_bwd_fill(int device, int64_t B, int32_t N, int32_t D, const float* dpooled, float* dX, void* stream);

/* ---- weight gradients on tensor cores (tcgen05, MN-major operands, 3xTF32) ---------------------
 * wgrad_tc: same contract as ltgnn_wgrad for Do in {64, 128}, Di a multiple of 32 (<= 256); ~3x faster.
 * pipe_head_bwd_w: dW1 [H, 3D] = dpre^T [x_u, x_v, |x_u - x_v|] and db1 [H] = column sums of dpre, where
 *   dpre[r, j] = dlogit[r] * w2[j] * (hidden[r, j] > 0 ? gate_scale : 0), with the gate read as 1 bit per unit from
 *   hmask (ltgnn_pipe_head_fwd) and the operands formed on the fly; dw2 [H] = sum_r dlogit[r] * hidden[r, :] is
 *   recovered from the same accumulators (hidden is linear in W1, b1 under the gate), so the forward need not save
 *   the hidden activations.
 * ws: ltgnn_tgrad_ws_floats(device, Di) floats; pipe head: ltgnn_pipe_head_ws_floats(device).  Deterministic.
 */
int64_t ltgnn_tgrad_ws_floats(int device, int32_t No);
int ltgnn_wgrad_tc(int device, int64_t M, int32_t Do, int32_t Di, const float* G, const float* X, float* dW,
                   int accumulate, float* ws, void* stream);
int ltgnn_pipe_head_bwd_w(int device, int64_t B, int32_t N, int32_t P, int32_t D, int32_t H, const float* X,
                          const int32_t* ends, const float* W1, const float* b1, const float* w2, const uint32_t* hmask,
                          const float* dlogit, float gate_scale, float* dW1, float* db1, float* dw2, float* ws,
                          void* stream);
int64_t ltgnn_pipe_head_ws_floats(int device);

/* ---- frozen TCN predictor, one dilated causal convolution over gathered rows ---------
 * Reference: models/predictor.py:17-52 (CausalConv1d -> LayerNorm -> ReLU, residual add), evaluated only on the
 * (window, position) rows the last time step depends on (models/utils.py:169-216 reads nothing else).
 *   Y[m, :] = [res[res_row[m], :] +] relu(LayerNorm(bias + sum_tap X[src[tap * M + m], :] W[tap]^T))
 * X [x_rows, C], src int32 [taps, M] (-1 = zero row: the left padding), W [taps, C, C] (tap 0 = the oldest input;
 * torch's conv weight permuted (2, 0, 1)), bias / gamma / beta [C] (gamma = beta = null: no LayerNorm), res [r_rows, C]
 * with res_row int32 [M] or both null, Y [M, C].  C = 128.  tcgen05, 3xTF32; the weight is split into TF32 hi / lo and laid
 * out as shared-memory operand stages once per call (into ws: ltgnn_tcn_ws_floats(taps) floats) and streamed by bulk copies.
 */
int64_t ltgnn_tcn_ws_floats(int32_t taps);
int ltgnn_tcn_conv(int device, int64_t M, int32_t C, int32_t taps, const float* X, const int32_t* src, const float* W,
                   const float* bias, const float* gamma, const float* beta, float eps, int relu, const float* res,
                   const int32_t* res_row, float* Y, float* ws, void* stream);

/* ---- shared per-sensor GRU encoder (detector.py:28-73) -----------------------
 * Sequence q = b*S + s has input [r[b,t,s], tf[b,t,0..F)] at step t (the reference's cat([rr, tf]) order);
 * weights in torch.nn.GRU layout (gate order r, z, n): w_ih [3H, 1+F], w_hh [3H, H], b_ih, b_hh [3H].  H = 64.
 * gru_fwd:    h_last [B*S, H] = hidden state after step L-1 (h_0 = 0).  For training also pass
 *             hseq (every state, logical [L*Qp, H]) and gates (r, z, n, W_hn h + b_hn; logical [L*Qp, 4H]; with
 *             save_hn = 0 only r, z, n: [L*Qp, 3H], for gru_bwd_dg_hn),
 *             Qp = B*S rounded up to 128, both in the kernels' blocked-32 layout [rows/32][W/4][32][4]; NULL otherwise.
 *             State lives in tensor memory; recurrent + input GEMM fused on tcgen05 (3xTF32).
 * gru_bwd_dg: back-propagation through time given dh_last [B*S, H]: writes dG (blocked-32, [L*Qp, 4H]), the gradient wrt
 *             the four pre-activation groups (r, z, W_hn h + b_hn, W_in x + b_in).
 * gru_bwd_w:  dBfused [4H, 96] = sum_(t,q) dG^T [h_{t-1} | x | tf | 1 | 0]: columns 0..H-1 are d w_hh, H..H+F
 *             d w_ih, H+F+1 the bias gradients (rows: r, z, hn-part, in-part).  ws: ltgnn_gru_ws_floats() floats.
 */
int ltgnn_gru_fwd(int device, int64_t B, int32_t L, int32_t S, int32_t F, int32_t H, const float* r, const float* tf,
                  const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* h_last,
                  float* hseq, float* gates, int32_t save_hn, void* stream);
int ltgnn_gru_bwd_dg(int device, int64_t Q, int32_t L, int32_t H, const float* w_hh, const float* gates,
                     const float* hseq, const float* dh_last, float* dG, void* stream);
/* gru_bwd_dg_hn: the BPTT for gates saved without the fourth group (gru_fwd with save_hn = 0): W_hn h_{t-1} + b_hn is
 * rebuilt per step by one 64-column tensor-core GEMM; 8.75 GB less written and read at L = 288.  Same dG. */
int ltgnn_gru_bwd_dg_hn(int device, int64_t Q, int32_t L, int32_t H, const float* w_hh, const float* b_hh,
                        const float* gates3, const float* hseq, const float* dh_last, float* dG, void* stream);
/* Recompute form of the BPTT: the forward saves only hseq (pass gates = NULL to gru_fwd), the gates are rebuilt per step
 * from h_{t-1} on the tensor cores.  gru_inproj: P [B*L, 3H] = W_ih[:, 1:] tf + b_ih (+ b_hh for the r and z rows), the
 * part of the pre-activations the S sensors of a window share.  gru_bwd_dg_rc: same dG as gru_bwd_dg. */
int ltgnn_gru_inproj(int device, int64_t B, int32_t L, int32_t F, int32_t H, const float* tf, const float* w_ih,
                     const float* b_ih, const float* b_hh, float* P, void* stream);
int ltgnn_gru_bwd_dg_rc(int device, int64_t B, int32_t L, int32_t S, int32_t F, int32_t H, const float* r,
                        const float* w_ih, const float* w_hh, const float* b_hh, const float* P, const float* hseq,
                        const float* dh_last, float* dG, void* stream);
int64_t ltgnn_gru_ws_floats(int device);
int ltgnn_gru_bwd_w(int device, int64_t B, int32_t L, int32_t S, int32_t F, int32_t H, const float* r, const float* tf,
                    const float* hseq, const float* dG, float* dBfused, float* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTGNN_H_ */
