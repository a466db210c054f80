/*
 * xdr.h -- C ABI of libxdr.so: the B200 (sm_100a) hot path of cross-domain recommender training.
 *
 * This is the drop-in boundary for the per-batch path of RecBole-CDR
 *     gather -> cross-domain map/transfer -> score + loss -> sparse gradient scatter-add
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/recbole_cdr/).  The reference itself is pure Python on PyTorch/ATen and has no FFI;
 * the binding a maintainer adds is the ctypes stub in INTEGRATION.md (mirrored by
 * recbole-cdr_b200/recbole_cdr_b200/_lib.py).
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; no C++ or torch types cross the boundary.
 *   - Every data pointer is a DEVICE pointer on the current CUDA device unless the name says "host".
 *     Tables are row-major fp32 [n_rows, dim] (nn.Embedding.weight layout), contiguous, 16-byte aligned;
 *     ids are int64 (torch.LongTensor), labels/scores/losses fp32.  dim % 4 == 0 and dim <= 256.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).  The library
 *     never synchronises, never allocates and never touches the default stream behind the caller's back.
 *   - The caller owns every buffer and keeps it alive until the stream work has completed.
 *   - Return value: 0 = XDR_OK, negative = error; xdr_last_error() returns a thread-local message.
 *   - `oob` (may be NULL): device int32 flag set to 1 by the kernel if any id is outside [0, n_rows)
 *     (such ids are skipped: they contribute zero rows and receive no gradient).  PyTorch raises
 *     IndexError for the same input; the Python wrapper turns the flag into that exception on request.
 *   - `ws`: caller-provided scratch of xdr_workspace_bytes() bytes, zero-filled ONCE at allocation and
 *     private to one stream; kernels leave it zeroed again (self-cleaning tickets).
 *   - Loss reductions are deterministic: per-block partials are summed in a fixed order by the last
 *     block to finish.  Gradient scatter uses fp32 vector atomics (red.global.add.v4.f32), so rows hit
 *     by duplicate ids are exact sums in an unspecified order (reference: index_add, also order-free).
 */
#ifndef XDR_H_
#define XDR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XDR_VERSION 100 /* 0.1.0 */

#define XDR_OK 0
#define XDR_ERR_INVALID (-1)     /* bad argument (null pointer, unsupported dim, negative size) */
#define XDR_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define XDR_ERR_UNSUPPORTED (-3) /* valid request that this build does not implement */

typedef void* xdr_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define XDR_API __attribute__((visibility("default")))
#else
#define XDR_API
#endif

/* pointwise data-loss kinds for xdr_point_fwd / xdr_point_bwd */
#define XDR_LOSS_MSE 0         /* nn.MSELoss on the raw dot score      (EMCDR-MF, emcdr.py:50,116)        */
#define XDR_LOSS_BCE_SIGMOID 1 /* nn.BCELoss on sigmoid(dot score)     (CMF cmf.py:79,94; BiTGCF bitgcf.py:227) */
#define XDR_LOSS_NONE 2        /* no data loss, EmbLoss term only      (BiTGCF ego-row regulariser, bitgcf.py:231-233) */

/* activations for the dense-layer family */
#define XDR_ACT_NONE 0
#define XDR_ACT_RELU 1
#define XDR_ACT_TANH 2
#define XDR_ACT_SIGMOID 3

/* ---- library ------------------------------------------------------------------------------------- */
XDR_API int xdr_version(void);
XDR_API const char* xdr_last_error(void);
/* SM count and compute capability of the current device (host out-pointers). */
XDR_API int xdr_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);
/* Bytes of zero-initialised scratch every reducing entry point needs (constant for a given build). */
XDR_API size_t xdr_workspace_bytes(void);

/* ---- A1: embedding-row gather / gradient scatter-add ------------------------------------------------
 * Replaces torch.nn.Embedding.__call__ (emcdr.py:99-100, conet.py:106-109, dtcdr.py:113-118, cmf.py:53-73,
 * bitgcf.py:221-224) and its backward embedding_dense_backward (index_add into the dense grad).
 * out[k, 0:dim] (row stride out_ld floats) = table[idx[k], :]; bit-exact copy.                          */
XDR_API int xdr_gather_rows(const float* table, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx,
                    float* out, int64_t out_ld, int32_t* oob, xdr_stream_t stream);
/* dst[idx[k], :] += scale * rows[k, 0:dim]   (rows has row stride rows_ld floats).                      */
XDR_API int xdr_scatter_add_rows(float* dst, int64_t n_rows, int dim, const int64_t* idx, int64_t n_idx,
                         const float* rows, int64_t rows_ld, float scale, int32_t* oob, xdr_stream_t stream);
/* out[k,:] = max(table_a[idx[k],:], table_b[idx[k],:])  -- DTCDR's element-wise max combine, dtcdr.py:113-119. */
XDR_API int xdr_gather_max2(const float* table_a, const float* table_b, int64_t n_rows, int dim, const int64_t* idx,
                    int64_t n_idx, float* out, int64_t out_ld, int32_t* oob, xdr_stream_t stream);
/* backward of xdr_gather_max2 (torch.maximum: gradient to the larger operand, split 0.5/0.5 on ties).  */
XDR_API int xdr_scatter_max2_bwd(const float* table_a, const float* table_b, int64_t n_rows, int dim, const int64_t* idx,
                         int64_t n_idx, const float* grad_rows, int64_t grad_ld, float scale, float* dst_a,
                         float* dst_b, int32_t* oob, xdr_stream_t stream);

/* ---- A2/A3: fused gather -> dot score -> BPR (+EmbLoss) -------------------------------------------------
 * Replaces EMCDR.calculate_source_loss / calculate_target_loss, BPR branch (emcdr.py:121-130, 144-153):
 *   loss = -mean(log(gamma + sigmoid(s(u,i+) - s(u,i-)))) + reg_weight * (||Eu[u]||_F + ||Ei[i+]||_F) / B
 * out8[0]=loss, [1]=BPR term, [2]=||Eu[u]||_F, [3]=||Ei[i+]||_F, [4]=EmbLoss value.
 * pos_score/neg_score [B] are written by fwd and consumed by bwd.                                          */
XDR_API int xdr_bpr_fwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                const int64_t* user, const int64_t* pos_item, const int64_t* neg_item, int64_t batch,
                float gamma, float reg_weight, float* pos_score, float* neg_score, float* out8, void* ws,
                int32_t* oob, xdr_stream_t stream);
/* user_dst[u] += scale*g*dL/dEu[u], item_dst[i+-] likewise; g = *grad_loss (device scalar, NULL => 1).
 * dst may be a dense gradient table (scale = 1) or the weight table itself (scale = -lr: fused SGD).       */
XDR_API int xdr_bpr_bwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                const int64_t* user, const int64_t* pos_item, const int64_t* neg_item, int64_t batch,
                float gamma, float reg_weight, const float* pos_score, const float* neg_score,
                const float* out8, const float* grad_loss, float scale, float* user_dst, float* item_dst,
                xdr_stream_t stream);

/* ---- A3/A13/A16: fused gather -> dot score -> pointwise loss (+EmbLoss) ---------------------------------
 * Replaces EMCDR MF branch (emcdr.py:111-120, 134-143), CMF per-domain term (cmf.py:75-98) and the two
 * halves of BiTGCF.calculate_loss (bitgcf.py:221-247).  loss_kind: XDR_LOSS_*.  `label` may be NULL for
 * XDR_LOSS_NONE.  score[B] = raw dot product (pre-sigmoid), written by fwd, read by bwd.
 * out8 as in xdr_bpr_fwd ([1] = data-loss term).                                                          */
XDR_API int xdr_point_fwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                  const int64_t* user, const int64_t* item, const float* label, int64_t batch, int loss_kind,
                  float reg_weight, float* score, float* out8, void* ws, int32_t* oob, xdr_stream_t stream);
XDR_API int xdr_point_bwd(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                  const int64_t* user, const int64_t* item, const float* label, int64_t batch, int loss_kind,
                  float reg_weight, const float* score, const float* out8, const float* grad_loss, float scale,
                  float* user_dst, float* item_dst, xdr_stream_t stream);

/* ---- A4/A7/A14: dense-layer family (the mapping MLP, cross-stitch units, NeuMF towers) ------------------
 * Y[m,n] = act( sum_k X[m,k]*W[n,k] + bias[n] + mask[m] * sum_k X2[m,k]*W2[n,k] )
 * W, W2 are nn.Linear.weight layout [N, K] (W2 = crossparas[l].weight, conet.py:122: x_t @ weight.t()).
 * bias, X2/W2 may be NULL.  mask[m] = (mask_ids[m] < mask_lt) when mask_ids != NULL, else 1
 * (conet.py:113-116).  Replaces nn.Linear/torch.mm/activation in emcdr.py:86-93, conet.py:118-138,
 * recbole MLPLayers (dtcdr.py:61-67).                                                                     */
/* Engine of the three GEMM entry points below: 1 = tcgen05.mma on three bf16 operand planes (bf16x6: fp32-faithful products,
 * fp32 accumulation in tensor memory) for M >= 128, N % 16 == 0 (16..128), K % 16 == 0 (16..256) and 16-byte aligned
 * operands, fp32 FMA otherwise; 2 = tcgen05 only for the calls it measured faster on a B200 (forward and input gradient of
 * wide layers: K >= 192 and N >= 64, i.e. CoNet's layer 0), fp32 FMA for the rest; 0 (default) = fp32 FMA for every shape --
 * on a B200 the two tie per call at the BASELINE model shapes (profiles/r2_dense_engines.jsonl), so tcgen05 is opt-in at
 * the library level; the drop-in CoNet class switches its own calls to engine 1, which wins at its 32768-row launches
 * (profiles/r2_conet_stacked.md).  Returns the previous setting.                                                        */
XDR_API int xdr_set_dense_engine(int engine);
XDR_API int xdr_dense_fwd(const float* X, const float* W, const float* bias, const float* X2, const float* W2,
                  const int64_t* mask_ids, int64_t mask_lt, int act, float* Y, int64_t M, int N, int K,
                  xdr_stream_t stream);
/* dZ = dY * act'(Y) (Y = the layer's activated output); dZ may alias dY.                                   */
XDR_API int xdr_act_bwd(const float* Y, const float* dY, int act, float* dZ, int64_t count, xdr_stream_t stream);
/* dX[m,k] (=|+=) mask[m] * sum_n dZ[m,n]*W[n,k];  accumulate != 0 adds into dX.                            */
XDR_API int xdr_dense_bwd_input(const float* dZ, const float* W, const int64_t* mask_ids, int64_t mask_lt, float* dX,
                        int64_t M, int N, int K, int accumulate, xdr_stream_t stream);
/* dW[n,k] += sum_m mask[m]*dZ[m,n]*X[m,k];  db[n] += sum_m mask[m]*dZ[m,n] (db may be NULL).        */
XDR_API int xdr_dense_bwd_weight(const float* dZ, const float* X, const int64_t* mask_ids, int64_t mask_lt, float* dW,
                         float* db, int64_t M, int N, int K, xdr_stream_t stream);

/* ---- A4: mapping loss tail: MSE between mapped rows and gathered target rows ------------------------------
 * Replaces nn.MSELoss(mapping(Es[idx]), Et[idx]) in EMCDR.calculate_map_loss (emcdr.py:156-168).
 * out8[0] = mean((Y - Et[idx])^2) over n_idx*dim elements.                                                  */
XDR_API int xdr_mse_rows_fwd(const float* Y, const float* tgt_tab, int64_t n_rows, int dim, const int64_t* idx,
                     int64_t n_idx, float* out8, void* ws, int32_t* oob, xdr_stream_t stream);
/* dY = g*2*(Y - Et[idx])/(n_idx*dim);  tgt_dst[idx] += scale * (-dY)  (target embedding is NOT detached).   */
XDR_API int xdr_mse_rows_bwd(const float* Y, const float* tgt_tab, int64_t n_rows, int dim, const int64_t* idx,
                     int64_t n_idx, const float* grad_loss, float scale, float* dY, float* tgt_dst,
                     xdr_stream_t stream);

/* ---- A8/A15: BCE on a logit column (the sigmoid output unit + nn.BCELoss) ---------------------------------
 * prob[m] = sigmoid(logit[m]); out8[0] = mean BCE(prob, label) with torch's log clamp at -100
 * (conet.py:140,196-197; dtcdr.py:121-124,186-187).                                                         */
XDR_API int xdr_bce_logit_fwd(const float* logit, const float* label, int64_t count, float* prob, float* out8, void* ws,
                      xdr_stream_t stream);
/* dlogit[m] = g * (p - y) / max(p*(1-p), 1e-12) * p*(1-p) / count  (BCELoss backward x sigmoid backward).   */
XDR_API int xdr_bce_logit_bwd(const float* prob, const float* label, int64_t count, const float* grad_loss,
                      float* dlogit, xdr_stream_t stream);

/* ---- A8: CoNet's regulariser sum_l ||H_l||_F over the cross-stitch matrices (conet.py:198-201: reg_loss += torch.norm(
 * para.weight) per layer -- reg_weight is read but never applied) ---------------------------------------------------------
 * mats_host / counts_host / dsts_host: HOST arrays of n_mats (1..8) device pointers / element counts.
 * norms[l] = ||H_l||_F, out[0] = sum_l norms[l]; one launch, one CTA, fixed summation order.                          */
XDR_API int xdr_frob_sum_fwd(const float* const* mats_host, const int64_t* counts_host, int n_mats, float* norms, float* out,
                     xdr_stream_t stream);
/* dsts[l][i] = g * H_l[i] / norms[l]  (0 where the norm is 0, as torch.norm's backward); g = *grad_loss (NULL: 1).     */
XDR_API int xdr_frob_sum_bwd(const float* const* mats_host, const int64_t* counts_host, int n_mats, const float* norms,
                     const float* grad_loss, float* const* dsts_host, xdr_stream_t stream);

/* ---- A4 / A14-A15 fused: gather -> small MLP -> loss head -> backward -> scatter-add in one persistent kernel ------------
 * Replaces the whole of EMCDR.calculate_map_loss (emcdr.py:156-168, mapping of emcdr.py:58-64,86-93) and of one
 * DTCDR.neumf_forward + BCELoss term (dtcdr.py:112-125,186-187, recbole MLPLayers) including their backward.
 *   layers: n_layers (1..3) nn.Linear weights W[l] [dims[l+1], dims[l]] / biases b[l] (NULL = none); hidden_act after
 *           every layer but the last; *_host arguments are HOST arrays of device pointers.
 *   in_mode 0: x = Au[idx_u]                                         (dims[0] = dim)
 *   in_mode 1: x = [max(Au[u], Bu[u]) | max(Ai[i], Bi[i])]           (dims[0] = 2*dim)
 *   head 0: loss = mean((y - T[idx_u])^2), T not detached            (dims[n] = dim)
 *   head 1: loss = mean BCE(sigmoid(y), label), prob[] optional      (dims[n] = 1)
 *   backward != 0: also dW[l] += , db[l] += , and scatter-adds scale * g * dL/drow into dAu/dBu/dAi/dBi (torch.maximum
 *   routing) and dT; g = *grad_loss.  backward == 0: loss (and prob) only.
 * Weights live in shared memory (W and W^T), activations of a 32-row tile never leave it, weight gradients accumulate
 * in registers across a CTA's tiles.  xdr_fused_mlp_supported() says whether a layer stack fits.                        */
XDR_API int xdr_fused_mlp_supported(int n_layers, const int* dims_host);
XDR_API int xdr_fused_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                               float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head,
                               const float* Au, const float* Bu, const float* Ai, const float* Bi, const float* T,
                               int64_t n_u, int64_t n_i, int dim, const int64_t* idx_u, const int64_t* idx_i,
                               const float* label, int64_t batch, int backward, const float* grad_loss, float scale,
                               float* dAu, float* dBu, float* dAi, float* dBi, float* dT, float* prob, float* out8,
                               void* ws, int32_t* oob, xdr_stream_t stream);

/* ---- A4 / A14-A15 fused, tensor-core engine ---------------------------------------------------------------------------------
 * Same contract, arguments and results as xdr_fused_mlp_step (EMCDR.calculate_map_loss emcdr.py:156-168; one DTCDR NeuMF
 * term dtcdr.py:112-125,186-187), but every layer product of a 64- or 32-row tile (X W^T, dZ W, dZ^T X) is a 3xTF32
 * mma.sync tile product with fp32 accumulation (agrees with fp32 FMA to ~1e-6 relative).  Restrictions on top of the
 * fp32 engine: every layer input width % 8 == 0, hidden widths % 8 == 0 (a final width < 8, i.e. the NeuMF output unit,
 * runs on the CUDA cores), at most 64 / 64 / 8 weight-gradient 16x8 tiles in layers 0 / 1 / 2.                           */
XDR_API int xdr_tc_mlp_supported(int n_layers, const int* dims_host);
XDR_API int xdr_tc_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                            float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head,
                            const float* Au, const float* Bu, const float* Ai, const float* Bi, const float* T,
                            int64_t n_u, int64_t n_i, int dim, const int64_t* idx_u, const int64_t* idx_i,
                            const float* label, int64_t batch, int backward, const float* grad_loss, float scale,
                            float* dAu, float* dBu, float* dAi, float* dBi, float* dT, float* prob, float* out8,
                            void* ws, int32_t* oob, xdr_stream_t stream);

/* ---- A4 fused, tcgen05 engine (tc5_mlp.cu) -----------------------------------------------------------------------------------
 * The EMCDR map step (in_mode 0, head 0, two layers [D, 128, D], D % 16 == 0, D <= 64) with all six products of a 128-row
 * tile on tcgen05.mma kind::f16 (bf16x3: bf16 hi / lo operand planes, fp32 accumulation in tensor memory, ~2^-16 relative per
 * product) and the weight-gradient accumulators resident in tensor memory over all tiles of a CTA.  Same arguments and
 * results as xdr_fused_mlp_step; anything else is XDR_ERR_INVALID.  Validated on a B200 (round 2); EMCDR's default engine.
 * backward == 2 (this entry point only) is the CORRECTION pass of an eager step: a caller that already ran backward == 1
 * with an upstream gradient of 1 at forward time (so that loss and gradients cost ONE launch) calls it from its autograd
 * backward with the real upstream gradient g (grad_loss, required): it adds (g - 1) x the gradients to the same destinations
 * and returns at once, before allocating anything, when g == 1 -- which is what loss.backward() passes.                       */
XDR_API int xdr_tc5_mlp_supported(int n_layers, const int* dims_host);
XDR_API int xdr_tc5_mlp_step(int n_layers, const int* dims_host, const float* const* W_host, const float* const* b_host,
                             float* const* dW_host, float* const* db_host, int hidden_act, int in_mode, int head,
                             const float* Au, const float* Bu, const float* Ai, const float* Bi, const float* T,
                             int64_t n_u, int64_t n_i, int dim, const int64_t* idx_u, const int64_t* idx_i,
                             const float* label, int64_t batch, int backward, const float* grad_loss, float scale,
                             float* dAu, float* dBu, float* dAi, float* dBi, float* dT, float* prob, float* out8,
                             void* ws, int32_t* oob, xdr_stream_t stream);

/* ---- A7-A8 fused: one CoNet tower pass (cross-stitch stack + BCE) with backward and scatter-add in ONE kernel --------------
 * Replaces CoNet.source_forward / target_forward + nn.BCELoss and their autograd backward for one domain batch
 * (conet.py:105-181, 196-197):  x_s = [Su[u] | Si[i]], x_t = [Tu[u] | Ti[i]];  per layer l
 *     x_s' = relu(Ws_l x_s + bs_l + m (H_l x_t)),  x_t' = relu(Wt_l x_t + bt_l + m (H_l x_s)),  m = (id < n_overlap)
 * (id = user, or item when mask_on_item; H_l = crossparas[l].weight shared by both directions), then
 *     loss = mean BCE(sigmoid(w_out . x_want + b_out), label),  want 0 = source tower, 1 = target tower.
 *   dims[0] = 2*dim, dims[1..n_layers] = mlp_hidden_size; *_host are HOST arrays of n_layers device pointers
 *   (weights [dims[l+1], dims[l]], 16-byte aligned); w_out [dims[n_layers]], b_out [1].
 *   backward != 0: dWs/dbs/dWt/dbt/dH/dw_out/db_out += (the unwanted tower's last layer receives zeros), and
 *   scale * g * dL/drow is scatter-added into dSu/dSi/dTu/dTi; g = *grad_loss.  dz1_scratch: xdr_tc_conet_scratch_bytes().
 * Layer products are 3xTF32 mma.sync tiles (fp32-equivalent to ~1e-6).  Supported stacks (xdr_tc_conet_supported):
 * 2*dim % 64 == 0 and <= 512, 1..4 layers, hidden widths % 8 == 0 and <= 64.                                               */
XDR_API int xdr_tc_conet_supported(int n_layers, const int* dims_host, int dim);
XDR_API size_t xdr_tc_conet_scratch_bytes(int64_t batch, int hidden0);
XDR_API int xdr_tc_conet_step(int n_layers, const int* dims_host, const float* const* Ws_host, const float* const* bs_host,
                              const float* const* Wt_host, const float* const* bt_host, const float* const* H_host,
                              float* const* dWs_host, float* const* dbs_host, float* const* dWt_host,
                              float* const* dbt_host, float* const* dH_host, const float* w_out, const float* b_out,
                              float* dw_out, float* db_out, int want, const float* Su, const float* Si, const float* Tu,
                              const float* Ti, int64_t n_u, int64_t n_i, int dim, const int64_t* user, const int64_t* item,
                              const float* label, int64_t batch, int mask_on_item, int64_t n_overlap, int backward,
                              const float* grad_loss, float scale, float* dSu, float* dSi, float* dTu, float* dTi,
                              float* dz1_scratch, float* prob, float* out8, void* ws, int32_t* oob, xdr_stream_t stream);

/* ---- F1: row-sparse optimizer step over the rows a batch touched ------------------------------------------------------------
 * Replaces optimizer.zero_grad() + optimizer.step() of recbole Trainer._train_epoch [recbole-1.0.1] for an embedding table
 * (`learner` sgd / adagrad / sparse_adam of Trainer._build_optimizer): G is a dense [n_rows, dim] gradient table that is zero
 * outside the rows the step kernels scatter-added into; ids [n] are the batch's ids for this table (duplicates allowed).
 * The first visitor of a row in this call (atomicMax on stamp[row] with step_id) updates W and the state rows from the row's
 * summed gradient and writes the gradient row back to zero.  stamp [n_rows] int32 starts at 0; step_id must be >= 1 and
 * strictly larger than in every earlier call on the same stamp table.
 *   XDR_OPT_SGD        w -= lr g                                                (torch.optim.SGD, no momentum / weight decay)
 *   XDR_OPT_ADAGRAD    S1 += g g;  w -= lr g / (sqrt(S1) + eps)                 (torch.optim.Adagrad, lr_decay 0 -- identical
 *                                                                               to the dense optimizer: a zero gradient is a no-op)
 *   XDR_OPT_LAZY_ADAM  S1 += (1-b1)(g-S1);  S2 += (1-b2)(g g-S2);
 *                      w -= lr sqrt(1-b2^t)/(1-b1^t) S1 / (sqrt(S2) + eps)      (torch.optim.SparseAdam; t = adam_t >= 1)
 * beta1/beta2 are doubles: torch forms (1 - beta) and the bias corrections in double before rounding to fp32.               */
#define XDR_OPT_SGD 0
#define XDR_OPT_ADAGRAD 1
#define XDR_OPT_LAZY_ADAM 2
XDR_API int xdr_sparse_optim_rows(int kind, float* W, float* G, float* S1, float* S2, int32_t* stamp, const int64_t* ids,
                                  int64_t n, int64_t n_rows, int dim, int step_id, int64_t adam_t, float lr, float eps,
                                  double beta1, double beta2, int32_t* oob, xdr_stream_t stream);

/* ---- F2: full-sort scoring fused with history masking and top-k ---------------------------------------------------------------
 * Replaces full_sort_predict's dense score matrix (emcdr.py:208-233, cmf.py:107-112: matmul(user_e, all_item_e^T)) together
 * with what recbole's full-sort evaluation does to it next [recbole-1.0.1]: PAD column and each user's history set to -inf,
 * torch.topk(k).  user_vecs [batch, dim]: the user-side vectors (gathered, and mapped where the model maps them);
 * item_tab rows [first_item, n_items) are the candidates (first_item = 1 skips PAD); hist_ptr [batch + 1] / hist_ids: CSR of
 * ASCENDING item ids to exclude per user (both NULL: no masking).  out_score / out_id [batch, k]: score descending, ties by
 * ascending item id; users with fewer than k candidates are padded with (-inf, -1).  The [batch, n_items] matrix is never
 * materialised; scores are 3xTF32 tensor-core dot products (fp32-equivalent).  dim % 8 == 0, k <= 128.                    */
XDR_API size_t xdr_topk_workspace_bytes(int64_t batch, int k);
/* xdr_full_sort_topk_tc5: the same contract with the score blocks on tcgen05.mma (kind::tf32, 3xTF32, accumulators in tensor
 * memory, 128 users per CTA, warp-specialised loader / issuer / epilogue pipeline over mbarriers); dim % 8 == 0, dim <= 64.  */
XDR_API int xdr_full_sort_topk_tc5(const float* user_vecs, int64_t batch, const float* item_tab, int64_t n_items, int dim,
                                   int64_t first_item, const int64_t* hist_ptr, const int64_t* hist_ids, int k,
                                   float* out_score, int64_t* out_id, void* topk_ws, size_t topk_ws_bytes,
                                   xdr_stream_t stream);
/* Self-test of the tcgen05 building blocks (tc5.cuh): D[128, N] = A[128, K] B[N, K]^T in 3xTF32 on one CTA, each operand
 * staged K-major (a_mn = b_mn = 0) in shared memory.  Row-major fp32 device pointers; N % 16 == 0, N <= 256, K % 8 == 0.
 * MN-major TF32 operands (a_mn / b_mn = 1) are refused with XDR_ERR_UNSUPPORTED: on a B200 the SWIZZLE_NONE layouts
 * reproduce the product for K-major 32-bit operands only (profiles/r2_ubench_tcgen05.txt); kind::f16 takes both majors.   */
XDR_API int xdr_tc5_selftest(const float* A, const float* B, int N, int K, int a_mn, int b_mn, float* D, xdr_stream_t stream);
/* The same product as bf16x3 on tcgen05.mma kind::f16 (bf16 hi / lo operand planes, 8 elements per 16-byte chunk, K % 16 == 0):
 * the operand format planned for the tcgen05 training kernels.  ~2^-16 relative per product.  a_mn = 2: A is staged as a
 * row-block-major tile (the physical layout of an MN-major operand) and read through its K-major view (tc5.cuh RowBlock16). */
XDR_API int xdr_tc5_selftest_bf16(const float* A, const float* B, int N, int K, int a_mn, int b_mn, float* D, xdr_stream_t stream);
XDR_API int xdr_full_sort_topk(const float* user_vecs, int64_t batch, const float* item_tab, int64_t n_items, int dim,
                               int64_t first_item, const int64_t* hist_ptr, const int64_t* hist_ids, int k,
                               float* out_score, int64_t* out_id, void* topk_ws, size_t topk_ws_bytes, xdr_stream_t stream);

/* ---- A6: EMCDR predict tail: select mapped vs. target row, then dot ------------------------------------------
 * Replaces the torch.where + mul + sum of EMCDR.predict, OVERLAP/BOTH phase (emcdr.py:191-205):
 *   e = (sel_ids[b] < n_overlap) ? mapped[b, :] : tgt_tab[sel_ids[b], :];   score[b] = e . other_tab[other_ids[b], :]
 * `mapped` [batch, dim] = mapping(Es[sel_ids]) produced by xdr_gather_rows + xdr_dense_fwd.
 * overlap_users: sel = user ids, other = target item table; overlap_items: sel = item ids, other = target user table. */
XDR_API int xdr_select_dot(const float* mapped, const float* tgt_tab, int64_t n_sel_rows, const int64_t* sel_ids,
                   int64_t n_overlap, const float* other_tab, int64_t n_other_rows, const int64_t* other_ids, int dim,
                   int64_t batch, float* score, int32_t* oob, xdr_stream_t stream);

/* ---- A17: the trainer-step hot loop, K batches in ONE persistent software-pipelined launch ---------------------------
 * Replaces K iterations of recbole Trainer._train_epoch's inner loop [recbole-1.0.1] (driven by
 * CrossDomainTrainer.fit, trainer/trainer.py:59-73) around EMCDR.calculate_source_loss / calculate_target_loss
 * (emcdr.py:110-154) or one CMF domain term (cmf.py:75-98):  for k < n_steps: loss_k = L(batch_k); backward.
 * Batch k reads ids user[k*step_stride + 0..batch), item_a[...], item_b[...] (pairwise) / label[...] (pointwise).
 * out8[k*8 + 0..7] as in xdr_bpr_fwd / xdr_point_fwd; per-batch losses and gradient contributions are identical to
 * calling the fwd+bwd pair on batch k.  dst = gradient tables (scale 1): gradients of the K batches accumulate.
 * dst = the weight tables (scale = -lr): asynchronous SGD, a batch may read rows up to 4 steps stale.
 * Restrictions (else XDR_ERR_UNSUPPORTED: use the per-step entry points): batch % 4 == 0, step_stride % 4 == 0,
 * id arrays 16-byte aligned, and at most 40 warp tasks per CTA per step, i.e. ceil(batch / #SMs) <= 160 at dim <= 64
 * (batch <= 23680 on a 148-SM part; rows live in registers, two tasks in flight per warp).
 * steps_ws: xdr_steps_workspace_bytes(n_steps) bytes of scratch.  No initialisation needed: the library zeroes it the
 * first time it sees the pointer; after that the hand-off words carry step tags that grow from launch to launch, so the
 * caller must leave its contents alone between launches (pass a different pointer after writing to it).
 * Launch: all CTAs of these kernels wait for each other, so the grid must be co-resident.  The library checks
 * occupancy x #SMs >= grid before the launch (XDR_ERR_UNSUPPORTED) and launches with cudaLaunchCooperativeKernel (the
 * driver then starts the grid only when every CTA fits, whatever else runs on the device); xdr_set_coop_launch(0)
 * falls back to plain launches (returns the previous setting).                                                       */
XDR_API int xdr_set_coop_launch(int on);
XDR_API size_t xdr_steps_workspace_bytes(int n_steps);
XDR_API int xdr_train_steps(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                            const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                            int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                            float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst,
                            float* out8, void* steps_ws, size_t steps_ws_bytes, int32_t* oob, xdr_stream_t stream);

/* One chunk of K pairwise steps fed from a PINNED HOST id block, enqueued with ONE call (the engine of trainer.FusedStepRunner):
 * on copy_stream -- wait for buf_free_event (the event recorded after the launch that last read dev_ids; NULL: nothing to wait
 * for), copy host_ids [n_steps][3][batch] (user, item+, item-) into dev_ids, record ids_ready_event; on stream -- wait for it,
 * run xdr_train_steps over dev_ids (step stride 3 * batch), copy the [n_steps][8] records into host_out8 (pinned; NULL: no
 * copy) and record launch_done_event (NULL: none).  Replaces, per chunk, the host side of K iterations of recbole
 * Trainer._train_epoch: `interaction.to(device)` and the loss read-back around calculate_loss / backward (reference
 * trainer.py:59-73 via recbole-1.0.1).  The events and streams are the caller's (cudaEvent_t / cudaStream_t handles).       */
XDR_API int xdr_train_steps_host(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                                 const int64_t* host_ids, int64_t* dev_ids, int64_t batch, int n_steps, float gamma,
                                 float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst,
                                 float* out8, float* host_out8, void* steps_ws, size_t steps_ws_bytes, int32_t* oob,
                                 xdr_stream_t copy_stream, void* buf_free_event, void* ids_ready_event,
                                 void* launch_done_event, xdr_stream_t stream);

/* The same launch with LAZILY ZEROED gradient tables (single GPU).  A scatter-add (RED) into a gradient line that is not in
 * L2 costs a DRAM read and, later, a write-back: twice the bytes the reference's `zeros + index_add` needs to produce.
 * touch_map (xdr_touch_map_bytes(n_users, n_items) bytes, 16-byte aligned; user part first) holds 2 bits per destination
 * row.  A row whose bits are clear COUNTS AS ZERO whatever the table holds there: the loader warp that gathers a row also
 * claims its destination row (bit 0), and the first claimer stores a full row of zeros (full-line stores allocate in L2 without
 * reading DRAM), fences and sets bit 1 ("filled"); later claimers wait for bit 1; the scatter-adds then hit L2.  So `optimizer.zero_grad()` (a dense N x D fill in the reference, emcdr.py tables via recbole
 * Trainer._train_epoch) becomes clearing the map -- clear_map != 0 does that first, making user_dst / item_dst the gradient of
 * exactly this launch's K batches on the rows the map marks afterwards (bit 0 of a row's pair), e.g. for a row-sparse
 * optimizer; clear_map == 0 keeps accumulating into the marked rows.  Rows with clear bits are never written and may hold
 * stale data.  The destination tables must not be the weight tables.  Needs the staged kernel (XDR_ERR_UNSUPPORTED
 * otherwise: zero the tables and use xdr_train_steps).  Everything else as xdr_train_steps.                                */
/* Hot rows of the following xdr_train_steps launches (plain scatter-add destinations, single GPU): device arrays of int64 row
 * ids, e.g. the most popular items of the catalogue (item popularity is a property of the dataset: computed once).  Rows that
 * thousands of interactions per step name serialise their REDs in one L2 slice (measured: Zipf(1.05) items, 37 us per step
 * against 3.5 us for uniform ids); every CTA instead adds the gradients of the listed rows in SHARED memory over the whole
 * launch and adds them to the destination once at the end -- the same sums in another order (the reference's index_add is
 * order-dependent at 1e-7 too).  At most 64 rows in total (fewer for rows wider than 64 floats); the arrays must stay valid
 * while launches use them; n_hot_users = n_hot_items = 0 switches it off.                                                */
XDR_API int xdr_steps_set_hot_rows(const int64_t* hot_users, int n_hot_users, const int64_t* hot_items, int n_hot_items);
XDR_API size_t xdr_touch_map_bytes(int64_t n_users, int64_t n_items);
XDR_API int xdr_train_steps_lazy(const float* user_tab, const float* item_tab, int64_t n_users, int64_t n_items, int dim,
                                 const int64_t* user, const int64_t* item_a, const int64_t* item_b, const float* label,
                                 int64_t step_stride, int64_t batch, int n_steps, int pairwise, int loss_kind, float gamma,
                                 float reg_weight, const float* grad_loss, float scale, float* user_dst, float* item_dst,
                                 float* out8, void* steps_ws, size_t steps_ws_bytes, uint32_t* touch_map, int clear_map,
                                 int32_t* oob, xdr_stream_t stream);

/* ---- A10-A12: BiTGCF graph propagate-and-transfer ------------------------------------------------------------------------
 * xdr_spmm_csr: S[r,:] = sum_e val[e] * X[col[e],:] over CSR work items -- replaces torch.sparse.mm(L, E) of
 * BiTGCF.graph_layer (bitgcf.py:131).  Rows are pre-cut into work items of bounded length (heavy-tailed item degrees):
 * work_row/beg/end [n_work], work_split[w] != 0 when the row has several items (those rows, listed in split_rows, are
 * zeroed and accumulated with vector atomics; single-item rows are stored directly).
 * xdr_prop_elementwise: mode 0  out = A + B + A*B        E' = E + S + E*S, bitgcf.py:132-133 (A = E, B = S)
 *                       mode 1  out = A * (1 + B)        backward: dS = dE' * (1 + E)         (A = dE', B = E)
 *                       mode 2  out = A * (1 + B) + C    backward: dE = dE' * (1 + S) + L.dS  (A = dE', B = S, C = L.dS)
 * xdr_transfer_norm_fwd: transfer_layer (bitgcf.py:137-172) on both domains + F.normalize (bitgcf.py:185-186) in one
 * pass: Ps/Pt [N, dim] propagated tables -> Es/Et transferred tables (next layer's input) and Ns/Nt normalised rows
 * (row stride n_ld floats, so they can land in a slot of the layer-concat buffer).  N = n_users + n_items nodes, users
 * first; rows [0, n_ov_users) and [n_users, n_users + n_ov_items) are mixed, all others pass through.  deg_s/deg_t [N]:
 * per-node degree in each domain (bitgcf.py:79-82).  xdr_transfer_norm_bwd is its backward (dEs2/dEt2: gradient that
 * reaches the transferred tables from the next layer, NULL for the last layer).                                       */
XDR_API int xdr_spmm_csr(const int64_t* work_row, const int64_t* work_beg, const int64_t* work_end,
                         const uint8_t* work_split, int64_t n_work, const int64_t* split_rows, int64_t n_split_rows,
                         const int64_t* col, const float* val, const float* X, int dim, float* S, xdr_stream_t stream);
XDR_API int xdr_prop_elementwise(const float* A, const float* B, const float* C, float* out, int64_t count, int mode,
                                 xdr_stream_t stream);
XDR_API int xdr_transfer_norm_fwd(const float* Ps, const float* Pt, int64_t n_users, int64_t n_items, int64_t n_ov_users,
                                  int64_t n_ov_items, int dim, float lam_s, float lam_t, const float* deg_s,
                                  const float* deg_t, float* Es, float* Et, float* Ns, float* Nt, int64_t n_ld,
                                  xdr_stream_t stream);
XDR_API int xdr_transfer_norm_bwd(const float* Es, const float* Et, const float* dNs, const float* dNt, int64_t n_ld,
                                  const float* dEs2, const float* dEt2, int64_t n_users, int64_t n_items,
                                  int64_t n_ov_users, int64_t n_ov_items, int dim, float lam_s, float lam_t,
                                  const float* deg_s, const float* deg_t, float* dPs, float* dPt, xdr_stream_t stream);

/* ---- A18: uniform negative draw with per-user rejection ----------------------------------------------------------------
 * Replaces CrossDomainSourceSampler._uni_sampling + AbstractSampler.sample_by_key_ids (sampler/crossdomain_sampler.py:
 * 220-221, 139-176) and recbole's target-domain Sampler.  out[j*n_keys + p] (j < num) is a valid item id of the domain
 * drawn uniformly and never in used[key_ids[p]] (CSR, columns sorted per row: the get_used_ids sets of :229-250).
 * Candidate k in [0, n_valid) maps to id k+1 if k+1 < n_overlap else k+1+n_gap (source domain: n_gap = n_target_only
 * items; target domain: n_gap = 0, n_overlap = item_num).  Draws are Philox4x32-10(key = seed, counter = position,
 * attempt, stream_id): deterministic and bit-identical to oracle/sampler_oracle.py.  *status |= 1 if some position
 * exhausted max_attempts (a user that used every item: the reference raises ValueError at construction), |= 2 for a key
 * id outside [0, n_rows) (the reference: ValueError('user_id ... not exist')).                                        */
XDR_API int xdr_neg_sample_uniform(const int64_t* key_ids, int64_t n_keys, int num, const int64_t* used_rowptr,
                                   const int64_t* used_col, int64_t n_rows, int64_t n_overlap, int64_t n_gap,
                                   int64_t n_valid, uint64_t seed, uint32_t stream_id, int max_attempts, int64_t* out,
                                   int32_t* status, xdr_stream_t stream);

/* ---- E1: row-sharded tables over peer memory (one process per GPU, NVLink) --------------------------------------------
 * The reference is single-device (no collective anywhere); this is the B200 addition.  A table of
 * n global rows is split block-cyclically over G = n_shards GPUs (G a power of two <= 8): global row r lives on shard
 * r mod G at local row r div G.  Each rank allocates its shard, exports it with xdr_ipc_export, opens the peers' with
 * xdr_ipc_open, and passes all G shard pointers (local and peer-mapped) to xdr_train_steps_sharded: the kernel gathers
 * rows with LDG and scatter-adds gradients with RED directly through the peer mappings, so the NVLink transfers are
 * issued by the same warps that score the batch -- there is no separate all-to-all step and no NCCL call on the data
 * path.  The batch stays data-parallel: every rank runs its own K batches (its own per-batch losses) against the shared
 * sharded tables; accumulated gradients equal a single-GPU run over the union of the batches.
 * n_users / n_items are GLOBAL row counts; *_shards are HOST arrays of G device pointers.                             */
XDR_API int xdr_train_steps_sharded(const float* const* user_shards, const float* const* item_shards,
                                    float* const* user_dst_shards, float* const* item_dst_shards, int n_shards,
                                    int64_t n_users, int64_t n_items, int dim, const int64_t* user, const int64_t* item_a,
                                    const int64_t* item_b, const float* label, int64_t step_stride, int64_t batch,
                                    int n_steps, int pairwise, int loss_kind, float gamma, float reg_weight,
                                    const float* grad_loss, float scale, float* out8, void* steps_ws,
                                    size_t steps_ws_bytes, const float* staged_item_a, const float* staged_item_b,
                                    int32_t* oob, xdr_stream_t stream);
/* Peer gather: out[k,:] = shard[idx[k] mod G][idx[k] div G, :].  Run one chunk ahead of xdr_train_steps_sharded on a
 * second stream, it pulls the chunk's item rows over NVLink into dense local blocks [n_steps][batch][dim] that the
 * persistent kernel then reads sequentially (staged_item_a / staged_item_b; NULL = gather straight from the shards):
 * thousands of row requests in flight hide the NVLink latency that the persistent kernel's registers cannot.          */
XDR_API int xdr_gather_rows_sharded(const float* const* shards, int n_shards, int64_t n_rows, int dim, const int64_t* idx,
                                    int64_t n_idx, int64_t idx_batch, int64_t idx_step_stride, float* out, int64_t out_ld,
                                    int32_t* oob, xdr_stream_t stream);
/* Row-sharded SpMM (BiTGCF over G GPUs): as xdr_spmm_csr, but the operand X is block-cyclically row-sharded
 * over x_shards[0..G) (peer-mapped, each [ceil(N/G), dim]): column id c is row c div G of shard c mod G; the work items
 * cover THIS rank's rows of L (torch.sparse.mm(L, E), bitgcf.py:131, restricted to the rows the rank owns) and S is
 * local.  The neighbour-row gathers cross NVLink directly; the caller orders the launch after every owner has written
 * its shard (one tiny all-reduce on the stream).                                                                      */
XDR_API int xdr_spmm_csr_sharded(const int64_t* work_row, const int64_t* work_beg, const int64_t* work_end,
                                 const uint8_t* work_split, int64_t n_work, const int64_t* split_rows,
                                 int64_t n_split_rows, const int64_t* col, const float* val, const float* const* x_shards,
                                 int n_shards, int dim, float* S, xdr_stream_t stream);
/* CUDA-IPC plumbing.  export: 64-byte handle of the allocation containing dev_ptr + byte offset of dev_ptr inside it
 * (host outputs).  open: map a peer's allocation, returns its base (add the exported offset).  close: unmap.            */
XDR_API int xdr_ipc_export(const void* dev_ptr, unsigned char* handle64_host, int64_t* offset_host);
XDR_API int xdr_ipc_open(const unsigned char* handle64_host, void** base_out_host);
XDR_API int xdr_ipc_close(void* base);

#ifdef __cplusplus
}
#endif
#endif /* XDR_H_ */
