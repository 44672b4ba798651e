/* pfo_b200.h -- C ABI of the B200-native PfoTGNRec hot path (libpfo_b200.so).
 *
 * Drop-in boundary: the reference has no FFI of its own -- its
 * boundary is the Python import surface of model/tgn.py, modules/*.py and utils/utils.py.
 * The host side (pfotgnrec_b200/overlay/*) mirrors those classes and binds the entry points
 * below through ctypes.  Every function:
 *   - takes plain device pointers and sizes (no torch types); the caller owns every buffer;
 *   - is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - returns 0 on success, else the cudaError_t of the failed launch / bad argument;
 *   - keeps no global state and makes no hidden allocation.
 * "file:line" cites the reference code each entry point replaces (paths relative to the
 * reference repository root).
 */
#ifndef PFO_B200_H
#define PFO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFO_ABI_VERSION 3
int pfo_abi_version(void);

/* ---- K1: temporal neighbour sampling --- utils/utils.py:150-161 (find_before) and
 * :163-220 (get_temporal_neighbor); the Python loop over queries with np.searchsorted.
 * CSR adjacency sorted by (node, timestamp, stream order): rowptr[N+1], adj_nbr/adj_eidx/adj_ts[2E].
 * Outputs are [Q, n] row-major, right-aligned / left zero-padded; out_dt = fp32(t - fp32(edge time))
 * (modules/embedding_module.py:133-135).  uniform != 0 selects the uniform-with-replacement mode
 * (:193-204) driven by the shared Philox stream (counter = (query, call, slot, 3)) with
 * call = call_id + *call_ctr (call_ctr may be NULL): the device part lets a captured CUDA graph draw a fresh
 * stream on every replay.  lanes_per_query (most-recent mode): 1, 4, 8 or 32 lanes of a warp share one query's
 * lower-bound search ((lanes+1)-ary, one dependent probe round per factor of lanes+1); 0 = chosen from n_queries.
 * Node-sharded use (queries routed to the rank that owns their CSR row): q_nodes[q] < 0 marks an empty routing slot
 * (its outputs are zero rows); q_ids (may be NULL) = the id of each query in the un-sharded query list, which keys the
 * uniform mode's stream so the draws do not depend on the number of ranks; ld_out (0 = n_neighbors) = row stride of
 * the four outputs, so that they can be column blocks of one reply row; out_etime may be NULL. */
int pfo_neighbor_sample(const int64_t* rowptr, const int32_t* adj_nbr, const int32_t* adj_eidx,
                        const double* adj_ts, const int32_t* q_nodes, const double* q_ts,
                        int64_t n_queries, int n_neighbors, int uniform, uint64_t seed, uint32_t call_id,
                        const uint32_t* call_ctr, int lanes_per_query, const int32_t* q_ids, int64_t ld_out,
                        int32_t* out_nbr, int32_t* out_eidx, float* out_etime, float* out_dt, void* stream);

/* ---- touched-node compaction (replaces the O(n_nodes) clone + Python loop of
 * modules/memory_updater.py:46-51 and modules/message_aggregator.py:40-50: only nodes read
 * in this batch are updated).  mark: set bit `id` for every id >= 0 (ids == 0 are skipped when
 * skip_zero != 0: node 0 is padding).  compact: ascending unique ids, slot_of_node[id] = rank,
 * *n_unique = count; the bitmap is left zeroed. */
int pfo_mark_nodes(const int32_t* ids, int64_t count, int skip_zero, uint32_t* bitmap, void* stream);
int64_t pfo_compact_workspace_ints(int64_t n_nodes);
int pfo_compact_nodes(uint32_t* bitmap, int64_t n_nodes, int32_t* workspace, int32_t* uniq_ids,
                      int32_t* slot_of_node, int32_t* n_unique, void* stream);
int pfo_map_slots(const int32_t* ids, int64_t count, int skip_zero, const int32_t* slot_of_node,
                  int32_t* out, void* stream);

/* ---- dense contractions, exact fp32 --- torch.nn.GRUCell / RNNCell (modules/memory_updater.py:31,47,
 * 60,68), nn.MultiheadAttention projections (model/temporal_attention.py:26-32,70), MergeLayer
 * (utils/utils.py:4-17).  C[m,:N] = epi(alpha * (A[row(m),:K] . W^T + bias * brs[m])), row(m) = a_idx ?
 * a_idx[m] : m (negative -> zero row); W(n,k) = w_transposed ? W[k*ldw+n] : W[n*ldw+k]; act 1 = relu;
 * rows with row_zero[m] != 0 are written as zero; outputs where relu_gate <= 0 are zeroed;
 * accumulate adds into C.  m_dev (optional) holds the live row count on the device. */
int pfo_linear_f32(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                   int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                   float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                   float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                   int accumulate, void* stream);
/* dW[n,k] = sum_m G[m,n] * A[row(m),k]; db[n] = sum_m G[m,n] (db may be null); two-stage,
 * deterministic reduction through `workspace` (pfo_wgrad_workspace_floats floats). */
int64_t pfo_wgrad_workspace_floats(int64_t M, int N, int K, int with_bias);
int pfo_wgrad_f32(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                  int64_t M, const int32_t* m_dev, int N, int K, float* dW, int64_t lddw, float* db,
                  int accumulate, float* workspace, void* stream);
/* ---- the same two contractions on the tcgen05 tensor cores, fed by TMA straight from the fp32 tensors
 * (kind::tf32, accumulators in TMEM).  passes = 3: error-compensated 3xTF32 (a = rna_tf32(a) + residual,
 * three MMAs per K step) -- the default "fp32" mode, 1e-5 contract; passes = 1: plain TF32, the fast mode
 * (2e-2 contract).  Operand layouts TMA cannot describe (a_idx gathers, rows not 16-byte aligned, K beyond
 * one accumulator) are routed to the FFMA kernels above inside the call.  pfo_wgrad_tf32 needs
 * pfo_wgrad_tf32_workspace_floats floats of workspace. */
int pfo_linear_tf32(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                    int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                    float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                    float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                    int accumulate, int passes, void* stream);
int64_t pfo_wgrad_tf32_workspace_floats(int64_t M, int N, int K, int with_bias);
int pfo_wgrad_tf32(const float* G, int64_t ldg, const float* A, int64_t lda, const int32_t* a_idx,
                   int64_t M, const int32_t* m_dev, int N, int K, float* dW, int64_t lddw, float* db,
                   int accumulate, float* workspace, int passes, void* stream);
/* same contract as pfo_linear_f32 with bf16 operands / fp32 accumulation on tcgen05 tensor cores
 * (TMEM accumulators); selected by the host when gemm_mode == "bf16" (2e-2 contract). */
int pfo_linear_bf16(const float* A, int64_t lda, const int32_t* a_idx, const float* W, int64_t ldw,
                    int w_transposed, const float* bias, const float* bias_row_scale, int64_t ld_brs,
                    float* C, int64_t ldc, int64_t M, const int32_t* m_dev, int N, int K,
                    float alpha, int act, const int32_t* row_zero, const float* relu_gate, int64_t ld_gate,
                    int accumulate, void* stream);

/* ---- memory updater gates on the unique touched nodes --- modules/memory_updater.py:35-53
 * (get_updated_memory) restricted to nodes that are read.  cell: 0 GRU, 1 RNN, 2 no memory.
 * gather_state snapshots the rows of the unique nodes (HG memory, XG pending message with row stride ldx >= rawp,
 * valid_u, lu_u = last_update') before persist / store overwrite them; Hcat (optional, row stride ldh) receives a
 * second copy of the memory row -- the tail of the operand row [cell input | memory] of the merged contraction.
 * merged == 0: GI / GH are the input / hidden pre-activations ([U,3d] or [U,d], two GEMMs).  merged != 0: GI is G4 =
 * [r | z | n_i | n_h] ([U,4d], GRU; r and z already summed) or the single pre-activation ([U,d], RNN) of ONE GEMM over
 * [cell input | memory] with the block weight pfo_pack_cell builds from weight_ih / weight_hh / bias_ih / bias_hh
 * (kx = input width, kxp = its padded width in the operand row); dGI is then [U,4d] / [U,d], GH / dGH are unused, and
 * pfo_unpack_cell_grads adds the block weight's gradient back into the four tensors.  Hnew = updated memory rows,
 * H0 = Hnew + node_feat rows (modules/embedding_module.py:93-98). */
int pfo_gather_state(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int raw,
                     const float* memory, const float* pend_msg, int64_t rawp, const uint8_t* pend_valid,
                     const float* pend_ts, const float* last_update,
                     float* HG, float* XG, int64_t ldx, float* Hcat, int64_t ldh, uint8_t* valid_u, float* lu_u,
                     void* stream);
int pfo_cell_forward(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int cell, int merged,
                     const float* GI, const float* GH, int64_t ldg, const float* HG,
                     const uint8_t* valid_u, const float* node_feat, float* Hnew, float* H0, void* stream);
int pfo_cell_backward(const int32_t* uniq, const int32_t* n_uniq, int64_t u_max, int d, int cell, int merged,
                      const float* GI, const float* GH, int64_t ldg, const float* HG,
                      const uint8_t* valid_u, const float* dH, float* dGI, float* dGH, void* stream);
int pfo_pack_cell(const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int d, int kx, int kxp,
                  int cell, float* Wc, float* bc, void* stream);
int pfo_unpack_cell_grads(const float* gWc, const float* gbc, int d, int kx, int kxp, int cell, float* gW_ih,
                          float* gW_hh, float* gb_ih, float* gb_hh, void* stream);

/* ---- persist + message store --- model/tgn.py:185-206 (update_memory for positives, clear,
 * get_raw_messages x2, store_raw_messages), model/tgn.py:357-378, modules/memory.py:35-37,
 * modules/message_aggregator.py:38-55 (`last`).  last_pos[N] is int32 scratch that must hold -1.
 * other_emb_* (optional, [B,d]): embedding of the other endpoint in place of its memory row
 * (use_destination_embedding_in_message, tgn.py:362-365); self_emb_* (optional): own embedding in place of the own
 * memory row (use_source_embedding_in_message, tgn.py:360-361). */
int pfo_persist_rank(const int32_t* src, const int32_t* dst, int B, int d, const int32_t* slot_of_node,
                     const float* Hnew, const uint8_t* pend_valid, const float* pend_ts,
                     float* memory, float* last_update, int32_t* last_pos, void* stream);
int pfo_store_messages(const int32_t* src, const int32_t* dst, const int32_t* eidx, const double* ts,
                       int B, int d, int F, const float* memory, const float* last_update,
                       const float* edge_feat, const float* tw, const float* tb,
                       const float* other_emb_for_src, const float* other_emb_for_dst,
                       const float* self_emb_for_src, const float* self_emb_for_dst,
                       float* pend_msg, int64_t rawp, float* pend_ts, uint8_t* pend_valid,
                       int32_t* last_pos, void* stream);
/* `mean` aggregator --- modules/message_aggregator.py:62-81: pending message = mean of the raw messages the node's
 * interactions of this batch appended (in append order), pending time = that of the last one.  sorted_node / order
 * [2B]: the batch's (node, position = side * B + event) pairs sorted by node with a stable sort. */
int pfo_store_messages_mean(const int32_t* src, const int32_t* dst, const int32_t* eidx, const double* ts,
                            int B, int d, int F, const int32_t* sorted_node, const int32_t* order,
                            const float* memory, const float* last_update,
                            const float* edge_feat, const float* tw, const float* tb,
                            const float* other_emb_for_src, const float* other_emb_for_dst,
                            const float* self_emb_for_src, const float* self_emb_for_dst,
                            float* pend_msg, int64_t rawp, float* pend_ts, uint8_t* pend_valid,
                            int32_t* last_pos, void* stream);

/* node-sharded variants (pfotgnrec_b200/dist.py): build the message rows where the events live (from the
 * unique-node rows fetched from the owners), apply persist + last-wins where the nodes live.  key = global
 * position (side * B_global + event index): the largest key per node wins, independent of the shard count. */
int pfo_build_messages(const int32_t* src_slot, const int32_t* dst_slot, const int32_t* eidx, const double* ts,
                       int B, int d, int F, const float* Hnew, const float* lu_u, const float* edge_feat,
                       const float* tw, const float* tb, const float* other_emb_for_src,
                       const float* other_emb_for_dst, float* rows, int64_t ldr, float* t32_out, void* stream);
int pfo_apply_messages(const int32_t* node, const int32_t* key, int64_t R, int d, int raw,
                       const int32_t* slot_of_node, const float* Hnew, const float* rows, int64_t ldr,
                       const float* t32, float* memory, float* last_update, float* pend_msg, int64_t rawp,
                       float* pend_ts, uint8_t* pend_valid, int32_t* last_pos, void* stream);
/* routed forms (device-side exchange plans): build writes each message into a row of `ldr` >= raw + 3 words followed by
 * [owner-local node id = node / n_ranks | key = key_base + event (+ key_side on the destination side) | fp32 time];
 * apply reads those three words from the received rows and skips empty slots (node id < 0). */
int pfo_build_routed_messages(const int32_t* src_slot, const int32_t* dst_slot, const int32_t* src_node,
                              const int32_t* dst_node, const int32_t* eidx, const double* ts, int B, int d, int F,
                              const float* Hnew, const float* lu_u, const float* edge_feat, const float* tw,
                              const float* tb, const float* other_emb_for_src, const float* other_emb_for_dst,
                              int n_ranks, int key_base, int key_side, float* rows, int64_t ldr, void* stream);
int pfo_apply_routed_messages(const float* rows, int64_t ldr, int64_t R, int d, int raw,
                              const int32_t* slot_of_node, const float* Hnew, float* memory, float* last_update,
                              float* pend_msg, int64_t rawp, float* pend_ts, uint8_t* pend_valid,
                              int32_t* last_pos, void* stream);

/* ---- device-side exchange plans of the node-sharded mode (no counterpart in the single-process reference; SURVEY.md
 * section 8e).  plan: slot[i] = (ids[i] mod n_ranks) * cap + arrival order inside that bucket, -1 for dropped rows
 * (ids[i] < 0, i >= *n_valid when n_valid != NULL, or bucket full -> *overflow |= 1); counts[g] = rows destined to
 * rank g; local_id[i] = ids[i] / n_ranks (may be NULL).  scatter_rows / gather_words move rows of w 32-bit words into /
 * out of their slots (gather fills rows whose slot is < 0 with `fill`); n_valid (device, may be NULL) bounds the rows
 * walked when the table is sized for the worst case.  pack / unpack_queries: the request rows of
 * the neighbour exchange, [local node id | timestamp (2 words) | query id]. */
int pfo_route_plan(const int32_t* ids, int64_t n_rows, const int32_t* n_valid, int n_ranks, int cap,
                   int32_t* counts, int32_t* slot, int32_t* local_id, int32_t* overflow, void* stream);
int pfo_scatter_rows(const void* src, int64_t lds, const int32_t* slot, int64_t M, const int32_t* n_valid, int w,
                     void* dst, int64_t ldd, void* stream);
int pfo_gather_words(const void* src, int64_t lds, const int32_t* slot, int64_t M, int w, void* dst, int64_t ldd,
                     uint32_t fill, void* stream);
int pfo_pack_queries(const int32_t* local_id, const double* q_ts, const int32_t* q_ids, const int32_t* slot,
                     int64_t n_queries, int32_t* out_rows, void* stream);
int pfo_unpack_queries(const int32_t* in_rows, int64_t n_rows, int32_t* q_nodes, double* q_ts, int32_t* q_ids,
                       void* stream);
/* fused unpacking of the two request / reply exchanges: unroute_neighbors = reply rows [nbr | eidx | dt] (3n words) at
 * slot[q] -> three dense [Q, n] arrays; route_reply_rows (owner) = [updated memory row | last_update'] per received id;
 * unroute_rows (requester) = rows of the unique-node table from their slots plus H0 = memory' + node_feat[uniq]. */
int pfo_unroute_neighbors(const int32_t* back, const int32_t* slot, int64_t n_queries, int n, int32_t* nbr,
                          int32_t* eidx, float* dt, void* stream);
int pfo_route_reply_rows(const float* Hnew_own, const float* lu_own, const int32_t* slots_own, int64_t n_rows, int d,
                         float* reply, void* stream);
int pfo_unroute_rows(const float* back, const int32_t* slot, const int32_t* uniq, const int32_t* n_valid,
                     const float* node_feat, int64_t n_rows, int d, float* Hnew, float* lu_u, float* H0, void* stream);


/* ---- jodie time-projection embedding --- modules/embedding_module.py:57-61, model/tgn.py:260-266 */
int pfo_time_embedding_fwd(const int32_t* q_nodes, const double* q_ts, int64_t Q, int64_t n_src, int d,
                           const int32_t* slot_of_node, const float* Hnew, const float* lu_u,
                           float mean_src, float std_src, float mean_dst, float std_dst,
                           const float* W, const float* b, float* td_out, float* emb, void* stream);
int pfo_time_embedding_bwd(const int32_t* q_nodes, int64_t Q, int d, const int32_t* slot_of_node,
                           const float* Hnew, const float* td, const float* W, const float* b,
                           const float* dEmb, float* dHnew, float* dWdb, float* workspace,
                           int64_t workspace_floats, void* stream);

/* small row utilities (gather with idx < 0 -> zero row; scatter-add = its gradient) */
/* ---- optimiser step --- main.py:123,389 (torch.optim.Adam(lr), defaults otherwise) over ONE flat parameter / gradient /
 * moment buffer (the trainer lays the parameters out in the engine's operand order and the step's backward writes its
 * gradients straight into the flat gradient buffer).  *step = t >= 1, kept on the device. */
int pfo_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, const int32_t* step, void* stream);

/* ---- TimeEncode.forward alone --- model/time_encoding.py:17-25: out_cos[m, c] = cos(fmaf(t[m], w[c], b[c])) (fmaf ==
 * nn.Linear(1, d) bit for bit), out_sin optional.  mode 0 / 1 = the fp64 quadrant reduction the fused kernels use
 * (any |x| < 2^44), 2 = fp32 Cody-Waite reduction (|x| < 2^17; kept for the comparison test), 3 = the cosine-only
 * half-turn reduction + one even polynomial of the neighbour forward kernel (any |x| < 2^44; out_sin as in mode 0). */
int pfo_time_encode(const float* t, const float* w, const float* b, int64_t M, int d, int mode,
                    float* out_cos, float* out_sin, void* stream);
int pfo_reduce_partials(const float* partial, int rows, int cols, float* out, int accumulate, void* stream);
int pfo_scatter_add_rows(const float* src, int64_t lds, const int32_t* idx, int64_t M, int d,
                         float* dst, int64_t ldd, void* stream);
int pfo_gather_rows(const float* src, int64_t lds, const int32_t* idx, int64_t M, int d,
                    float* dst, int64_t ldd, void* stream);

/* ---- K4 neighbour level: fused gather + TimeEncode + masked softmax over sampled neighbours ---
 * model/temporal_attention.py:34-90, model/time_encoding.py:17-25, modules/embedding_module.py:141-154.
 * Weight-absorbed multi-head attention (see csrc/attention_kernels.cu).  QK/XB/dQK/dXB are
 * [Q, H, ekp] with segment [h (d) | e (F) | te (d) | psum | valid | one | 0..] (ekp >= 2d+F+3; `valid` = the
 * query has a neighbour, `one` = 1: constant columns that let the host fold biases into the consuming GEMM);
 * XB / dXB rows are ldxb / lddxb floats apart.  T is the feature table the
 * neighbour rows are gathered from (idx < 0 = padded neighbour); P [Q,H,n] = softmax weights.
 * Dropout on the attention weights (p_drop > 0) draws from the shared Philox stream keyed by
 * (query, slot, step + *step_dev) -- one Philox4x32 block per (query, slot), word h = head h; step_dev (optional)
 * is a device-resident batch counter.
 * pfo_attn_nbr_fwd_rows: the same kernel with the query operand shared between queries -- query q reads row
 * qk_row[q] of QK (NULL = row q).  qk_h depends on the query NODE only (the query's own time encoding is te(0),
 * model/temporal_attention.py:48-50), so when a batch asks about few distinct nodes at many times -- full-ranking
 * evaluation, evaluation.py:84-115: every user against all stocks -- the caller computes one row per node. */
int pfo_attn_nbr_fwd_rows(const float* QK, const int32_t* qk_row, const float* T, int64_t ldt, const int32_t* idx,
                          const int32_t* eidx, const float* dt, const float* efeat, const float* tw, const float* tb,
                          int64_t Q, int n, int d, int F, int H, int ekp, float p_drop, uint64_t seed, uint32_t step,
                          const uint32_t* step_dev, float* XB, int64_t ldxb, float* P, int32_t* invalid, void* stream);
int pfo_attn_nbr_fwd(const float* QK, const float* T, int64_t ldt, const int32_t* idx, const int32_t* eidx,
                     const float* dt, const float* efeat, const float* tw, const float* tb,
                     int64_t Q, int n, int d, int F, int H, int ekp, float p_drop, uint64_t seed, uint32_t step,
                     const uint32_t* step_dev, float* XB, int64_t ldxb, float* P, int32_t* invalid, void* stream);
int64_t pfo_attn_nbr_bwd_workspace_floats(int d);
int pfo_attn_nbr_bwd(const float* QK, const float* dXB, int64_t lddxb, const float* P, const int32_t* invalid,
                     const float* T, int64_t ldt, const int32_t* idx, const int32_t* eidx, const float* dt,
                     const float* efeat, const float* tw, const float* tb,
                     int64_t Q, int n, int d, int F, int H, int ekp, float p_drop, uint64_t seed, uint32_t step,
                     const uint32_t* step_dev, float* dQK, float* dT, int64_t lddt, float* dtw_dtb, int accumulate,
                     float* workspace, void* stream);

/* ---- parameter-side folding of the attention layer --- model/temporal_attention.py:26-90 (q/k/v/out projections of
 * nn.MultiheadAttention, TimeEncode(0) of the query) and utils/utils.py:4-17 (MergeLayer fc1).  Forward builds the
 * GEMM operands the per-query kernels consume: Wqk [H*ekp, d], cqk [H*ekp] (qk_h = Wqk_h h_q + cqk_h) and
 * Wc1T [H*ekp + d, d] (fc1([out_proj(attn) | h_q]) as one contraction over [XB | h_q], biases in the `valid` and
 * `one` rows).  Backward carries d/dWqk, d/dcqk, d/dWc1T back to the reference's tensors (all outputs fully
 * written; q_proj / k_proj / v_proj weights are [2d, 2d] / [2d, 2d+F] / [2d, 2d+F], in_proj_bias [6d], out_proj
 * [2d, 2d] + [2d], fc1 [d, 3d] + [d], tb = TimeEncode bias [d]).  fp64 arithmetic; `workspace` holds
 * pfo_fold_attention_workspace_doubles doubles and must be kept from the forward to the backward call. */
int64_t pfo_fold_attention_workspace_doubles(int d, int F, int H);
int pfo_fold_attention_fwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                           const float* bo, const float* W1, const float* b1, const float* tb,
                           int d, int F, int H, int ekp, double* workspace,
                           float* Wqk, float* cqk, float* Wc1T, void* stream);
int pfo_fold_attention_bwd(const float* Wq, const float* Wk, const float* Wv, const float* b_in, const float* Wo,
                           const float* bo, const float* W1, const float* b1, const float* tb,
                           int d, int F, int H, int ekp, double* workspace,
                           const float* gWqk, const float* gcqk, const float* gWc1T,
                           float* gWq, float* gWk, float* gWv, float* gb_in, float* gWo, float* gbo,
                           float* gW1, float* gb1, float* gtb, void* stream);

/* ---- K6: BPR loss forward + backward --- main.py:321-337 / :364-381.  du/dp/dn may be null
 * (forward only).  workspace: 1024 floats. */
int pfo_bpr(const float* eu, const float* ep, const float* en, int B, int k, int d,
            float* du, float* dp, float* dn, float* loss, float grad_scale, float* workspace, void* stream);

/* ---- evaluation scoring + ranking --- evaluation.py:107-115,134-138 (stable argsort, reversed).
 * scores [B, 1+n_cand]; pos_rank[b] = position of the true item; top_idx [B, topk] candidate
 * positions (0 = the true item). */
int pfo_eval_score(const float* es, const float* ed, const float* ec, int B, int n_cand, int d, int topk,
                   float* scores, int32_t* pos_rank, int32_t* top_idx, void* stream);

/* ---- evaluation metric block --- evaluation.py:127-207 (per-interaction loop: Recall/NDCG@{1,3,5}, :11-21, and
 * return_sharpe_at_k, :23-36, in and out of sample) and :209-258 (means / fraction-positive).  pos_rank / top_idx are
 * pfo_eval_score's outputs (topk >= 5); pos_item / cand are item ids, stock = id - item_offset; the portfolio CSR
 * holds 0-based stocks (an empty row = the reference's ['']); logret_* are float64 [n_days, n_stocks, n_returns]
 * log(p[1:]/p[:-1]) of time_feature_past / time_feature_future (n_returns <= 32).  per_event is float64 [B, 18] =
 * [recall@1,3,5 | ndcg@1,3,5 | d_return_in@1,3,5 | d_sharpe_in@1,3,5 | d_return_out@1,3,5 | d_sharpe_out@1,3,5],
 * rounded exactly like the reference's numpy fp64 arithmetic.  acc (optional, float64 [31]) accumulates over
 * calls: column sums [0..18), positive counts of columns 6..17 [18..30), number of interactions [30]. */
int pfo_eval_metrics(const int32_t* pos_rank, const int32_t* top_idx, int topk, const int32_t* pos_item,
                     const int32_t* cand, int n_cand, int item_offset, const int32_t* day_idx,
                     const int64_t* port_ptr, const int32_t* port_items, const double* logret_past,
                     const double* logret_future, int n_stocks, int n_returns, int B,
                     double* per_event, double* acc, void* stream);

/* ---- K5: candidate sampling + mean-variance efficient selection --- utils/utils.py:65-114
 * (RandEdgeSampler) and main.py:197-304 (inline block).  Stocks are 0-based indices; logret is
 * float64 [n_days, n_stocks, n_returns].  sample != 0 draws cand[:,1:] from the Philox stream
 * (cand[:,0] = pos_stock - item_offset); sample == 0 takes cand as input.  p_neg is [n_neg-th lowest .. lowest].
 * item_offset: pos_stock is read, and p_pos / p_neg are written, as stock index + item_offset (the item ids of the
 * interaction stream, upper_u + 1 + stock); cand stays 0-based. */
int pfo_mv_select(const int64_t* event_ids, const int32_t* day_idx, const int32_t* pos_stock,
                  const int64_t* port_ptr, const int32_t* port_items,
                  const int32_t* items_sorted, int n_items_universe,
                  const double* logret, int n_stocks, int n_returns,
                  int B, int K, double gamma, double lam, int n_pos, int n_neg, uint64_t seed, int sample,
                  int32_t* cand, double* y_out, int32_t* p_pos, int32_t* p_neg, int item_offset, void* stream);
int pfo_sample_candidates(const int64_t* event_ids, const int64_t* port_ptr, const int32_t* port_items,
                          const int32_t* items_sorted, int n_items_universe, int B, int size,
                          uint64_t seed, int32_t* out, void* stream);

/* ---- peer-memory transport of the node-sharded exchanges (no counterpart in the reference; csrc/peer_kernels.cu).
 * alloc: cudaMalloc'ed, zeroed arena + its 64-byte CUDA IPC handle; open / close: map / unmap a peer's arena; the first
 * pfo_peer_header_bytes() of every arena hold the barrier flags and epochs.  push: block g (block_words 32-bit words) of
 * the local send buffer -> arena of rank g at recv_off_bytes + rank * block_words * 4 (`bases`: HOST array of the
 * n_ranks arena addresses in this process).  barrier `id`: signal every peer, wait for every peer (error_word |= 2 after
 * timeout_s instead of hanging). */
int pfo_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int pfo_peer_open(const unsigned char* handle64, void** ptr);
int pfo_peer_close(void* ptr);
int pfo_peer_free(void* ptr);
int pfo_peer_max_ranks(void);
int pfo_peer_max_barriers(void);
int64_t pfo_peer_header_bytes(void);
int pfo_peer_push(const void* send, const uint64_t* bases, int64_t recv_off_bytes, int n_ranks, int rank,
                  int64_t block_words, void* stream);
int pfo_peer_barrier(const uint64_t* bases, int n_ranks, int rank, int id, uint32_t* error_word, double timeout_s,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFO_B200_H */
