/* b200rec.h — C ABI of libb200rec.so: hand-written sm_100a kernels for the HSTU multi-head
 * train / eval hot path.
 *
 * The reference (zhykoties/Multi-Head-Recommendation-with-Human-Priors) has NO native/FFI
 * interface: its hot path is eager PyTorch inside code/REC/model/IDNet/hstu.py and
 * code/REC/evaluator/collector.py.  Each entry point below therefore cites the reference
 * Python op sequence (file:line under code/REC/) it replaces.  The Python host
 * (b200rec.hstu.HSTU, same constructor / forward / predict / compute_item_all as the
 * reference class) binds these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; b200rec_last_error() gives
 *     the message of the last failure on the calling thread.
 *   - all pointers are DEVICE pointers unless the name ends in _host; no function allocates
 *     or synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - dtype codes: B200REC_F32 = 0, B200REC_BF16 = 1.  "act" tensors use the compute dtype of
 *     the mode (bf16 production / fp32 verification); residual stream, LN statistics, logits,
 *     losses and parameter gradients are always fp32.
 *   - jagged layout: only valid tokens are stored, T = number of tokens; tok_b[t] is the
 *     sequence of token t, tok_pos[t] its absolute position in the padded frame [0, L),
 *     seq_off[b]..seq_off[b+1] the token range of sequence b (tokens of one sequence are
 *     contiguous and ordered by position).
 */
#ifndef B200REC_H
#define B200REC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REC_F32 0
#define B200REC_BF16 1

const char* b200rec_last_error(void);
int b200rec_version(void);
/* 1 when the current device is sm_100 (tcgen05/TMA kernels usable). */
int b200rec_device_is_sm100(void);

/* ------------------------------------------------------------------ embedding (SURVEY §8 a1,a2)
 * hstu.py:637,670,752 `self.item_embedding(ids)`; :640-643 position embedding add. */
/* out[r,:] = table[ids[r],:]          (out dtype f32 or bf16) */
int b200rec_gather_rows(const float* table, int D, const int64_t* ids, int64_t n_ids,
                        void* out, int out_dtype, void* stream);
/* x[t,:] = table[items[tok_b[t]*LP + tok_pos[t]],:] + pos_table[tok_pos[t],:] */
int b200rec_embed_tokens(const float* table, const float* pos_table, const int64_t* items,
                         const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int D,
                         float* x, void* stream);
/* hstu.py:670-672 / :605-606: gather + L2 normalise.  out_hat[r] = row/||row|| (act dtype),
 * inv_norm[r] = 1/||row||.  table==NULL -> rows are read from `rows_in` (fp32 [n,D]) instead. */
int b200rec_gather_l2norm(const float* table, const float* rows_in, int D, const int64_t* ids,
                          int64_t n, void* out_hat, int out_dtype, float* inv_norm, void* stream);
/* backward of x_hat = x/||x||:  dx = (dxh - xh * <xh, dxh>) * inv_norm.   dxh fp32. */
int b200rec_l2norm_bwd(const void* x_hat, int act_dtype, const float* inv_norm, const float* d_xhat,
                       int64_t n, int D, float* dx, int accumulate, void* stream);

/* d_pos[pos,:] = sum over sequences (ascending b) of dx0[tok_index[b*LP+pos],:]; pos in [0, L). */
int b200rec_pos_emb_grad(const float* dx0, const int32_t* tok_index, int B, int LP, int L, int D,
                         float* d_pos, void* stream);
/* Decode-head ResBlock backward (llm_heads.py:26-40) on the [T, H, D] head block:
 * dz = d_hd * silu'(z) (act dtype; skipped when z == NULL), dy[t,:] = sum_h d_hd[t,h,:]. */
int b200rec_resblock_bwd(const float* d_hd, const void* z, int act_dtype, int64_t T, int H, int D,
                         void* dz, float* dy, void* stream);

/* Deterministic embedding gradient (replaces autograd's atomic embedding_dense_backward).
 * ids[n] (i64) name the table row of each of the n gradient rows grad_rows[n,D] (fp32).  Rows are
 * radix-sorted by id (stable), and each unique id's rows are summed in ascending input position by
 * one warp-group: bit-reproducible.  id 0 (padding_idx, hstu.py:413) and ids < 0 get no gradient.
 * Outputs: uniq_ids[<=n], uniq_rows[<=n, D], n_uniq (device int32), row_slot[N] is NOT touched.
 * workspace: b200rec_scatter_add_workspace_bytes(n). */
size_t b200rec_scatter_add_workspace_bytes(int64_t n_ids);
int b200rec_scatter_add_sorted(const int64_t* ids, int64_t n_ids, const float* grad_rows, int D,
                               int64_t* uniq_ids, float* uniq_rows, int32_t* n_uniq,
                               void* workspace, size_t workspace_bytes, void* stream);
/* Owner-side gradient reduction of a ROW-SHARDED table inside one NVLink node (replaces the reference's dense-gradient
 * reduce of trainer.py:434-453 / an all-to-all of gradient rows): ids[W * n_per] are the global ids behind the gradient
 * rows of all W ranks (rank-major), src_ptrs_dev[r] is rank r's gradient-row buffer [n_per, D] mapped with CUDA IPC
 * (b200rec_ipc_import).  Keeps the ids this rank owns (id % W == rank, id > 0), returns their LOCAL rows id / W in
 * uniq_ids and the per-id sums in ascending (rank, index) order: the rows are read over NVLink while they are summed. */
int b200rec_scatter_add_sorted_peer(const int64_t* ids, int64_t n_per, int W, int rank, const void* src_ptrs_dev, int D,
                                    int64_t* uniq_ids, float* uniq_rows, int32_t* n_uniq, void* workspace,
                                    size_t workspace_bytes, void* stream);
/* Row lookup in a row-sharded table (global id g lives on rank g % W at local row g / W): shard_ptrs_dev[r] = rank r's
 * shard mapped with CUDA IPC; out[i,:] = shard[ids[i] % W][ids[i] / W, :], zero row for ids[i] < 0.  Replaces the
 * embedding gather of hstu.py:637,670,752 when the table does not fit / is not replicated; no host sync, fixed size. */
int b200rec_gather_rows_sharded(const void* shard_ptrs_dev, int W, int D, const int64_t* ids, int64_t n_ids,
                                float* out, void* stream);
/* CUDA IPC plumbing for the two calls above: export = (64-byte cudaIpcMemHandle_t of the cudaMalloc block containing
 * ptr, byte offset of ptr inside it); import maps a peer's block and returns its base; close unmaps it. */
int b200rec_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out);
int b200rec_ipc_import(const void* handle64, void** base_out);
int b200rec_ipc_close(void* base);
/* dense[uniq_ids[i],:] (+)= uniq_rows[i,:] for i < *n_uniq (rows are unique -> deterministic). */
int b200rec_rows_to_dense(const int64_t* uniq_ids, const float* uniq_rows, const int32_t* n_uniq,
                          int64_t max_rows, int D, float* dense, int accumulate, void* stream);

/* ------------------------------------------------------------------ norms (a4)
 * hstu.py:213-219 F.layer_norm(x,[D],eps) without affine. */
int b200rec_layernorm_fwd(const float* x, int T, int D, float eps, void* y, int y_dtype,
                          float* mean, float* rstd, void* stream);
/* dx = LN'(x; dy) (+ residual_grad if not NULL).  dy in act dtype, leading dimension ldy.  If dx_act is
 * not NULL it receives a copy of dx in the dy dtype (the GEMM operand of the next block's backward). */
int b200rec_layernorm_bwd(const void* dy, int dy_dtype, int ldy, const float* x, const float* mean,
                          const float* rstd, int T, int D, const float* residual_grad, float* dx,
                          void* dx_act, void* stream);
/* hstu.py:277 o_input = u * LN(attn): oin = u * LN(a).  u has leading dimension ldu (it is a column
 * slice of the uvqk activation). */
/* dropout_p > 0 applies F.dropout to oin (hstu.py:281-285) with a stateless Philox4x32-10 keep-mask keyed by
 * (seed, layer, *rng_step_dev, element): statistically equivalent to torch's, not bit-matched; the backward
 * regenerates the same mask.  rng_step_dev is a device counter (b200rec_counter_add) so graph replays differ. */
int b200rec_gate_ln_fwd(const void* u, int ldu, const float* a, int T, int D, float eps, void* oin,
                        int act_dtype, float* mean, float* rstd, float dropout_p, uint32_t seed,
                        uint32_t layer, const int64_t* rng_step_dev, void* stream);
int b200rec_counter_add(int64_t* counter_dev, int64_t v, void* stream);
/* backward: du = d_oin*LN(a) ; da = LN'(a; d_oin*u) (act dtype).  Writes d_pre_u = du * silu'(pre_u) directly
 * (pre_u = pre-activation slice, ld ldu) into d_pre_u (ld ldu). */
int b200rec_gate_ln_bwd(const void* d_oin, const void* u, const void* pre_u, int ldu, const float* a,
                        const float* mean, const float* rstd, int T, int D, void* d_pre_u, void* da,
                        int act_dtype, float dropout_p, uint32_t seed, uint32_t layer,
                        const int64_t* rng_step_dev, void* stream);
/* y = cast(x) elementwise, n elements (fp32 -> act dtype). */
int b200rec_cast(const float* x, int64_t n, void* y, int y_dtype, void* stream);
/* col_sum[j] = sum_i x[i, j]  (deterministic: 256-row slabs, then the slabs in ascending order by the block that finishes
 * last for its 32-column tile; one launch), x act dtype or fp32, ld = ldx.  workspace: b200rec_colsum_workspace_bytes(rows,
 * cols) bytes, 256-byte aligned; its first 4096 bytes (one uint32 counter per 32-column tile, cols <= 32768) must be ZERO on
 * entry and are zero again when the kernel ends, so one zero-initialised buffer serves every call of a stream. */
size_t b200rec_colsum_workspace_bytes(int rows, int cols);
int b200rec_colsum(const void* x, int x_dtype, int ldx, int rows, int cols, float* out,
                   int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ GEMM (a4,a7,a10,a15)
 * C[M,N] = epilogue( A[M,K] * B[N,K]^T ).  Replaces torch.matmul / nn.Linear / einsum at
 * hstu.py:243 (uvqk), :280 (_o), llm_heads.py:40 (ResBlock), hstu.py:611-613 (NCE logits / fix),
 * :979 (eval scoring) and their autograd backward.
 * in_dtype BF16 -> tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), TMA-fed, persistent;
 * in_dtype F32  -> SIMT fp32 verification kernel (same epilogues).
 * a_major/b_major: 0 = K-major (A[m*lda+k], B[n*ldb+k]); 1 = MN-major (A[k*lda+m], B[k*ldb+n]). */
enum {
  B200REC_EPI_STORE = 0,      /* C = alpha*acc                                              */
  B200REC_EPI_ACCUM = 1,      /* C += alpha*acc (C fp32)                                    */
  B200REC_EPI_SILU_DUAL = 2,  /* C2 = acc (pre-activation), C = silu(acc)                   */
  B200REC_EPI_BIAS_RESID = 3, /* C = acc + bias[n] + resid[m, n]                            */
  B200REC_EPI_RESBLOCK = 4,   /* z = acc + bias[n]; C2 = z; C = resid[m, n % n_split] + silu(z) */
  B200REC_EPI_GT_BITS = 5     /* C is uint32 [M, N/32]: bit (n%32) of word n/32 = acc > alpha;
                                 optional C2 = u8[M], set to 1 for rows with any bit (zero it first) */
  ,
  B200REC_EPI_FOLD_HEADS = 6  /* eval scoring (hstu.py:979-999 + collector.py:241-275 fused): rows are
                                 (user, head) with `fold_hp` (power of two <= 32) rows per user; per item
                                 column the masked max over the user's heads goes to C = fval fp32
                                 [M/fold_hp, N] and the arg-max head to C2 = fhead u8 [M/fold_hp, N].
                                 Masks: fold_head_on[m] (u8, 0 = head off / padding row),
                                 fold_head_cat[h] (category gating head h, -1 none) against
                                 fold_item_tags[n] (bit c = item n has tag c), global item id 0
                                 (n*fold_id_stride + fold_id_offset == 0).  bf16 (tcgen05) path only. */
  ,
  B200REC_EPI_NCE_EXP = 7     /* fused sampled-softmax forward: C = bf16 softmax numerators + per-row partial
                                 sums (nce_* fields, described below b200rec_gemm).  bf16 (tcgen05) path only. */
  ,
  B200REC_EPI_FOLD_ITEMS = 8  /* FOLD_HEADS with the operands swapped: A = item rows [M = items, K], B = user heads
                                 [N = users * fold_hp, K] (fold_hp in {1, 2, 4, 8, 12, 16}: H = 12 needs no padding to
                                 16).  An accumulator row is one item, so the masked max / arg-max over a user's heads
                                 is taken in registers of one thread (no shuffles, no shared-memory transpose) and the
                                 epilogue hides under the MMAs.  Outputs as FOLD_HEADS: C = fval fp32 [users, ldc >= M],
                                 C2 = fhead u8 [users, ldc2 >= M], or the streamed candidate lists (fold_thr != NULL).
                                 Masks: fold_on_bits[user] (bit h = head h on), fold_head_cat[h], fold_item_tags[m],
                                 global item id m*fold_id_stride + fold_id_offset == 0.  Tile order: all user tiles of
                                 one item tile run back to back (the table streams from HBM once).  bf16 path only. */
};
typedef struct {
  int M, N, K;
  const void* A; int64_t lda; int a_major;
  const void* B; int64_t ldb; int b_major;
  int in_dtype;
  void* C; int64_t ldc; int c_dtype;
  void* C2; int64_t ldc2; int c2_dtype;
  int epilogue;
  float alpha;
  const float* alpha_dev; /* optional device scalar: STORE/ACCUM use alpha * (*alpha_dev) */
  const float* bias;
  const float* resid; int64_t ldr;
  /* n_split > 0: RESBLOCK reads resid[m, n % n_split] (one residual row shared by the H head blocks
   * of a [T, H*D] output).  If additionally c_split_stride > 0, column n is stored at
   * C + (n / n_split) * c_split_stride + m*ldc + n % n_split (per-head [H, T, D] blocks). */
  int n_split; int64_t c_split_stride; int64_t c2_split_stride;
  /* optional split-K workspace (fp32): when the output has few tiles and K is long (weight gradients) the
   * tcgen05 path splits each tile's k-range over idle SMs and sums the partials in fixed order.  Requires
   * epilogue STORE, fp32 C, ldc == N % 4 == 0 contiguous rows; NULL disables.  The same workspace serves the tail
   * split (b200rec_gemm_use_tail_split below), which has no layout requirement on C. */
  void* splitk_ws; size_t splitk_ws_bytes;
  /* B200REC_EPI_FOLD_HEADS only */
  int fold_hp;
  const uint8_t* fold_head_on; const int32_t* fold_head_cat; const uint32_t* fold_item_tags;
  int64_t fold_id_offset; int64_t fold_id_stride;
  /* FOLD_HEADS, streamed variant (fold_thr != NULL): nothing of size users x items is written.  A folded score that is
   * finite and >= fold_thr[user] is appended to the user's candidate list instead: slot = atomicAdd(fold_cnt[user]),
   * fold_keys[user * fold_cap + slot] = (~order_key(score) << 32) | (local item index << 5) | arg-max head, so that an
   * ascending sort of the keys IS the tie rule (value desc, item id asc, head asc); N < 2^27 rows per shard.  Slots >=
   * fold_cap are dropped; fold_cnt still counts them so the caller can detect the overflow.  C / C2 unused. */
  const float* fold_thr; uint32_t* fold_cnt; uint64_t* fold_keys; int fold_cap;
  /* streamed variant only: fold_groups = G > 1 lets a user own G groups of fold_hp heads (rows ordered (user, group,
   * head); head index = group * fold_hp + h <= 31), so H = 12 runs as 3 x 4 without padding to 16.  Every group appends
   * to the SAME user's list: an item may then appear once per group, b200rec_topk_from_candidates(dedupe = 1) keeps the
   * best entry (value desc, head asc). */
  int fold_groups;
  /* GT_BITS: optional rank-1 addend, bit = acc + gt_row[m] * gt_col[n] > alpha (fp32[M], fp32[N]; both or neither).
   * Used for the PRUNED false-negative filter: a K = 64 prefix product plus the product of the tail norms is an upper
   * bound of the full cosine (Cauchy-Schwarz), b200rec_gt_bits_verify then settles the few surviving pairs exactly. */
  const float* gt_row; const float* gt_col;
  /* STORE / ACCUM: optional per-row factor, C[m,:] (+)= alpha * row_scale[m] * acc[m,:] (fp32[M]; NULL = 1). */
  const float* row_scale;
  /* B200REC_EPI_NCE_EXP only (fused sampled-softmax forward, hstu.py:600-619 + cross_entropy): see below. */
  const float* nce_mref; const float* nce_thr; float* nce_stats; const float* nce_logit_scale;
  /* B200REC_EPI_FOLD_ITEMS only: uint32[users], bit h set = head h of that user takes part (fold_head_on is unused). */
  const uint32_t* fold_on_bits;
} b200rec_gemm_args;
int b200rec_gemm(const b200rec_gemm_args* args, void* stream);
/* B200REC_EPI_NCE_EXP (bf16 tcgen05 path): the logits GEMM of the sampled-softmax loss with the softmax numerators
 * produced in the epilogue, so that no [T, Nneg] fp32 logits tensor ever reaches HBM.  With acc = q_hat . n_hat (cosine),
 * tau2 = exp(clamp(*nce_logit_scale, 0, ln 100)) * log2(e) and a per-row reference nce_mref[m] (log2 domain):
 *     E[m, n]  = bf16( 2^(tau2 * acc - nce_mref[m]) )                      -> C (bf16 [M, ldc])
 *     nce_stats[m, part, 0..3] = { sum E, sum E * acc, #(acc > nce_thr[m]), 0 } over the columns of `part`,
 * part = 2 * (n / BN) + (which of the two epilogue warps), n_parts = b200rec_gemm_nce_parts(N).  Sums use the ROUNDED
 * E so that later subtractive corrections (false-negative filter) stay consistent with the stored tile.  The combine
 * step (b200rec_nce_combine) turns the partials into per-offset log-sum-exp / loss / gradient scalars; the backward
 * GEMMs consume E directly: dq = row_scale * (E @ n_hat), dn = E^T @ (row_scale * q_hat). */
/* Programmatic dependent launch of the tcgen05 GEMM kernels (default on; env B200REC_PDL=0 disables): the kernel's
 * prologue (barriers, TMEM allocation, tensor-map prefetch) overlaps the previous kernel's tail, every global access
 * happens after griddepcontrol.wait. */
void b200rec_gemm_use_pdl(int on);
/* Tail split of the tcgen05 GEMM (default on; env B200REC_TAIL_SPLIT=0 disables).  When the output has a few tiles
 * more than a whole number of waves (config B's O-proj / d_oin / dn: 84 tiles on 74 CTA pairs), the r tiles of the last
 * wave are cut into s = units / r k-ranges, one per otherwise idle CTA (pair); the fp32 partial accumulators go to
 * splitk_ws and a second small kernel sums them in ascending split order (deterministic) and applies the epilogue
 * (STORE / ACCUM / SILU_DUAL / BIAS_RESID / RESBLOCK).  Needs splitk_ws (16-byte aligned, r * s tiles of fp32). */
void b200rec_gemm_use_tail_split(int on);
int b200rec_gemm_nce_parts(int N);
/* Pruned false-negative filter (hstu.py:613-614: fix_logits = target @ neg^T > nce_thres), exact result, ~16x fewer
 * FLOPs than the full product:  tail_norm[r] = || x_hat[r, k0:] ||  (bf16 rows);  GT_BITS GEMM over the first k0
 * columns with gt_row / gt_col = the tail norms marks every pair whose cosine CAN exceed the threshold;
 * gt_bits_verify recomputes the full fp32 dot product of each marked pair, clears the bits that fail and sets
 * row_any[m] = 1 for rows that keep a bit (zero row_any first). */
int b200rec_tail_norm(const void* x_hat, int64_t n, int D, int k0, float* out, void* stream);
/* Same bound without an epilogue term: out bf16 [n, k0 + 16] = (x_hat[:, :k0], tail norm rounded UP to bf16, 15 zeros);
 * a plain GT_BITS GEMM of two such operands (K = k0 + 16) computes prefix product + |tail_a| |tail_b| on the tensor cores. */
int b200rec_prefix_aug(const void* x_hat, int64_t n, int D, int k0, void* out, void* stream);
int b200rec_gt_bits_verify(uint32_t* bits, int64_t M, int n_words, int N, const void* a_hat, const void* b_hat, int D,
                           float thres, uint8_t* row_any, void* stream);
/* The same over n_sets negative sets in one launch (grid.y = set): the sets share a_hat; set s uses bits + s * bits_stride
 * (uint32 elements), b_hat + s * b_stride (bf16 elements) and row_any + s * any_stride (bytes). */
int b200rec_gt_bits_verify_sets(uint32_t* bits, int64_t M, int n_words, int N, const void* a_hat, const void* b_hat,
                                int D, float thres, uint8_t* row_any, int n_sets, int64_t bits_stride,
                                int64_t b_stride, int64_t any_stride, void* stream);
/* n_groups independent problems args[0..n_groups).  Problems of identical shape / layout / dtypes with a
 * plain STORE or ACCUM epilogue (the per-head NCE GEMMs of hstu.py:697, the per-layer weight gradients)
 * run as ONE persistent tcgen05 launch (16 problems per launch), so that small problems share waves
 * instead of each paying its own tail; anything else is executed problem by problem.  Outputs must not
 * alias each other. */
int b200rec_gemm_grouped(const b200rec_gemm_args* args, int n_groups, void* stream);

/* ------------------------------------------------------------------ HSTU attention (a3,a5)
 * hstu.py:137-160: per head, A = silu(q k^T) / n_pad * [key valid & j <= i], out = A v.
 * q,k,v are column slices of the [T, 4D] activation (leading dimension ld, act dtype), heads are
 * contiguous dh-wide column groups.  Jagged over seq_off[B+1]; key_valid[t] (u8) masks keys
 * (hstu.py:1025).  out fp32 [T, D]. */
int b200rec_hstu_attn_fwd(const void* q, const void* k, const void* v, int ld, int act_dtype,
                          const int32_t* seq_off, const uint8_t* key_valid, int B, int T,
                          int n_heads, int dh, float inv_n, int max_len, float* out, void* stream);
/* backward: recomputes S.  d_out act dtype [T,D].  Writes d_pre_{q,k,v} =
 * d{q,k,v} * silu'(pre_{q,k,v}) into the [T,4D] pre-activation-gradient buffer (ld, act dtype). */
int b200rec_hstu_attn_bwd(const void* q, const void* k, const void* v, const void* pre_q,
                          const void* pre_k, const void* pre_v, int ld, int act_dtype,
                          const int32_t* seq_off, const uint8_t* key_valid, int B, int T,
                          int n_heads, int dh, float inv_n, int max_len, const void* d_out,
                          void* d_pre_q, void* d_pre_k, void* d_pre_v, void* stream);

/* Tensor-core variants (bf16 only, dh in {32, 64}): tcgen05.mma for QK^T / AV / the backward
 * contractions, TMA-fed, scores kept in TMEM / shared memory.  act, pre and d_pre point at column 0 of
 * the [T, 4D] u|v|q|k buffers (ld = 4D); d_out is bf16 [T, D]; out is fp32 [T, D]. */
int b200rec_hstu_attn_tc_fwd(const void* act, int ld, const int32_t* seq_off, const uint8_t* key_valid,
                             int B, int T, int n_heads, int dh, float inv_n, float* out, void* stream);
int b200rec_hstu_attn_tc_bwd(const void* act, const void* pre, int ld, const int32_t* seq_off,
                             const uint8_t* key_valid, int B, int T, int n_heads, int dh, float inv_n,
                             const void* d_out, void* d_pre, void* stream);

/* Short-sequence variants (bf16, dh in {32, 64}, every real sequence <= max_len <= 64 tokens): one CTA
 * per (sequence, head), warp-level mma.sync on register-resident score blocks, backward (dq, dk, dv)
 * in ONE launch without atomics.  Same buffers as the _tc_ variants.  A sequence longer than 64 tokens
 * is only legal if none of its keys is valid (the all-padding dummy row of static-shape mode); its
 * outputs / gradients are zeros. */
int b200rec_hstu_attn_seq_fwd(const void* act, int ld, const int32_t* seq_off, const uint8_t* key_valid,
                              int B, int T, int n_heads, int dh, float inv_n, int max_len, float* out,
                              void* stream);
int b200rec_hstu_attn_seq_bwd(const void* act, const void* pre, int ld, const int32_t* seq_off,
                              const uint8_t* key_valid, int B, int T, int n_heads, int dh, float inv_n,
                              int max_len, const void* d_out, void* d_pre, void* stream);

/* ------------------------------------------------------------------ NCE / sampled softmax (a9-a12)
 * hstu.py:600-619 nce_loss + :697 cross_entropy + :704-713 per-offset means, restructured:
 * one query row per (head, token) shared by all offsets p (SURVEY A.4 dedup identity).
 *
 * For one (negative set, head): logits[T, n_neg] (fp32) = q_hat @ neg_hat^T already computed by
 * b200rec_gemm; same_bits[B*LP, n_neg/32] = bits of (t_hat @ neg_hat^T > nce_thres).
 * Token (t,p) is valid iff tok_ok[(tok_b[t]*LP + tok_pos[t]+1+p)*tok_ok_ld + tok_ok_col] != 0
 * (caller folds attn-mask & category tag into tok_ok; p_mask selects offsets served by this head).
 * Outputs per (t,p): loss[t*P+p] (0 if invalid), g0[t*P+p] = coef*(softmax_0 - 1),
 * dscale[t*P+p] = coef * sum_k softmax-grad_k * z_k, rank0[t*P+p] = #negatives with logit > pos
 * (or -1), nvalid[t*P+p] = #unmasked logits incl. pos.  If G != NULL also writes
 * G[t, j] = tau * sum_p coef_p * softmax_p[j] (act dtype), the gradient w.r.t. the cosine logits.
 * coef[p] (device fp32[P]) = lambda_p * w_c / max(cnt_p, 1)  (hstu.py:708-712, 850-852).
 * row_any[B*LP] (u8, from the GT_BITS epilogue's C2: row has any filtered negative) and pos_ws
 * (fp32 [T*P] scratch) enable the register-resident fast path (n_neg <= 8192, n_neg % 4 == 0);
 * pass NULL for either to use the generic shared-memory kernel. */
int b200rec_nce_loss_fwd(const float* logits, int64_t ld_logits, int n_neg,
                         const uint32_t* same_bits, const uint8_t* row_any, float* pos_ws,
                         const void* q_hat, int64_t ldq,
                         const void* t_hat, int act_dtype, int D, const int32_t* tok_b,
                         const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask,
                         const uint8_t* tok_ok, int tok_ok_ld, int tok_ok_col, const float* coef,
                         const float* logit_scale, float* loss, float* g0, float* dscale,
                         int32_t* rank0, int32_t* nvalid, void* G, int64_t ldg, void* stream);
/* REMI's interest-aware hard negatives (REC/model/IDNet/remi.py:198-277 with beta_ihn > 0, + autograd) on the same
 * inputs and with the same outputs as b200rec_nce_loss_fwd (generic shared-memory kernel only):
 *   loss = coef * (logaddexp(z_pos, LSE_j((beta+1) z_j) - LSE_j(beta z_j) + log n_neg) - z_pos),  z = tau * cos,
 * the sums over the negatives the false-negative filter keeps, n_neg = all negatives of the row (remi.py:244).
 * g0 = d loss / d z_pos, dscale = d loss / d log tau, G[t, j] = d loss / d cos-logit (act dtype), rank0 / nvalid as above
 * (rank0 is filled for every served offset).  A row without any kept negative has loss 0 and zero gradients. */
int b200rec_nce_ihn_loss_fwd(const float* logits, int64_t ld_logits, int n_neg, const uint32_t* same_bits,
                             const void* q_hat, int64_t ldq, const void* t_hat, int act_dtype, int D,
                             const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask,
                             const uint8_t* tok_ok, int tok_ok_ld, int tok_ok_col, const float* coef,
                             const float* logit_scale, float beta, float* loss, float* g0, float* dscale,
                             int32_t* rank0, int32_t* nvalid, void* G, int64_t ldg, void* stream);
/* gscale (nullable device scalar) multiplies the upstream gradient in the two pos_bwd calls. */
/* coef[p] = lam[p] * w / max(cnt[p], 1)   (hstu.py:708-712, 850-852) */
int b200rec_nce_coef(const int32_t* cnt, const float* lam, float w, int P, float* coef, void* stream);
/* Relative position bias variant (north_star (b); reference hstu.py:74-134 builds the bias module but never applies it,
 * so this path is OFF for parity and only runs when config['apply_relative_attention_bias'] is set):
 *   A = silu(q k^T + bias_d[i - j]) / n_pad * mask,   bias_d fp32[max_len] indexed by the query-key distance.
 * bwd additionally fills dbias_part [b200rec_hstu_attn_bias_ws_floats(B, n_heads, max_len) / max_len rows, max_len]
 * (caller zero-fills); d bias_d = its column sums (fixed order: deterministic).  fp32 / bf16 SIMT kernels. */
int b200rec_hstu_attn_bias_fwd(const void* q, const void* k, const void* v, int ld, int act_dtype,
                               const int32_t* seq_off, const uint8_t* key_valid, int B, int T, int n_heads,
                               int dh, float inv_n, int max_len, const float* bias_d, float* out, void* stream);
size_t b200rec_hstu_attn_bias_ws_floats(int B, int n_heads, int max_len);
int b200rec_hstu_attn_bias_bwd(const void* q, const void* k, const void* v, const void* pre_q,
                               const void* pre_k, const void* pre_v, int ld, int act_dtype,
                               const int32_t* seq_off, const uint8_t* key_valid, int B, int T, int n_heads,
                               int dh, float inv_n, int max_len, const float* bias_d, const void* d_out,
                               void* d_pre_q, void* d_pre_k, void* d_pre_v, float* dbias_part, void* stream);

/* ------------------------------------------------------------------ prior-switch aux heads (a13)
 * hstu.py:512-544, 731-805 + layers.py:16-84: Linear(D -> 1) per prior category on the body output of EVERY context
 * position, weighted BCE (mode 0, pos_weight[a]) or asymmetric loss (mode 1).  The reference body is dense, so padded
 * positions count too: a left-padded query row sees no valid key, its attention output is 0 and each block only adds
 * its output bias, i.e. y_pad = E[item] + P[pos] + sum_l b_o^l (bo_sum).
 *   switch_rows:  out[b * Ls + j, :] = tok_index[b, l0 + j] >= 0 ? y[token] : table[items_idx[b, l0 + j]] + pos_emb[l0 + j] + bo_sum
 *   switch_loss:  logits / element losses / (logit >= 0) == target flags / dlogit = d total / d logit (grad_norm folds
 *                 prior_switch_loss_weight and the mean); target[b, l, a] = any_p tags[b, l + 1 + p, head_cat[a]]
 *   switch_bwd:   dW, db (ascending row order), d rows routed to dy[token] (valid) or pad_rows (padded; zero elsewhere)
 *   switch_pos_grad: dpos[l0 + j, :] += sum_b pad_rows[b * Ls + j, :] */
int b200rec_switch_rows(const float* y, const int32_t* tok_index, const int64_t* items_idx, const float* table,
                        const float* pos_emb, const float* bo_sum, int B, int LP, int l0, int Ls, int D,
                        float* out, void* stream);
int b200rec_switch_loss(const float* rows, int64_t R, int Ls, int l0, int D, const float* W_aux,
                        const float* b_aux, int n_act, const int32_t* head_cat, const float* pos_w,
                        const int64_t* tags, int LP, int C_tag, int P, int mode, float gamma_pos,
                        float gamma_neg, float clip, float eps, float grad_norm, float* logits, float* loss_el,
                        float* correct, float* dlogit, void* stream);
int b200rec_switch_bwd(const float* dlogit, const float* rows, const float* W_aux, int64_t R, int Ls, int l0,
                       int LP, int n_act, int D, const int32_t* tok_index, const float* gscale,
                       float* dW, float* db, float* dy, float* pad_rows, void* stream);
int b200rec_switch_pos_grad(const float* pad_rows, int B, int Ls, int l0, int D, float* dpos, void* stream);

/* Fused sampled-softmax path (bf16 production mode; replaces hstu.py:600-629 + cross_entropy without any fp32
 * [T, Nneg] tensor in HBM).  Order: b200rec_nce_pos_ref -> logits GEMM with B200REC_EPI_NCE_EXP -> b200rec_nce_combine.
 *   nce_pos_ref: pos_cos[T,P] (NaN = offset not served by this head / invalid token), mref[T] = reference exponent of the
 *                row (max valid positive logit in the log2 domain, floored at tau2 - 100; tau2 for unused rows),
 *                thr[T] = cosine of the offset-0 positive (+inf if absent): rank threshold of the top-k logging.
 *   nce_combine: from the GEMM's partial sums: loss / g0 (= d loss / d positive logit / tau) / dscale / rank0 / nvalid
 *                per (token, offset) exactly as b200rec_nce_loss_fwd returns them; row_scale[T] with
 *                d loss / d cos-logit[t, j] = row_scale[t] * E[t, j]; offsets whose target filters negatives
 *                (same_bits / row_any from the GT_BITS GEMM) subtract those stored numerators and E is patched at the
 *                filtered entries; qs[T, D] (bf16, nullable) = row_scale[t] * q_hat[t, :], the B operand of dn. */
int b200rec_nce_pos_ref(const void* q_hat, int64_t ldq, const void* t_hat, int D, const int32_t* tok_b,
                        const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask, const uint8_t* tok_ok,
                        int tok_ok_ld, int tok_ok_col, const float* logit_scale, float* pos_cos, float* mref,
                        float* thr, void* stream);
int b200rec_nce_combine(const float* stats, int n_parts, void* E, int64_t lde, int n_neg,
                        const uint32_t* same_bits, const uint8_t* row_any, const float* pos_cos,
                        const float* mref, const void* q_hat, int64_t ldq, int D, const int32_t* tok_b,
                        const int32_t* tok_pos, int T, int LP, int P, const float* coef,
                        const float* logit_scale, float* loss, float* g0, float* dscale, int32_t* rank0,
                        int32_t* nvalid, float* row_scale, void* qs, int64_t ldqs, void* stream);
/* Grouped forms: the loss of config B runs 12 (negative set, head) jobs over the same tokens; one launch per job left
 * every one of these row kernels launch / tail bound.  grid.y = job, at most 16 jobs per launch (longer lists are cut into
 * consecutive launches).  nce_combine_grouped: `ld_loss` is the row pitch of the loss outputs, so the jobs' [T, P] losses
 * can be column blocks of ONE [T, n_jobs * P] tensor and a single b200rec_colsum reduces them all; g0 / dscale / rank0 /
 * nvalid keep pitch P. */
typedef struct {
  const void* q_hat; uint32_t p_mask; int tok_ok_col; float* pos_cos; float* mref; float* thr;
} b200rec_nce_pos_ref_job;
int b200rec_nce_pos_ref_grouped(const b200rec_nce_pos_ref_job* jobs, int n_jobs, int64_t ldq, const void* t_hat, int D,
                                const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                                const uint8_t* tok_ok, int tok_ok_ld, const float* logit_scale, void* stream);
typedef struct {
  const float* stats; void* E; const uint32_t* same_bits; const uint8_t* row_any; const float* pos_cos;
  const float* mref; const void* q_hat; const float* coef; float* loss; float* g0; float* dscale; int32_t* rank0;
  int32_t* nvalid; float* row_scale; void* qs;
} b200rec_nce_combine_job;
int b200rec_nce_combine_grouped(const b200rec_nce_combine_job* jobs, int n_jobs, int n_parts, int64_t lde, int n_neg,
                                int64_t ldq, int D, const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                                int64_t ld_loss, const float* logit_scale, int64_t ldqs, void* stream);
/* positive-logit backward over several jobs: bwd_q -- the jobs of ONE call must have distinct d_qhat slices (grid.y =
 * job); bwd_t -- every job accumulates into the same d_that rows, in job order inside the kernel (deterministic). */
typedef struct { const float* g0; const void* q_hat; float* d_qhat; } b200rec_nce_pos_bwd_job;
int b200rec_nce_pos_bwd_q_grouped(const b200rec_nce_pos_bwd_job* jobs, int n_jobs, const void* t_hat, int act_dtype,
                                  int D, const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                                  const float* logit_scale, const float* gscale, int64_t ldd, void* stream);
int b200rec_nce_pos_bwd_t_grouped(const b200rec_nce_pos_bwd_job* jobs, int n_jobs, int64_t ldq, int act_dtype, int D,
                                  const int32_t* tok_index, int B, int LP, int P, const float* logit_scale,
                                  const float* gscale, float* d_that, void* stream);
/* cnt[p] = #valid tokens at offset p for this (head / category):  (hstu.py:705-707) */
int b200rec_nce_count(const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                      const uint8_t* tok_ok, int tok_ok_ld, int tok_ok_col, int32_t* cnt,
                      void* stream);
/* Every (validity column, loss weight) key of a step at once: cnt[c * P + p] (int32 workspace, n_col * P) = #valid tokens of
 * tok_ok column c at offset p, coef[k * P + p] = lam[p] * key_w[k] / max(cnt[key_col[k], p], 1).  key_col / key_w are HOST
 * arrays (n_keys <= 32); tok_ok is [rows, n_col] (n_col <= 32). */
int b200rec_nce_coefs(const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, const uint8_t* tok_ok, int n_col,
                      const int32_t* key_col, const float* key_w, int n_keys, const float* lam, int32_t* cnt,
                      float* coef, void* stream);
/* Top-k logging scalars of one job (hstu.py:621-629) from rank0 / nvalid [T, P]: over the rows whose offset 0 is served,
 * out[0..6] = { #rows, mean #logits per row, acc@1, acc@5, acc@10, acc@50, acc@100 }; acc_ws: 64 bytes of workspace. */
int b200rec_nce_topk_logs(const int32_t* rank0, const int32_t* nvalid, int T, int P, void* acc_ws, float* out,
                          void* stream);
/* positive-logit backward, deterministic:
 *   d_qhat[t,:]  += tau * sum_p g0[t,p] * t_hat[b, pos+1+p,:]            (query side)
 *   d_that[r,:]  += tau * sum_p g0[t(b, pos_r-1-p), p] * q_hat[t(...),:] (target side, gather form)
 * tok_index[b*LP + pos] = t or -1. */
int b200rec_nce_pos_bwd_q(const float* g0, const void* t_hat, int act_dtype, int D,
                          const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                          const float* logit_scale, const float* gscale, float* d_qhat, int64_t ldd,
                          void* stream);
int b200rec_nce_pos_bwd_t(const float* g0, const void* q_hat, int64_t ldq, int act_dtype, int D,
                          const int32_t* tok_index, int B, int LP, int P, const float* logit_scale,
                          const float* gscale, float* d_that, void* stream);
/* out[0] (+)= scale * sum_i x[i]   deterministic fixed-order tree. */
int b200rec_reduce_sum(const float* x, int64_t n, float scale, float* out, int accumulate,
                       void* stream);

/* ------------------------------------------------------------------ eval (a15-a17)
 * hstu.py:979-999 scores + prior masks, trainer.py:724-726 id-0 / history masks and
 * collector.py:241-275 cross-head merge, fused via the identity merge == topk(max over heads).
 * scores[B*H, N] fp32 (row = b*H + h) from b200rec_gemm.  head_cat[h] = category whose item-tag
 * bit masks head h (-1: none); item_tag_bits[N] u32 (bit c = item has tag c); head_on[B*H] u8
 * (prior_given_at_test / switch masks); hist_off[B+1], hist_items: per-user history ids.
 * Tie rule: value desc, item id asc, head asc.
 * Row i of `scores` is global item id i*id_stride + id_offset (1, 0 for an unsharded table; W, rank for
 * a table sharded by id % W): id 0 / history masks and the returned ids are in global ids.
 * Outputs topk_idx[B,K] i64, topk_val[B,K] f32, topk_head[B,K] i32.
 * workspace: b200rec_topk_workspace_bytes(B, N). */
size_t b200rec_topk_workspace_bytes(int B, int64_t N);
int b200rec_score_mask_topk(const float* scores, int64_t ld_scores, int B, int H, int64_t N, int K,
                            const int32_t* head_cat, const uint32_t* item_tag_bits,
                            const uint8_t* head_on, const int32_t* hist_off,
                            const int64_t* hist_items, int split_mode, int64_t id_offset,
                            int64_t id_stride, int64_t* topk_idx, float* topk_val, int32_t* topk_head,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Last stage of the STREAMED eval (collector.py:191-275 without any users x items tensor): per user, the candidates the
 * FOLD_HEADS epilogue appended (keys / cnt, capacity cap) minus history items and global id 0, ordered by
 * (value desc, item id asc); writes the first K.  overflow[0] is set to 1 if any user's count exceeded cap or fewer
 * than K candidates survived (the caller then re-runs that batch through the materialising path). */
int b200rec_topk_from_candidates(const uint64_t* keys, const uint32_t* cnt, int cap, int B, int K, int dedupe,
                                 const int32_t* hist_off, const int64_t* hist_items, int64_t id_offset,
                                 int64_t id_stride, int64_t* topk_idx, float* topk_val, int32_t* topk_head,
                                 int32_t* overflow, void* stream);
/* Second half of b200rec_score_mask_topk for scores already folded over heads (B200REC_EPI_FOLD_HEADS):
 * history suppression on fval, then per-user radix select + sort.  Same tie rule and id mapping.
 * fval / fhead rows have leading dimension ld >= N (a multiple of 4 keeps the 16-byte loads). */
int b200rec_topk_select(float* fval, const uint8_t* fhead, int B, int64_t N, int64_t ld, int K,
                        const int32_t* hist_off, const int64_t* hist_items, int64_t id_offset,
                        int64_t id_stride, int64_t* topk_idx, float* topk_val, int32_t* topk_head,
                        void* stream);
/* In-place masks for the reference-compatible predict() that returns [B,H,N] scores
 * (hstu.py:983-999): rows of switched-off heads and items outside the head's category -> -inf. */
int b200rec_apply_score_masks(float* scores, int64_t ld_scores, int B, int H, int64_t N,
                              const int32_t* head_cat, const uint32_t* item_tag_bits,
                              const uint8_t* head_on, void* stream);
/* collector.py:300-316: hit[b, k] |= topk_idx[b,k] in positive_i[b, 0:p+1]; pos_len quirk.
 * out[n_p, B, K+1] int32 for the n_p entries of pred_list (host array). */
int b200rec_hit_matrix(const int64_t* topk_idx, const int64_t* positive_i, int B, int K, int Pe,
                       const int32_t* pred_list_host, int n_p, int32_t* out, void* stream);

/* ------------------------------------------------------------------ batch construction (§8 f N2)
 * One kernel per batch instead of the reference's per-sample Python in DataLoader workers
 * (data/dataset/trainset.py:70-177, evalset.py:81-155, collate_fn.py:59-90).
 * Interaction data in CSR form: user_seq[user_off[u] .. user_off[u+1]) = item ids of user u (event_seq: the
 * parallel event types, nullable), train_len[u] = length of the training prefix, samples = (uid, context_end)
 * pairs (dataload.valid_sample_locations), batch_index[B] picks the samples of this batch.
 * Train row: [pad x (L-ctx) | seq[start:end+pred] | pad x (P-pred)]; pads are distinct random items outside the
 * row (pad_random) or 0; mask 1 on real positions; neg_items[b, s, :] = n_neg DISTINCT items drawn uniformly
 * from pool s (category pools cat_items[cat_off[s]..cat_off[s+1]) for s < n_pools, then the global pool
 * [1, item_num)), none of them in the padded row; with probability neg_sample_mix_ratio a category set draws
 * from the global pool (trainset.py:72-78).  tags[b, pos, c] = item_tags[item, c] (u8 [N, C], pads included,
 * trainset.py:165-167) or, when item_tags is NULL and event_seq is not, the one-hot event type at real
 * positions (trainset.py:147-153); tags may be NULL.  Randomness: Philox4x32-10 keyed on (seed, step, row, slot). */
int b200rec_build_train_batch(const int64_t* user_seq, const int64_t* user_off, const int32_t* train_len,
                              const int32_t* event_seq, const int64_t* sample_uid, const int32_t* sample_end,
                              const int64_t* batch_index, int B, int L, int P, int pad_random, int64_t item_num,
                              int n_sets, int n_neg, const int64_t* cat_items, const int64_t* cat_off, int n_pools,
                              float neg_sample_mix_ratio, const uint8_t* item_tags, int C, uint64_t seed,
                              uint64_t step, int64_t* items, int64_t* neg_items, int64_t* mask, int64_t* tags,
                              void* stream);
/* Eval rows: phase 0 (valid) history = seq[:train_len], targets = the next Pe items; phase 1 (test) history =
 * seq[:-Pe], targets = the last Pe.  item_seq[b] = last L history items left-padded with 0; (hist_u, hist_i) =
 * (row, item) for EVERY history item, written at hist_off[b] (host-side prefix sum of the history lengths);
 * target_tags like `tags` above (nullable). */
int b200rec_build_eval_batch(const int64_t* user_seq, const int64_t* user_off, const int32_t* train_len,
                             const int32_t* event_seq, const int64_t* uids, int B, int L, int Pe, int phase,
                             const uint8_t* item_tags, int C, const int64_t* hist_off, int64_t* item_seq,
                             int64_t* item_target, int64_t* target_tags, int64_t* hist_u, int64_t* hist_i,
                             void* stream);

/* ------------------------------------------------------------------ optimizer (§8 f N1)
 * torch.optim.AdamW semantics (trainer.py:296-299), fused single pass, fp32 state. */
/* coef_dev (nullable): device fp32[4] = {lr, 1-beta1^t, sqrt(1-beta2^t), t} overriding lr / step, advanced
 * by b200rec_adamw_tick — lets a captured CUDA graph contain the optimizer step. */
int b200rec_adamw(float* p, float* m, float* v, const float* g, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, float grad_scale,
                  const float* coef_dev, void* stream);
int b200rec_adamw_tick(float* coef_dev, float beta1, float beta2, void* stream);
/* Lazy dense-equivalent AdamW on the item table (exact torch.optim.AdamW values, deferred): a row with zero
 * gradient at step j needs only its own (p, m, v) and the step's scalars, so rows nobody reads are not touched.
 * last[row] (int32, zero-initialised) = last step applied to the row; hist (float4[cap]) = {1 - lr*wd,
 * lr / (1-beta1^j), 1 / sqrt(1-beta2^j), lr} of every step j, written by b200rec_adamw_tick_hist (which replaces b200rec_adamw_tick).
 *   b200rec_adamw_rows_catchup  brings the rows `ids` (duplicates allowed; NULL = all n_ids = n_rows rows) up to
 *                               the current step BEFORE something reads them (lookups, eval, checkpoint);
 *   b200rec_adamw_rows_lazy     applies the current step (tick already run) to the rows with a gradient.
 * id_stride > 1 (catchup only): `ids` are GLOBAL ids of a row-sharded table; ids with id % id_stride != id_offset
 * are skipped and the others address local row id / id_stride (every rank can pass the same all-gathered list).
 * hist is a RING of hist_cap (a power of two) entries indexed by step & (hist_cap - 1): the caller must bring every
 * row up to date (catchup with ids = NULL) at least once per hist_cap - 1 steps, so no row ever looks back further. */
int b200rec_adamw_tick_hist(float* coef_dev, void* hist, int cap, float beta1, float beta2,
                            float weight_decay, void* stream);
int b200rec_adamw_rows_catchup(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* ids,
                               int64_t n_ids, int32_t* last, const void* hist, int hist_cap, const float* coef_dev,
                               float beta1, float beta2, float eps, float weight_decay, int64_t id_stride,
                               int64_t id_offset, void* stream);
int b200rec_adamw_rows_lazy(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* uniq_ids,
                            const float* uniq_rows, const int32_t* n_uniq, int64_t max_rows, int32_t* last,
                            const void* hist, int hist_cap, const float* coef_dev, float beta1, float beta2, float eps,
                            float weight_decay, float grad_scale, void* stream);
/* One launch for many dense tensors: table_dev = array of {float* p, m, v; const float* g; bf16* shadow;
 * int64 n}, blocks_dev = int64 pairs {tensor index, first element of a 4096-element chunk}.  `shadow`
 * (nullable) receives the updated parameter rounded to bf16 — the GEMM operand of the next step, so the
 * weights are never re-cast. */
int b200rec_adamw_multi(const void* table_dev, const int64_t* blocks_dev, int n_blocks, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step,
                        float grad_scale, const float* coef_dev, void* stream);
/* dense-equivalent AdamW over an embedding table whose gradient is given in compact form
 * (uniq_ids, uniq_rows, n_uniq): rows without gradient use g = 0 (momentum still moves them). */
int b200rec_adamw_rows(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* uniq_ids,
                       const float* uniq_rows, const int32_t* n_uniq, int32_t* row_slot_ws, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step,
                       float grad_scale, const float* coef_dev, void* stream);

/* ------------------------------------------------------------------ prior construction
 * Co-occurrence graph of the reference's clustering scripts (code/item-clustering.py:152-162: all pairs of the distinct
 * items of a user's training window; code/user-clustering.py:268-290: all pairs of the distinct users of an item,
 * capped): `members` holds every group's members sorted ascending and de-duplicated, group g = members[group_off[g] ..
 * group_off[g+1]) (at most `cap` of them are used when cap > 0, the reference's MAX_USERS_PER_ITEM slice).
 * count: counts[g] = n (n - 1) / 2.  emit: keys[pair_off[g] + p] = (a << 32) | b over all pairs a < b of the group
 * (pair_off = exclusive prefix sum of counts); a 64-bit sort + unique of the keys is the reference's edge set. */
int b200rec_group_pairs_count(const int64_t* group_off, int64_t G, int cap, int64_t* counts, void* stream);
int b200rec_group_pairs_emit(const int32_t* members, const int64_t* group_off, const int64_t* pair_off, int64_t G,
                             int cap, uint64_t* keys, void* stream);
/* Modularity local moving (the phase shared by Louvain and by igraph's Leiden, which the reference calls at
 * code/item-clustering.py:240-245 with objective "modularity" and a resolution): node i's neighbouring communities are
 * given as runs (run_comm ascending, run_w = summed integer edge weight; no self loops) in [node_off[i], node_off[i+1]).
 * best_move: best[i] = argmax_C  w(i->C) - gamma * deg[i] * (tot[C] - [C == comm[i]] deg[i]) / two_m  if it beats
 * staying strictly (ties: smaller id), only for nodes with (i & 1) == parity, never from a singleton to a singleton of
 * larger id; IEEE double, no contraction.  apply: moves every node to best[i], updates tot / csize with integer atomics
 * (exact in any order) and counts the moves in *moved. */
int b200rec_louvain_best_move(const int64_t* node_off, const int32_t* run_comm, const int64_t* run_w,
                              const int32_t* comm, const int64_t* deg, const int64_t* tot, const int32_t* csize,
                              int64_t n, int64_t two_m, double gamma, int parity, int32_t* best, void* stream);
int b200rec_louvain_apply(const int32_t* best, int32_t* comm, const int64_t* deg, int64_t* tot, int32_t* csize,
                          int64_t n, uint64_t* moved, void* stream);

/* ------------------------------------------------------------------ ComiRec-SA readout on the HSTU body
 * Replaces REC/model/IDNet/comirec.py:232-300 (+ autograd): self-attentive multi-interest pooling over the causal
 * context and the hard readout against the future targets.  fp32, jagged tokens (valid positions only).
 *   pool_fwd:   u[t, k, :] = sum_{t' <= t in t's sequence} softmax_{t'}(a[t', k]) y[t', :]  in one online-softmax pass per
 *               (sequence, interest); M / S [T, K] = running maximum / denominator, kept for the backward.
 *   select_fwd: for token t and offset p, best = first argmax_k <u[t, k], traw[tok_b * LP + tok_pos + 1 + p]>,
 *               hd[t, p, :] = u[t, best, :], sel[t, p] = best.
 *   select_bwd: du[t, k, :] = sum_p [sel[t, p] == k] d_hd[t, p, :].
 *   pool_bwd:   dy[t, :] += ..., da[t, k] = ...  (closed form of the softmax-pooling gradient via suffix sums).
 *   tanh / tanh_bwd: the attention net's activation (in place). */
int b200rec_comi_tanh(float* z, int64_t n, void* stream);
int b200rec_comi_tanh_bwd(const float* h, float* dh, int64_t n, void* stream);
int b200rec_comi_pool_fwd(const float* a, const float* y, const int32_t* seq_off, int B, int K, int D, float* u, float* M,
                          float* S, void* stream);
int b200rec_comi_select_fwd(const float* u, const float* traw, const int32_t* tok_b, const int32_t* tok_pos, int T, int LP,
                            int P, int K, int D, float* hd, int32_t* sel, void* stream);
int b200rec_comi_select_bwd(const float* d_hd, const int32_t* sel, int T, int P, int K, int D, float* du, void* stream);
int b200rec_comi_pool_bwd(const float* du, const float* u, const float* y, const float* a, const float* M, const float* S,
                          const int32_t* seq_off, int B, int K, int D, float* dy, float* da, void* stream);
/* REMI's routing regulariser on the same routing logits (REC/model/IDNet/remi.py:156-196, 356-372, + autograd):
 *   var2[t, k] = (variance over the n valid window positions of the routing weights of interest k at token t / D)^2,
 *   rr = sum(var2) / seq_off[B_real]  (mean over the valid positions; the caller sums), da[t, k] = d rr / d a[t, k]
 * (nullable: forward only; otherwise scratch = [T, K, 3] floats).  One scan per (sequence, interest) instead of the
 * reference's [B*L, K, L] routing tensor.  Sequences >= B_real (the dummy sequence of static-token mode) are skipped:
 * the caller zero-fills var2 / da. */
int b200rec_comi_rr(const float* a, const int32_t* seq_off, int B_real, int K, int D, float* var2, float* scratch,
                    float* da, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H */
