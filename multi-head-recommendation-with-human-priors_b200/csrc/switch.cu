// Prior-switch auxiliary heads (reference hstu.py:512-544, 731-805, layers.py:16-84): per prior
// category c a Linear(D -> 1) on the body output predicts "does any of the next P targets carry category c"; the loss is
// a weighted BCE (pos_weight = (1 - p_c) / p_c) or the asymmetric loss, taken over ALL [B, L] context positions.
//
// The reference body is dense, so the loss also covers PADDED positions.  A left-padded query row attends to no valid
// key: its attention output is 0, LN(0) = 0, so every block only adds its output bias:  y_pad = E[item] + P[pos] +
// sum_l b_o^l.  switch_rows() rebuilds exactly that row for padded positions and copies the jagged body output for valid
// ones; the backward routes the gradient of padded rows to the pad item's embedding row, the position row and every
// block's output bias, exactly as autograd does in the reference.
#include "common.cuh"

#define SW_MAXC 32

// out[(b * Ls + j), :] for position l = l0 + j of sequence b
template <int MAXV>
__global__ void __launch_bounds__(256) switch_rows_kernel(const float* __restrict__ y, const int32_t* __restrict__ tok_index,
                                                          const int64_t* __restrict__ items_idx,
                                                          const float* __restrict__ table,
                                                          const float* __restrict__ pos_emb,
                                                          const float* __restrict__ bo_sum, int B, int LP, int l0,
                                                          int Ls, int D4, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= (int64_t)B * Ls) return;
  const int b = (int)(r / Ls), l = l0 + (int)(r - (int64_t)b * Ls);
  const int t = tok_index[(int64_t)b * LP + l];
  const int64_t id = items_idx[(int64_t)b * LP + l];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int c = lane + 32 * u;
    if (c >= D4) continue;
    float v[4];
    if (t >= 0) {
      load4<float>(y + ((int64_t)t * D4 + c) * 4, v);
    } else {
      float e[4], p[4], bo[4];
      load4<float>(table + (id * D4 + c) * 4, e);
      load4<float>(pos_emb + ((int64_t)l * D4 + c) * 4, p);
      load4<float>(bo_sum + c * 4, bo);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = e[k] + p[k] + bo[k];
    }
    store4<float>(out + (r * D4 + c) * 4, v);
  }
}

extern "C" int b200rec_switch_rows(const float* y, const int32_t* tok_index, const int64_t* items_idx, const float* table,
                                   const float* pos_emb, const float* bo_sum, int B, int LP, int l0, int Ls, int D,
                                   float* out, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && l0 >= 0 && l0 + Ls <= LP, "switch_rows: bad D / range");
  if (B == 0 || Ls == 0) return 0;
  const int blocks = ceil_div_i((int64_t)B * Ls, 8);
  if (D <= 512)
    switch_rows_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(y, tok_index, items_idx, table, pos_emb, bo_sum, B, LP,
                                                                    l0, Ls, D / 4, out);
  else
    switch_rows_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>(y, tok_index, items_idx, table, pos_emb, bo_sum, B, LP,
                                                                     l0, Ls, D / 4, out);
  B200_LAUNCH_OK();
  return 0;
}

struct SwitchLossCfg {
  int mode;                 // 0 = BCE with pos_weight, 1 = asymmetric loss
  float gamma_pos, gamma_neg, clip, eps;
  float grad_norm;          // prior_switch_loss_weight / normaliser: d(total) / d(sum of element losses)
};

// one warp per row: logits for the active heads, element losses, correctness flags and d(total)/d(logit)
template <int MAXV>
__global__ void __launch_bounds__(256) switch_loss_kernel(const float* __restrict__ rows, int64_t R, int Ls, int l0,
                                                          int D4, const float* __restrict__ Wa,
                                                          const float* __restrict__ ba, int n_act,
                                                          const int32_t* __restrict__ head_cat,
                                                          const float* __restrict__ pos_w,
                                                          const int64_t* __restrict__ tags, int LP, int C_tag, int P,
                                                          SwitchLossCfg cfg, float* __restrict__ logits,
                                                          float* __restrict__ loss_el, float* __restrict__ correct,
                                                          float* __restrict__ dlogit) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int b = (int)(r / Ls), l = l0 + (int)(r - (int64_t)b * Ls);
  float x[MAXV][4];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int c = lane + 32 * u;
    if (c < D4) load4<float>(rows + (r * D4 + c) * 4, x[u]);
  }
  for (int a = 0; a < n_act; ++a) {
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      const int c = lane + 32 * u;
      if (c < D4) {
        float w[4];
        load4<float>(Wa + ((int64_t)a * D4 + c) * 4, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(x[u][k], w[k], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane != 0) continue;
    const float z = acc + ba[a];
    // target: any of the next P positions carries category head_cat[a]   (hstu.py:733-736, 763-766)
    const int cat = head_cat[a];
    bool tgt = false;
    for (int p = 0; p < P; ++p) tgt |= tags[((int64_t)b * LP + l + 1 + p) * C_tag + cat] != 0;
    const float t = tgt ? 1.f : 0.f;
    const float s = 1.f / (1.f + expf(-z));
    float le, dz;
    if (cfg.mode == 0) {
      // F.binary_cross_entropy_with_logits(z, t, pos_weight = pw): l = -[pw t log s + (1 - t) log(1 - s)]
      const float pw = pos_w[a];
      const float sp_neg = fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));     // softplus(-z) = -log s
      const float sp_pos = fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));      // softplus(z)  = -log(1 - s)
      le = pw * t * sp_neg + (1.f - t) * sp_pos;
      dz = -pw * t * (1.f - s) + (1.f - t) * s;
    } else {
      // layers.py:53-84 (AsymmetricLoss), per element; the caller sums over positions and averages over the batch
      const float ds = s * (1.f - s);
      if (tgt) {
        const float sc = fmaxf(s, cfg.eps);
        const float lg = logf(sc);
        const float om = 1.f - s;                                         // 1 - pt
        const float wgt = cfg.gamma_pos > 0.f ? powf(om, cfg.gamma_pos) : 1.f;
        le = -lg * wgt;
        const float dlg = s > cfg.eps ? 1.f / s : 0.f;
        const float dw = cfg.gamma_pos > 0.f ? -cfg.gamma_pos * powf(om, cfg.gamma_pos - 1.f) : 0.f;
        dz = -(dlg * wgt + lg * dw) * ds;
      } else {
        const float raw = 1.f - s + cfg.clip;
        const bool clipped = cfg.clip > 0.f ? raw > 1.f : false;
        const float m = cfg.clip > 0.f ? fminf(raw, 1.f) : 1.f - s;
        const float mc = fmaxf(m, cfg.eps);
        const float lg = logf(mc);
        const float om = 1.f - m;
        const float wgt = cfg.gamma_neg > 0.f ? powf(om, cfg.gamma_neg) : 1.f;
        le = -lg * wgt;
        const float dm = clipped ? 0.f : -ds;
        const float dlg = m > cfg.eps ? 1.f / m : 0.f;
        const float dw = cfg.gamma_neg > 0.f ? -cfg.gamma_neg * powf(om, cfg.gamma_neg - 1.f) : 0.f;
        dz = -(dlg * wgt + lg * dw) * dm;
      }
    }
    logits[r * n_act + a] = z;
    loss_el[r * n_act + a] = le;
    correct[r * n_act + a] = ((z >= 0.f) == tgt) ? 1.f : 0.f;
    dlogit[r * n_act + a] = dz * cfg.grad_norm;
  }
}

extern "C" int b200rec_switch_loss(const float* rows, int64_t R, int Ls, int l0, int D, const float* W_aux,
                                   const float* b_aux, int n_act, const int32_t* head_cat, const float* pos_w,
                                   const int64_t* tags, int LP, int C_tag, int P, int mode, float gamma_pos,
                                   float gamma_neg, float clip, float eps, float grad_norm, float* logits, float* loss_el,
                                   float* correct, float* dlogit, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && n_act >= 1 && n_act <= SW_MAXC && l0 + Ls + P <= LP,
                 "switch_loss: bad D / heads / window (l0 + Ls + P must fit the [L + P] row)");
  if (R == 0) return 0;
  SwitchLossCfg cfg;
  cfg.mode = mode; cfg.gamma_pos = gamma_pos; cfg.gamma_neg = gamma_neg; cfg.clip = clip; cfg.eps = eps;
  cfg.grad_norm = grad_norm;
  const int blocks = ceil_div_i(R, 8);
  if (D <= 512)
    switch_loss_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(rows, R, Ls, l0, D / 4, W_aux, b_aux, n_act, head_cat,
                                                                    pos_w, tags, LP, C_tag, P, cfg, logits, loss_el,
                                                                    correct, dlogit);
  else
    switch_loss_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>(rows, R, Ls, l0, D / 4, W_aux, b_aux, n_act, head_cat,
                                                                     pos_w, tags, LP, C_tag, P, cfg, logits, loss_el,
                                                                     correct, dlogit);
  B200_LAUNCH_OK();
  return 0;
}

// dW[a, d] = gscale * sum_r dlogit[r, a] * rows[r, d] ; db[a] = gscale * sum_r dlogit[r, a].  One thread per (a, d),
// rows summed in ascending order: deterministic.
__global__ void __launch_bounds__(128) switch_bwd_w_kernel(const float* __restrict__ dlogit,
                                                           const float* __restrict__ rows, int64_t R, int n_act, int D,
                                                           const float* __restrict__ gscale, float* __restrict__ dW,
                                                           float* __restrict__ db) {
  const int a = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const float gs = gscale ? *gscale : 1.f;
  if (d < D) {
    float acc = 0.f;
    for (int64_t r = 0; r < R; ++r) acc = fmaf(dlogit[r * n_act + a], rows[r * D + d], acc);
    dW[(int64_t)a * D + d] = acc * gs;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int64_t r = 0; r < R; ++r) acc += dlogit[r * n_act + a];
    db[a] = acc * gs;
  }
}

// d rows[r, :] = gscale * sum_a dlogit[r, a] * W[a, :], routed: valid position -> added to dy[token]; padded position ->
// pad_rows[r, :] (zero for valid positions), from which the caller feeds the embedding scatter, the position gradient
// and every block's output-bias gradient.
template <int MAXV>
__global__ void __launch_bounds__(256) switch_bwd_rows_kernel(const float* __restrict__ dlogit,
                                                              const float* __restrict__ Wa, int64_t R, int Ls, int l0,
                                                              int LP, int n_act, int D4,
                                                              const int32_t* __restrict__ tok_index,
                                                              const float* __restrict__ gscale, float* __restrict__ dy,
                                                              float* __restrict__ pad_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int b = (int)(r / Ls), l = l0 + (int)(r - (int64_t)b * Ls);
  const int t = tok_index[(int64_t)b * LP + l];
  const float gs = gscale ? *gscale : 1.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int c = lane + 32 * u;
    if (c >= D4) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < n_act; ++a) {
      const float g = dlogit[r * n_act + a];
      float w[4];
      load4<float>(Wa + ((int64_t)a * D4 + c) * 4, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(g, w[k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] *= gs;
    float zero[4] = {0.f, 0.f, 0.f, 0.f};
    if (t >= 0) {
      float cur[4];
      load4<float>(dy + ((int64_t)t * D4 + c) * 4, cur);
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[k] += acc[k];
      store4<float>(dy + ((int64_t)t * D4 + c) * 4, cur);      // one position per token: no write conflict
      store4<float>(pad_rows + (r * D4 + c) * 4, zero);
    } else {
      store4<float>(pad_rows + (r * D4 + c) * 4, acc);
    }
  }
}

extern "C" int b200rec_switch_bwd(const float* dlogit, const float* rows, const float* W_aux, int64_t R, int Ls, int l0,
                                  int LP, int n_act, int D, const int32_t* tok_index, const float* gscale,
                                  float* dW, float* db, float* dy, float* pad_rows, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && n_act >= 1 && n_act <= SW_MAXC, "switch_bwd: bad D / heads");
  if (R == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch_bwd_w_kernel<<<dim3(ceil_div_i(D, 128), n_act), 128, 0, st>>>(dlogit, rows, R, n_act, D, gscale, dW, db);
  if (dy != nullptr) {
    const int blocks = ceil_div_i(R, 8);
    if (D <= 512)
      switch_bwd_rows_kernel<4><<<blocks, 256, 0, st>>>(dlogit, W_aux, R, Ls, l0, LP, n_act, D / 4, tok_index, gscale, dy,
                                                        pad_rows);
    else
      switch_bwd_rows_kernel<16><<<blocks, 256, 0, st>>>(dlogit, W_aux, R, Ls, l0, LP, n_act, D / 4, tok_index, gscale, dy,
                                                         pad_rows);
  }
  B200_LAUNCH_OK();
  return 0;
}

// dpos[l, :] += sum_b pad_rows[(b, l - l0), :]   (one thread per (l, column), batch summed in ascending order)
__global__ void __launch_bounds__(128) switch_pos_grad_kernel(const float* __restrict__ pad_rows, int B, int Ls, int l0,
                                                              int D, float* __restrict__ dpos) {
  const int j = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += pad_rows[((int64_t)b * Ls + j) * D + d];
  dpos[(int64_t)(l0 + j) * D + d] += acc;
}

extern "C" int b200rec_switch_pos_grad(const float* pad_rows, int B, int Ls, int l0, int D, float* dpos, void* stream) {
  if (B == 0 || Ls == 0) return 0;
  switch_pos_grad_kernel<<<dim3(ceil_div_i(D, 128), Ls), 128, 0, (cudaStream_t)stream>>>(pad_rows, B, Ls, l0, D, dpos);
  B200_LAUNCH_OK();
  return 0;
}
