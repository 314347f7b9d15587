// Sampled-softmax (NCE) loss over deduplicated query rows.
// One block owns one query row (head, token): its cosine logits against the negative set are read
// once into shared memory and shared by every prediction offset p that uses this head; the
// false-negative filter (hstu.py:613-614) comes from a bit matrix built by the GEMM epilogue.
#include "common.cuh"

#define NCE_THREADS 256
#define NCE_MAXP 16
#define NCE_MAX_JOBS 16   // jobs per grouped launch of the row kernels (job table by value in the kernel parameters)

struct Stats {
  float m, s, w;  // max, sum exp(z-m), sum exp(z-m)*z
  int gt;         // #logits greater than the reference positive logit
  int cnt;        // #logits counted
};

__device__ __forceinline__ void stats_merge(Stats& a, const Stats& b) {
  float m = fmaxf(a.m, b.m);
  float fa = a.m == -INFINITY ? 0.f : __expf(a.m - m);
  float fb = b.m == -INFINITY ? 0.f : __expf(b.m - m);
  a.s = a.s * fa + b.s * fb;
  a.w = a.w * fa + b.w * fb;
  a.m = m;
  a.gt += b.gt;
  a.cnt += b.cnt;
}

__device__ __forceinline__ void stats_merge2(Stats& a, const Stats& b) {   // log2-domain variant
  float m = fmaxf(a.m, b.m);
  float fa = a.m == -INFINITY ? 0.f : exp2f(a.m - m);
  float fb = b.m == -INFINITY ? 0.f : exp2f(b.m - m);
  a.s = a.s * fa + b.s * fb;
  a.w = a.w * fa + b.w * fb;
  a.m = m;
  a.gt += b.gt;
  a.cnt += b.cnt;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Block reduction of log2-domain Stats: maximum first, ONE rescale per partial, then plain fixed-tree sums
// (a merge-by-merge tree spends ~25 instructions and 2 MUFU per step; this is ~1/3 of the instructions).
__device__ Stats block_stats2(Stats v, Stats* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = v.m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float f = v.m == -INFINITY ? 0.f : ex2_ftz(v.m - m);
  float s = v.s * f, w = v.w * f;
  int gt = v.gt, cnt = v.cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    w += __shfl_xor_sync(0xffffffffu, w, o);
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __syncthreads();
  if (lane == 0) {
    Stats r;
    r.m = m; r.s = s; r.w = w; r.gt = gt; r.cnt = cnt;
    red[wid] = r;
  }
  __syncthreads();
  float M = red[0].m;
  for (int i = 1; i < nw; ++i) M = fmaxf(M, red[i].m);
  Stats r;
  r.m = M; r.s = 0.f; r.w = 0.f; r.gt = 0; r.cnt = 0;
  for (int i = 0; i < nw; ++i) {
    const float fi = red[i].m == -INFINITY ? 0.f : ex2_ftz(red[i].m - M);
    r.s = fmaf(red[i].s, fi, r.s);
    r.w = fmaf(red[i].w, fi, r.w);
    r.gt += red[i].gt;
    r.cnt += red[i].cnt;
  }
  return r;
}

// deterministic block reduction of Stats (fixed tree), result broadcast
__device__ Stats block_stats(Stats v, Stats* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stats t;
    t.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    t.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    t.w = __shfl_xor_sync(0xffffffffu, v.w, o);
    t.gt = __shfl_xor_sync(0xffffffffu, v.gt, o);
    t.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o);
    // order the merge by lane so both partners compute the identical value
    if (lane & o) { Stats u = t; stats_merge(u, v); v = u; } else { stats_merge(v, t); }
  }
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  Stats r = red[0];
  for (int i = 1; i < nw; ++i) stats_merge(r, red[i]);
  return r;
}

template <typename TA>
__global__ void __launch_bounds__(NCE_THREADS)
nce_loss_fwd_kernel(const float* __restrict__ logits, int64_t ld_logits, int n_neg,
                    const uint32_t* __restrict__ same_bits, const TA* __restrict__ q_hat, int64_t ldq,
                    const TA* __restrict__ t_hat, int D, const int32_t* __restrict__ tok_b,
                    const int32_t* __restrict__ tok_pos, int LP, int P, uint32_t p_mask,
                    const uint8_t* __restrict__ tok_ok, int tok_ok_ld, int tok_ok_col,
                    const float* __restrict__ coef, const float* __restrict__ logit_scale, float* __restrict__ loss,
                    float* __restrict__ g0, float* __restrict__ dscale, int32_t* __restrict__ rank0,
                    int32_t* __restrict__ nvalid, TA* __restrict__ G, int64_t ldg) {
  extern __shared__ __align__(16) float sm[];
  float* z = sm;               // [n_neg]  scaled logits tau * cos
  float* qrow = sm + n_neg;    // [D]
  __shared__ Stats red_stats[NCE_THREADS / 32];
  __shared__ float red[40];
  __shared__ float s_pos[NCE_MAXP], s_lse[NCE_MAXP], s_coef[NCE_MAXP];
  __shared__ int s_valid[NCE_MAXP], s_masked[NCE_MAXP];
  __shared__ int s_any;

  const int t = blockIdx.x;
  const int b = tok_b[t], pos = tok_pos[t];
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f));  // clamp(0, ln 100)
  const int n_words = (n_neg + 31) >> 5;

  if (threadIdx.x < NCE_MAXP) {
    int p = threadIdx.x;
    int ok = 0;
    if (p < P && ((p_mask >> p) & 1u)) {
      int64_t r = (int64_t)b * LP + pos + 1 + p;
      ok = tok_ok[r * tok_ok_ld + tok_ok_col] != 0;
    }
    s_valid[p] = ok;
    s_masked[p] = 0;
    s_coef[p] = (p < P) ? coef[p] : 0.f;
  }
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (threadIdx.x < P && s_valid[threadIdx.x]) s_any = 1;
  __syncthreads();
  if (!s_any) {  // no prediction offset uses this row
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      loss[(int64_t)t * P + p] = 0.f;
      g0[(int64_t)t * P + p] = 0.f;
      dscale[(int64_t)t * P + p] = 0.f;
      rank0[(int64_t)t * P + p] = -1;
      nvalid[(int64_t)t * P + p] = 0;
    }
    if (G)
      for (int j = threadIdx.x; j < n_neg; j += blockDim.x) G[(int64_t)t * ldg + j] = from_f32<TA>(0.f);
    return;
  }

  // stage the query row and the scaled logits row
  for (int d = threadIdx.x; d < D; d += blockDim.x) qrow[d] = to_f32(q_hat[(int64_t)t * ldq + d]);
  for (int j = threadIdx.x; j < n_neg; j += blockDim.x) z[j] = tau * logits[(int64_t)t * ld_logits + j];
  __syncthreads();

  // positive logits and "row has filtered negatives" flags
  for (int p = 0; p < P; ++p) {
    if (!s_valid[p]) continue;
    int64_t r = (int64_t)b * LP + pos + 1 + p;
    float acc = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) acc += qrow[d] * to_f32(t_hat[r * D + d]);
    acc = block_sum(acc, red);
    uint32_t any = 0;
    for (int w = threadIdx.x; w < n_words; w += blockDim.x) any |= same_bits[r * n_words + w];
    if (any) atomicOr(&s_masked[p], 1);
    if (threadIdx.x == 0) s_pos[p] = tau * acc;
    __syncthreads();
  }

  // unfiltered row statistics (shared by every offset without filtered negatives)
  const float pos0 = s_valid[0] ? s_pos[0] : INFINITY;
  Stats all;
  all.m = -INFINITY; all.s = 0.f; all.w = 0.f; all.gt = 0; all.cnt = 0;
  {
    float m = -INFINITY;
    for (int j = threadIdx.x; j < n_neg; j += blockDim.x) m = fmaxf(m, z[j]);
    Stats loc;
    loc.m = m; loc.s = 0.f; loc.w = 0.f; loc.gt = 0; loc.cnt = 0;
    for (int j = threadIdx.x; j < n_neg; j += blockDim.x) {
      float e = __expf(z[j] - m);
      loc.s += e;
      loc.w += e * z[j];
      loc.gt += z[j] > pos0;
      loc.cnt += 1;
    }
    all = block_stats(loc, red_stats);
  }

  float a_common = 0.f;  // sum over unfiltered offsets of coef_p * exp(m_all - lse_p)
  for (int p = 0; p < P; ++p) {
    if (!s_valid[p]) {
      if (threadIdx.x == 0) {
        loss[(int64_t)t * P + p] = 0.f;
        g0[(int64_t)t * P + p] = 0.f;
        dscale[(int64_t)t * P + p] = 0.f;
        rank0[(int64_t)t * P + p] = -1;
        nvalid[(int64_t)t * P + p] = 0;
      }
      continue;
    }
    Stats st = all;
    if (s_masked[p]) {  // exact pass over the un-filtered negatives of this target
      const uint32_t* bits = same_bits + ((int64_t)b * LP + pos + 1 + p) * n_words;
      float m = -INFINITY;
      for (int j = threadIdx.x; j < n_neg; j += blockDim.x)
        if (!((bits[j >> 5] >> (j & 31)) & 1u)) m = fmaxf(m, z[j]);
      Stats loc;
      loc.m = m; loc.s = 0.f; loc.w = 0.f; loc.gt = 0; loc.cnt = 0;
      for (int j = threadIdx.x; j < n_neg; j += blockDim.x)
        if (!((bits[j >> 5] >> (j & 31)) & 1u)) {
          float e = __expf(z[j] - m);
          loc.s += e;
          loc.w += e * z[j];
          loc.gt += z[j] > s_pos[p];
          loc.cnt += 1;
        }
      st = block_stats(loc, red_stats);
    }
    const float zp = s_pos[p];
    const float M = fmaxf(st.m, zp);
    const float en = st.m == -INFINITY ? 0.f : __expf(st.m - M);
    const float ep = __expf(zp - M);
    const float denom = st.s * en + ep;
    const float lse = M + __logf(denom);
    const float sm0 = ep / denom;
    const float c = s_coef[p];
    if (threadIdx.x == 0) {
      s_lse[p] = lse;
      loss[(int64_t)t * P + p] = c * (lse - zp);
      g0[(int64_t)t * P + p] = c * (sm0 - 1.f);
      // d loss / d logit_scale = sum_k softmax_k z_k - z_0
      dscale[(int64_t)t * P + p] = c * ((st.w * en + ep * zp) / denom - zp);
      rank0[(int64_t)t * P + p] = (p == 0 || s_masked[p]) ? st.gt : -1;
      nvalid[(int64_t)t * P + p] = st.cnt + 1;
    }
    if (!s_masked[p]) a_common += c * __expf(all.m - lse);
  }
  if (G == nullptr) return;
  __syncthreads();
  // gradient w.r.t. the cosine logits:  G[j] = tau * sum_p coef_p * softmax_p[j]
  for (int j = threadIdx.x; j < n_neg; j += blockDim.x) {
    float g = a_common * __expf(z[j] - all.m);
    for (int p = 0; p < P; ++p) {
      if (s_valid[p] && s_masked[p]) {
        const uint32_t* bits = same_bits + ((int64_t)b * LP + pos + 1 + p) * n_words;
        if (!((bits[j >> 5] >> (j & 31)) & 1u)) g += s_coef[p] * __expf(z[j] - s_lse[p]);
      }
    }
    G[(int64_t)t * ldg + j] = from_f32<TA>(tau * g);
  }
}


// ---------------------------------------------------------------------------------------------
// Fast path (n_neg <= 8192, n_neg % 4 == 0): positive cosines come from a warp-per-(t,p) pre-kernel,
// "target row has filtered negatives" flags from the bit-GEMM epilogue, and the logits row lives in
// registers (32 per thread) between the statistics pass and the gradient pass: one global read of
// the row, one block reduction, no shared-memory staging.
template <typename TA>
__global__ void __launch_bounds__(256) nce_pos_kernel(const TA* __restrict__ q_hat, int64_t ldq,
                                                      const TA* __restrict__ t_hat, int D4,
                                                      const int32_t* __restrict__ tok_b,
                                                      const int32_t* __restrict__ tok_pos, int T, int LP, int P,
                                                      uint32_t p_mask, const uint8_t* __restrict__ tok_ok,
                                                      int tok_ok_ld, int tok_ok_col, float* __restrict__ pos_cos) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)T * P) return;
  const int t = (int)(w / P), p = (int)(w - (int64_t)t * P);
  float acc = 0.f;
  bool ok = false;
  if ((p_mask >> p) & 1u) {
    const int64_t r = (int64_t)tok_b[t] * LP + tok_pos[t] + 1 + p;
    ok = tok_ok[r * tok_ok_ld + tok_ok_col] != 0;
    if (ok) {
      for (int c = lane; c < D4; c += 32) {
        float a[4], b[4];
        load4<TA>(q_hat + (int64_t)t * ldq + c * 4, a);
        load4<TA>(t_hat + (r * D4 + c) * 4, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += a[k] * b[k];
      }
      acc = warp_sum(acc);
    }
  }
  if (lane == 0) pos_cos[w] = ok ? acc : __int_as_float(0x7fc00000);  // NaN marks "offset not served"
}

#define NCE_RV 8  // float4 vectors per thread -> up to 8192 negatives per row

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename TA, bool FULL>   // FULL: n_neg == 8192 exactly (every thread owns 8 vectors, no range predicates)
__global__ void __launch_bounds__(256, 4)
nce_loss_fwd_reg_kernel(const float* __restrict__ logits, int64_t ld_logits, int n_neg,
                        const uint32_t* __restrict__ same_bits, const uint8_t* __restrict__ row_any,
                        const float* __restrict__ pos_cos, const int32_t* __restrict__ tok_b,
                        const int32_t* __restrict__ tok_pos, int LP, int P, const float* __restrict__ coef,
                        const float* __restrict__ logit_scale, float* __restrict__ loss, float* __restrict__ g0,
                        float* __restrict__ dscale, int32_t* __restrict__ rank0, int32_t* __restrict__ nvalid,
                        TA* __restrict__ G, int64_t ldg) {
  __shared__ Stats red_stats[8];
  __shared__ float s_pos[NCE_MAXP], s_lse[NCE_MAXP], s_coef[NCE_MAXP];
  __shared__ int s_valid[NCE_MAXP], s_masked[NCE_MAXP];
  __shared__ int s_any, s_anym;
  const int t = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)tok_b[t] * LP + tok_pos[t] + 1;
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f));
  const int n_words = (n_neg + 31) >> 5;
  const int n_vec = n_neg >> 2;
  if (tid == 0) { s_any = 0; s_anym = 0; }
  __syncthreads();
  if (tid < NCE_MAXP) {
    int p = tid;
    float pc = p < P ? pos_cos[(int64_t)t * P + p] : __int_as_float(0x7fc00000);
    int ok = pc == pc;
    s_valid[p] = ok;
    s_pos[p] = ok ? tau * pc : 0.f;
    s_masked[p] = (ok && row_any[r0 + p]) ? 1 : 0;
    s_coef[p] = p < P ? coef[p] : 0.f;
    if (ok) s_any = 1;
    if (s_masked[p]) s_anym = 1;
  }
  __syncthreads();
  if (!s_any) {
    for (int p = tid; p < P; p += blockDim.x) {
      loss[(int64_t)t * P + p] = 0.f;
      g0[(int64_t)t * P + p] = 0.f;
      dscale[(int64_t)t * P + p] = 0.f;
      rank0[(int64_t)t * P + p] = -1;
      nvalid[(int64_t)t * P + p] = 0;
    }
    if (G) {
      float zero[4] = {0.f, 0.f, 0.f, 0.f};
      for (int v = tid; v < n_vec; v += 256) store4<TA>(G + (int64_t)t * ldg + v * 4, zero);
    }
    return;
  }
  // the row: thread owns vectors tid, tid+256, ...   z2 = tau * log2(e) * cos  (exp(z - m) == exp2(z2 - m2))
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const float tau2 = tau * LOG2E;
  const bool lean = !s_anym;   // block-uniform: no offset of this row filters negatives (the common case)
  float z[NCE_RV][4];
  float mraw = -INFINITY;
  int n_mine = 0;
#pragma unroll
  for (int k = 0; k < NCE_RV; ++k) {
    int v = tid + k * 256;
    if (FULL || v < n_vec) {
      load4<float>(logits + (int64_t)t * ld_logits + v * 4, z[k]);
#pragma unroll
      for (int e = 0; e < 4; ++e) mraw = fmaxf(mraw, z[k][e]);
      n_mine += 4;
    }
  }
  const float m2 = mraw * tau2;            // tau2 > 0: max commutes with the scaling
  const float pos0_2 = s_valid[0] ? s_pos[0] * LOG2E : INFINITY;
  Stats loc;   // in the log2 domain: m = max z2, s = sum 2^(z2-m), w = sum 2^(z2-m) * z2
  loc.m = m2; loc.s = 0.f; loc.w = 0.f; loc.gt = 0; loc.cnt = n_mine;
  if (lean) {
    // ~9 instructions per logit: the register row is overwritten by 2^(z2 - m2), which is all the
    // gradient pass needs (g_j = const * 2^(z2_j - m2))
    const float thr = s_valid[0] ? s_pos[0] * LOG2E / tau2 : INFINITY;   // rank threshold on the raw cosine
    float wraw = 0.f;
#pragma unroll
    for (int k = 0; k < NCE_RV; ++k) {
      int v = tid + k * 256;
      if (FULL || v < n_vec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float raw = z[k][e];
          const float ex = ex2_approx(fmaf(raw, tau2, -m2));
          loc.s += ex;
          wraw = fmaf(ex, raw, wraw);
          loc.gt += raw > thr;
          z[k][e] = ex;
        }
      }
    }
    loc.w = wraw * tau2;
  } else {
#pragma unroll
    for (int k = 0; k < NCE_RV; ++k) {
      int v = tid + k * 256;
      if (FULL || v < n_vec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          z[k][e] *= tau2;
          float ex = exp2f(z[k][e] - m2);
          loc.s += ex;
          loc.w = fmaf(ex, z[k][e], loc.w);
          loc.gt += z[k][e] > pos0_2;
        }
      }
    }
  }
  const Stats all = block_stats2(loc, red_stats);

  // exact pass for offsets whose target filters some negatives (rare): block-wide, one offset at a time
  __shared__ Stats s_st[NCE_MAXP];
  for (int p = 0; p < P && !lean; ++p) {
    if (!(s_valid[p] && s_masked[p])) continue;   // block-uniform
    const uint32_t* bits = same_bits + (r0 + p) * n_words;
    const float posp_2 = s_pos[p] * LOG2E;
    float mm = -INFINITY;
#pragma unroll
    for (int k = 0; k < NCE_RV; ++k) {
      int v = tid + k * 256;
      if (FULL || v < n_vec) {
        uint32_t wbits = bits[(v * 4) >> 5] >> ((v * 4) & 31);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (!((wbits >> e) & 1u)) mm = fmaxf(mm, z[k][e]);
      }
    }
    Stats l2;
    l2.m = mm; l2.s = 0.f; l2.w = 0.f; l2.gt = 0; l2.cnt = 0;
#pragma unroll
    for (int k = 0; k < NCE_RV; ++k) {
      int v = tid + k * 256;
      if (FULL || v < n_vec) {
        uint32_t wbits = bits[(v * 4) >> 5] >> ((v * 4) & 31);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (!((wbits >> e) & 1u)) {
            float ex = exp2f(z[k][e] - mm);
            l2.s += ex;
            l2.w = fmaf(ex, z[k][e], l2.w);
            l2.gt += z[k][e] > posp_2;
            l2.cnt += 1;
          }
      }
    }
    const Stats st = block_stats2(l2, red_stats);
    if (tid == 0) s_st[p] = st;
  }
  __syncthreads();
  // per-offset scalars: thread p owns offset p (instead of every thread repeating all of them)
  __shared__ float s_acoef[NCE_MAXP];
  if (tid < P) {
    const int p = tid;
    float acoef = 0.f;
    if (!s_valid[p]) {
      loss[(int64_t)t * P + p] = 0.f;
      g0[(int64_t)t * P + p] = 0.f;
      dscale[(int64_t)t * P + p] = 0.f;
      rank0[(int64_t)t * P + p] = -1;
      nvalid[(int64_t)t * P + p] = 0;
    } else {
      const Stats st = s_masked[p] ? s_st[p] : all;
      const float zp2 = s_pos[p] * LOG2E;               // positive logit, log2 domain
      const float M = fmaxf(st.m, zp2);
      const float en = st.m == -INFINITY ? 0.f : exp2f(st.m - M);
      const float ep = exp2f(zp2 - M);
      const float denom = st.s * en + ep;
      const float lse2 = M + __log2f(denom);
      const float c = s_coef[p];
      s_lse[p] = lse2;
      loss[(int64_t)t * P + p] = c * (lse2 - zp2) * LN2;
      g0[(int64_t)t * P + p] = c * (__fdividef(ep, denom) - 1.f);
      // d loss / d logit_scale = sum_k softmax_k z_k - z_0   (natural-log logits = z2 * ln2)
      dscale[(int64_t)t * P + p] = c * (__fdividef(st.w * en + ep * zp2, denom) - zp2) * LN2;
      rank0[(int64_t)t * P + p] = (p == 0 || s_masked[p]) ? st.gt : -1;
      nvalid[(int64_t)t * P + p] = st.cnt + 1;
      if (!s_masked[p]) acoef = c * exp2f(all.m - lse2);
    }
    s_acoef[p] = acoef;
  }
  if (G == nullptr) return;
  __syncthreads();
  float a_common = 0.f;  // sum over unfiltered offsets of coef_p * exp(m_all - lse_p)
  int any_masked = 0;
  for (int p = 0; p < P; ++p) {
    a_common += s_acoef[p];
    any_masked |= (s_valid[p] && s_masked[p]);
  }
  a_common *= tau;
  if (lean) {
    const float f = a_common * ex2_approx(m2 - all.m);   // per-thread rescale from the local to the row maximum
#pragma unroll
    for (int k = 0; k < NCE_RV; ++k) {
      int v = tid + k * 256;
      if (FULL || v < n_vec) {
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) g[e] = f * z[k][e];
        store4<TA>(G + (int64_t)t * ldg + v * 4, g);
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < NCE_RV; ++k) {
    int v = tid + k * 256;
    if (FULL || v < n_vec) {
      float g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) g[e] = a_common * exp2f(z[k][e] - all.m);
      if (any_masked) {
        for (int p = 0; p < P; ++p) {
          if (s_valid[p] && s_masked[p]) {
            const uint32_t* bits = same_bits + (r0 + p) * n_words;
            uint32_t wbits = bits[(v * 4) >> 5] >> ((v * 4) & 31);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!((wbits >> e) & 1u)) g[e] += tau * s_coef[p] * exp2f(z[k][e] - s_lse[p]);
          }
        }
      }
      store4<TA>(G + (int64_t)t * ldg + v * 4, g);
    }
  }
}

int b200rec_nce_loss_fwd(const float* logits, int64_t ld_logits, int n_neg, const uint32_t* same_bits,
                         const uint8_t* row_any, float* pos_ws, const void* q_hat, int64_t ldq, const void* t_hat, int act_dtype, int D,
                         const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask,
                         const uint8_t* tok_ok, int tok_ok_ld, int tok_ok_col, const float* coef,
                         const float* logit_scale, float* loss, float* g0, float* dscale, int32_t* rank0,
                         int32_t* nvalid, void* G, int64_t ldg, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP, "nce_loss_fwd: pred_len %d not in [1,%d]", P, NCE_MAXP);
  if (T == 0) return 0;
  if (row_any != nullptr && pos_ws != nullptr && n_neg % 4 == 0 && n_neg <= NCE_RV * 256 * 4 &&
      ld_logits % 4 == 0 && ldg % 4 == 0 && D % 4 == 0 && ldq % 4 == 0) {
    DISPATCH_ACT(act_dtype, TA, {
      nce_pos_kernel<TA><<<ceil_div_i((int64_t)T * P, 8), 256, 0, (cudaStream_t)stream>>>(
          (const TA*)q_hat, ldq, (const TA*)t_hat, D / 4, tok_b, tok_pos, T, LP, P, p_mask, tok_ok, tok_ok_ld,
          tok_ok_col, pos_ws);
      if (n_neg == NCE_RV * 256 * 4)
        nce_loss_fwd_reg_kernel<TA, true><<<T, 256, 0, (cudaStream_t)stream>>>(
            logits, ld_logits, n_neg, same_bits, row_any, pos_ws, tok_b, tok_pos, LP, P, coef, logit_scale, loss, g0,
            dscale, rank0, nvalid, (TA*)G, ldg);
      else
        nce_loss_fwd_reg_kernel<TA, false><<<T, 256, 0, (cudaStream_t)stream>>>(
            logits, ld_logits, n_neg, same_bits, row_any, pos_ws, tok_b, tok_pos, LP, P, coef, logit_scale, loss, g0,
            dscale, rank0, nvalid, (TA*)G, ldg);
    });
    B200_LAUNCH_OK();
    return 0;
  }
  size_t smem = (size_t)(n_neg + D) * sizeof(float);
  B200_CHECK_ARG(smem <= 200 * 1024, "nce_loss_fwd: n_neg=%d too large for the shared-memory row cache", n_neg);
  DISPATCH_ACT(act_dtype, TA, {
    { static bool once_1 = false; if (!once_1) { B200_CUDA_OK(cudaFuncSetAttribute(nce_loss_fwd_kernel<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem)); once_1 = true; } }
    nce_loss_fwd_kernel<TA><<<T, NCE_THREADS, smem, (cudaStream_t)stream>>>(
        logits, ld_logits, n_neg, same_bits, (const TA*)q_hat, ldq, (const TA*)t_hat, D, tok_b, tok_pos, LP, P,
        p_mask, tok_ok, tok_ok_ld, tok_ok_col, coef, logit_scale, loss, g0, dscale, rank0, nvalid, (TA*)G, ldg);
  });
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// REMI's interest-aware hard negatives (REC/model/IDNet/remi.py:198-277 with beta > 0, + autograd): per (token, offset)
//     loss = logaddexp(z_pos, LSE_j((beta + 1) z_j) - LSE_j(beta z_j) + log N) - z_pos,     z = tau * cos,
// the sums over the negatives the false-negative filter keeps, N = ALL negatives of the row (remi.py:244).  With
// sigma = e^(log_neg - lse) (the mass of the negative term), p1 / p2 = softmax of (beta + 1) z / beta z over the kept
// negatives:  d loss / d z_pos = -sigma,  d loss / d z_j = sigma ((beta + 1) p1_j - beta p2_j),
//             d loss / d log tau = sigma ((beta + 1) <p1, z> - beta <p2, z> - z_pos).
// Same block-per-query-row layout, inputs and outputs as nce_loss_fwd_kernel (rank0 / nvalid feed the same top-k
// logging); every offset takes the filtered pass (the two temperatures leave nothing to share between offsets).
template <typename TA>
__global__ void __launch_bounds__(NCE_THREADS)
nce_ihn_loss_fwd_kernel(const float* __restrict__ logits, int64_t ld_logits, int n_neg,
                        const uint32_t* __restrict__ same_bits, const TA* __restrict__ q_hat, int64_t ldq,
                        const TA* __restrict__ t_hat, int D, const int32_t* __restrict__ tok_b,
                        const int32_t* __restrict__ tok_pos, int LP, int P, uint32_t p_mask,
                        const uint8_t* __restrict__ tok_ok, int tok_ok_ld, int tok_ok_col,
                        const float* __restrict__ coef, const float* __restrict__ logit_scale, float beta,
                        float* __restrict__ loss, float* __restrict__ g0, float* __restrict__ dscale,
                        int32_t* __restrict__ rank0, int32_t* __restrict__ nvalid, TA* __restrict__ G, int64_t ldg) {
  extern __shared__ __align__(16) float sm[];
  float* z = sm;               // [n_neg]  scaled logits tau * cos
  float* qrow = sm + n_neg;    // [D]
  __shared__ Stats red_stats[NCE_THREADS / 32];
  __shared__ float red[40];
  __shared__ float s_pos[NCE_MAXP], s_coef[NCE_MAXP], s_c1[NCE_MAXP], s_c2[NCE_MAXP], s_ln1[NCE_MAXP], s_ln2[NCE_MAXP];
  __shared__ int s_valid[NCE_MAXP], s_live[NCE_MAXP];
  __shared__ int s_any;

  const int t = blockIdx.x;
  const int b = tok_b[t], pos = tok_pos[t];
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f));  // clamp(0, ln 100)
  const int n_words = (n_neg + 31) >> 5;

  if (threadIdx.x < NCE_MAXP) {
    int p = threadIdx.x;
    int ok = 0;
    if (p < P && ((p_mask >> p) & 1u)) {
      int64_t r = (int64_t)b * LP + pos + 1 + p;
      ok = tok_ok[r * tok_ok_ld + tok_ok_col] != 0;
    }
    s_valid[p] = ok;
    s_live[p] = 0;
    s_coef[p] = (p < P) ? coef[p] : 0.f;
  }
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (threadIdx.x < P && s_valid[threadIdx.x]) s_any = 1;
  __syncthreads();
  if (!s_any) {  // no prediction offset uses this row
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      loss[(int64_t)t * P + p] = 0.f;
      g0[(int64_t)t * P + p] = 0.f;
      dscale[(int64_t)t * P + p] = 0.f;
      rank0[(int64_t)t * P + p] = -1;
      nvalid[(int64_t)t * P + p] = 0;
    }
    if (G)
      for (int j = threadIdx.x; j < n_neg; j += blockDim.x) G[(int64_t)t * ldg + j] = from_f32<TA>(0.f);
    return;
  }

  for (int d = threadIdx.x; d < D; d += blockDim.x) qrow[d] = to_f32(q_hat[(int64_t)t * ldq + d]);
  for (int j = threadIdx.x; j < n_neg; j += blockDim.x) z[j] = tau * logits[(int64_t)t * ld_logits + j];
  __syncthreads();

  for (int p = 0; p < P; ++p) {  // positive logits
    if (!s_valid[p]) continue;
    int64_t r = (int64_t)b * LP + pos + 1 + p;
    float acc = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) acc += qrow[d] * to_f32(t_hat[r * D + d]);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) s_pos[p] = tau * acc;
    __syncthreads();
  }

  const float b1 = beta + 1.f;
  for (int p = 0; p < P; ++p) {
    if (!s_valid[p]) {
      if (threadIdx.x == 0) {
        loss[(int64_t)t * P + p] = 0.f;
        g0[(int64_t)t * P + p] = 0.f;
        dscale[(int64_t)t * P + p] = 0.f;
        rank0[(int64_t)t * P + p] = -1;
        nvalid[(int64_t)t * P + p] = 0;
      }
      continue;
    }
    const uint32_t* bits = same_bits + ((int64_t)b * LP + pos + 1 + p) * n_words;
    const float zp = s_pos[p];
    float m = -INFINITY;
    for (int j = threadIdx.x; j < n_neg; j += blockDim.x)
      if (!((bits[j >> 5] >> (j & 31)) & 1u)) m = fmaxf(m, z[j]);
    Stats l1, l2;                      // temperature (beta + 1) and beta; w accumulates e * z
    l1.m = b1 * m; l1.s = 0.f; l1.w = 0.f; l1.gt = 0; l1.cnt = 0;      // -inf stays -inf (beta > 0)
    l2.m = beta * m; l2.s = 0.f; l2.w = 0.f; l2.gt = 0; l2.cnt = 0;
    for (int j = threadIdx.x; j < n_neg; j += blockDim.x)
      if (!((bits[j >> 5] >> (j & 31)) & 1u)) {
        const float d = z[j] - m;
        const float e1 = expf(b1 * d), e2 = expf(beta * d);
        l1.s += e1;
        l1.w += e1 * z[j];
        l1.gt += z[j] > zp;
        l1.cnt += 1;
        l2.s += e2;
        l2.w += e2 * z[j];
      }
    const Stats st1 = block_stats(l1, red_stats);
    const Stats st2 = block_stats(l2, red_stats);
    const float c = s_coef[p];
    float lossv = 0.f, g0v = 0.f, dsv = 0.f, c1 = 0.f, c2 = 0.f, ln1 = 0.f, ln2 = 0.f;
    int live = 0;
    if (st1.cnt > 0) {                 // no kept negative: log_neg = -inf, the loss and every gradient are 0
      ln1 = st1.m + logf(st1.s);
      ln2 = st2.m + logf(st2.s);
      const float log_neg = ln1 - ln2 + logf((float)n_neg);
      const float M = fmaxf(zp, log_neg);
      const float lse = M + logf(expf(zp - M) + expf(log_neg - M));
      const float sig = expf(log_neg - lse);
      lossv = c * (lse - zp);
      g0v = -c * sig;
      dsv = c * sig * (b1 * st1.w / st1.s - beta * st2.w / st2.s - zp);
      c1 = c * sig * b1;
      c2 = c * sig * beta;
      live = 1;
    }
    if (threadIdx.x == 0) {
      loss[(int64_t)t * P + p] = lossv;
      g0[(int64_t)t * P + p] = g0v;
      dscale[(int64_t)t * P + p] = dsv;
      rank0[(int64_t)t * P + p] = st1.gt;
      nvalid[(int64_t)t * P + p] = st1.cnt + 1;
      s_c1[p] = c1; s_c2[p] = c2; s_ln1[p] = ln1; s_ln2[p] = ln2;
      s_live[p] = live;
    }
  }
  if (G == nullptr) return;
  __syncthreads();
  // gradient w.r.t. the cosine logits:  G[j] = tau * sum_p coef_p sigma_p ((beta + 1) p1_p[j] - beta p2_p[j])
  for (int j = threadIdx.x; j < n_neg; j += blockDim.x) {
    float g = 0.f;
    for (int p = 0; p < P; ++p) {
      if (s_live[p]) {
        const uint32_t* bits = same_bits + ((int64_t)b * LP + pos + 1 + p) * n_words;
        if (!((bits[j >> 5] >> (j & 31)) & 1u))
          g += s_c1[p] * expf(b1 * z[j] - s_ln1[p]) - s_c2[p] * expf(beta * z[j] - s_ln2[p]);
      }
    }
    G[(int64_t)t * ldg + j] = from_f32<TA>(tau * g);
  }
}

extern "C" int b200rec_nce_ihn_loss_fwd(const float* logits, int64_t ld_logits, int n_neg, const uint32_t* same_bits,
                             const void* q_hat, int64_t ldq, const void* t_hat, int act_dtype, int D,
                             const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask,
                             const uint8_t* tok_ok, int tok_ok_ld, int tok_ok_col, const float* coef,
                             const float* logit_scale, float beta, float* loss, float* g0, float* dscale,
                             int32_t* rank0, int32_t* nvalid, void* G, int64_t ldg, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP, "nce_ihn_loss_fwd: pred_len %d not in [1,%d]", P, NCE_MAXP);
  B200_CHECK_ARG(beta > 0.f && n_neg >= 1, "nce_ihn_loss_fwd: beta %f must be > 0 (beta <= 0 is b200rec_nce_loss_fwd)", beta);
  if (T == 0) return 0;
  size_t smem = (size_t)(n_neg + D) * sizeof(float);
  B200_CHECK_ARG(smem <= 200 * 1024, "nce_ihn_loss_fwd: n_neg=%d too large for the shared-memory row cache", n_neg);
  DISPATCH_ACT(act_dtype, TA, {
    {
      static size_t smem_set = 0;      // raised outside any capture by the eager warm-up of a graphed step
      if (smem > smem_set) {
        B200_CUDA_OK(cudaFuncSetAttribute(nce_ihn_loss_fwd_kernel<TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
      }
    }
    nce_ihn_loss_fwd_kernel<TA><<<T, NCE_THREADS, smem, (cudaStream_t)stream>>>(
        logits, ld_logits, n_neg, same_bits, (const TA*)q_hat, ldq, (const TA*)t_hat, D, tok_b, tok_pos, LP, P,
        p_mask, tok_ok, tok_ok_ld, tok_ok_col, coef, logit_scale, beta, loss, g0, dscale, rank0, nvalid, (TA*)G, ldg);
  });
  B200_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------ counts / coefs
__global__ void nce_count_kernel(const int32_t* __restrict__ tok_b, const int32_t* __restrict__ tok_pos, int T,
                                 int LP, int P, const uint8_t* __restrict__ tok_ok, int tok_ok_ld, int tok_ok_col,
                                 int32_t* __restrict__ cnt) {
  __shared__ int loc[NCE_MAXP];
  if (threadIdx.x < NCE_MAXP) loc[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T * P) {
    int t = i / P, p = i - t * P;
    int64_t r = (int64_t)tok_b[t] * LP + tok_pos[t] + 1 + p;
    if (tok_ok[r * tok_ok_ld + tok_ok_col]) atomicAdd(&loc[p], 1);
  }
  __syncthreads();
  if (threadIdx.x < P && loc[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], loc[threadIdx.x]);
}

int b200rec_nce_count(const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, const uint8_t* tok_ok,
                      int tok_ok_ld, int tok_ok_col, int32_t* cnt, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP, "nce_count: pred_len %d not in [1,%d]", P, NCE_MAXP);
  B200_CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * P, (cudaStream_t)stream));
  if (T == 0) return 0;
  nce_count_kernel<<<ceil_div_i((int64_t)T * P, 256), 256, 0, (cudaStream_t)stream>>>(tok_b, tok_pos, T, LP, P,
                                                                                    tok_ok, tok_ok_ld, tok_ok_col,
                                                                                    cnt);
  B200_LAUNCH_OK();
  return 0;
}

__global__ void nce_coef_kernel(const int32_t* __restrict__ cnt, const float* __restrict__ lam, float w, int P,
                                float* __restrict__ coef) {
  int p = threadIdx.x;
  if (p < P) coef[p] = lam[p] * w / fmaxf((float)cnt[p], 1.f);
}

extern "C" int b200rec_nce_coef(const int32_t* cnt, const float* lam, float w, int P, float* coef, void* stream) {
  nce_coef_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cnt, lam, w, P, coef);
  B200_LAUNCH_OK();
  return 0;
}

// All (validity column, loss weight) keys of a step at once: cnt[c, p] = #valid tokens of column c at offset p (one
// pass over the tokens for every column), coef[k, p] = lam[p] * w_k / max(cnt[col_k, p], 1).  3 launches instead of 3 per key.
#define NCE_MAX_COLS 32
struct NceCoefKeys { int col[NCE_MAX_JOBS * 2]; float w[NCE_MAX_JOBS * 2]; };

__global__ void nce_count_all_kernel(const int32_t* __restrict__ tok_b, const int32_t* __restrict__ tok_pos, int T, int LP,
                                     int P, const uint8_t* __restrict__ tok_ok, int n_col, int32_t* __restrict__ cnt) {
  __shared__ int loc[NCE_MAX_COLS * NCE_MAXP];
  for (int i = threadIdx.x; i < n_col * P; i += blockDim.x) loc[i] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T * P) {
    int t = i / P, p = i - t * P;
    int64_t r = (int64_t)tok_b[t] * LP + tok_pos[t] + 1 + p;
    for (int c = 0; c < n_col; ++c)
      if (tok_ok[r * n_col + c]) atomicAdd(&loc[c * P + p], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_col * P; k += blockDim.x)
    if (loc[k]) atomicAdd(&cnt[k], loc[k]);
}

__global__ void nce_coef_all_kernel(const int32_t* __restrict__ cnt, const float* __restrict__ lam,
                                    const __grid_constant__ NceCoefKeys keys, int n_keys, int P, float* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys * P) return;
  const int k = i / P, p = i - k * P;
  coef[i] = lam[p] * keys.w[k] / fmaxf((float)cnt[keys.col[k] * P + p], 1.f);
}

extern "C" int b200rec_nce_coefs(const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, const uint8_t* tok_ok,
                                 int n_col, const int32_t* key_col, const float* key_w, int n_keys, const float* lam,
                                 int32_t* cnt, float* coef, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP && n_col >= 1 && n_col <= NCE_MAX_COLS, "nce_coefs: P=%d n_col=%d", P, n_col);
  B200_CHECK_ARG(n_keys >= 1 && n_keys <= NCE_MAX_JOBS * 2, "nce_coefs: %d keys (1..%d)", n_keys, NCE_MAX_JOBS * 2);
  NceCoefKeys keys;
  for (int k = 0; k < n_keys; ++k) {
    B200_CHECK_ARG(key_col[k] >= 0 && key_col[k] < n_col, "nce_coefs: key column %d out of range", key_col[k]);
    keys.col[k] = key_col[k];
    keys.w[k] = key_w[k];
  }
  B200_CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * n_col * P, (cudaStream_t)stream));
  if (T > 0)
    nce_count_all_kernel<<<ceil_div_i((int64_t)T * P, 256), 256, 0, (cudaStream_t)stream>>>(tok_b, tok_pos, T, LP, P, tok_ok,
                                                                                          n_col, cnt);
  nce_coef_all_kernel<<<ceil_div_i(n_keys * P, 128), 128, 0, (cudaStream_t)stream>>>(cnt, lam, keys, n_keys, P, coef);
  B200_LAUNCH_OK();
  return 0;
}

// Top-k logging scalars of one job from the rank of the positive (hstu.py:621-629): over the rows with a served offset 0
// (nvalid[t, 0] > 0): out = { #rows, mean #logits, acc@1, @5, @10, @50, @100 }.  Integer sums -> exact in any order.
__global__ void __launch_bounds__(256) nce_topk_logs_kernel(const int32_t* __restrict__ rank0,
                                                            const int32_t* __restrict__ nvalid, int T, int P,
                                                            unsigned long long* __restrict__ acc, float* __restrict__ out) {
  __shared__ unsigned long long s[7];
  if (threadIdx.x < 7) s[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned long long loc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int nv = nvalid[(int64_t)t * P], r = rank0[(int64_t)t * P];
    if (nv > 0) {
      loc[0] += 1; loc[1] += (unsigned long long)nv;
      loc[2] += r < 1; loc[3] += r < 5; loc[4] += r < 10; loc[5] += r < 50; loc[6] += r < 100;
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (loc[k]) atomicAdd(&s[k], loc[k]);
  __syncthreads();
  if (threadIdx.x < 7 && s[threadIdx.x]) atomicAdd(&acc[threadIdx.x], s[threadIdx.x]);
  // the last block to finish turns the sums into the scalars
  __shared__ unsigned int last;
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd((unsigned int*)(acc + 7), 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last && threadIdx.x < 7) {
    __threadfence();
    const volatile unsigned long long* a = acc;
    const float den = fmaxf((float)a[0], 1.f);
    out[threadIdx.x] = threadIdx.x == 0 ? (float)a[0] : (float)a[threadIdx.x] / den;
  }
}

extern "C" int b200rec_nce_topk_logs(const int32_t* rank0, const int32_t* nvalid, int T, int P, void* acc_ws, float* out,
                                     void* stream) {
  B200_CUDA_OK(cudaMemsetAsync(acc_ws, 0, 8 * sizeof(unsigned long long), (cudaStream_t)stream));
  B200_CUDA_OK(cudaMemsetAsync(out, 0, 7 * sizeof(float), (cudaStream_t)stream));
  if (T == 0) return 0;
  nce_topk_logs_kernel<<<std::min(ceil_div_i(T, 256), 64), 256, 0, (cudaStream_t)stream>>>(
      rank0, nvalid, T, P, (unsigned long long*)acc_ws, out);
  B200_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------ positive-logit backward
struct NcePosBwdJob { const float* g0; const void* q_hat; float* d_qhat; };
struct NcePosBwdJobs { NcePosBwdJob j[NCE_MAX_JOBS]; };

// grid (T, jobs): the jobs of one launch write DISTINCT query-gradient slices (the caller groups them so)
template <typename TA>
__global__ void __launch_bounds__(256) nce_pos_bwd_q_kernel(const __grid_constant__ NcePosBwdJobs jobs,
                                                            const TA* __restrict__ t_hat, int D4,
                                                            const int32_t* __restrict__ tok_b,
                                                            const int32_t* __restrict__ tok_pos, int LP, int P,
                                                            const float* __restrict__ logit_scale,
                                                            const float* __restrict__ gscale, int64_t ldd) {
  const float* __restrict__ g0 = jobs.j[blockIdx.y].g0;
  float* __restrict__ d_qhat = jobs.j[blockIdx.y].d_qhat;
  const int t = blockIdx.x;
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f)) * (gscale ? *gscale : 1.f);
  const int64_t r0 = (int64_t)tok_b[t] * LP + tok_pos[t] + 1;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    float acc[4];
    load4<float>(d_qhat + (int64_t)t * ldd + c * 4, acc);
    for (int p = 0; p < P; ++p) {
      float g = g0[(int64_t)t * P + p];
      if (g != 0.f) {
        float v[4];
        load4<TA>(t_hat + ((r0 + p) * D4 + c) * 4, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += tau * g * v[k];
      }
    }
    store4<float>(d_qhat + (int64_t)t * ldd + c * 4, acc);
  }
}

static int nce_pos_bwd_q_launch(const NcePosBwdJobs& jobs, int n_jobs, const void* t_hat, int act_dtype, int D,
                                const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P,
                                const float* logit_scale, const float* gscale, int64_t ldd, void* stream) {
  if (T == 0) return 0;
  B200_CHECK_ARG(D % 4 == 0 && ldd % 4 == 0, "nce_pos_bwd_q: bad D/ldd");
  B200_CHECK_ARG(n_jobs >= 1 && n_jobs <= NCE_MAX_JOBS, "nce_pos_bwd_q: %d jobs (1..%d per launch)", n_jobs, NCE_MAX_JOBS);
  DISPATCH_ACT(act_dtype, TA, {
    nce_pos_bwd_q_kernel<TA><<<dim3(T, n_jobs), 256, 0, (cudaStream_t)stream>>>(jobs, (const TA*)t_hat, D / 4, tok_b,
                                                                              tok_pos, LP, P, logit_scale, gscale, ldd);
  });
  B200_LAUNCH_OK();
  return 0;
}

int b200rec_nce_pos_bwd_q(const float* g0, const void* t_hat, int act_dtype, int D, const int32_t* tok_b,
                          const int32_t* tok_pos, int T, int LP, int P, const float* logit_scale,
                          const float* gscale, float* d_qhat, int64_t ldd, void* stream) {
  NcePosBwdJobs jobs;
  jobs.j[0] = {g0, nullptr, d_qhat};
  return nce_pos_bwd_q_launch(jobs, 1, t_hat, act_dtype, D, tok_b, tok_pos, T, LP, P, logit_scale, gscale, ldd, stream);
}

extern "C" int b200rec_nce_pos_bwd_q_grouped(const b200rec_nce_pos_bwd_job* jobs_in, int n_jobs, const void* t_hat,
                                             int act_dtype, int D, const int32_t* tok_b, const int32_t* tok_pos, int T,
                                             int LP, int P, const float* logit_scale, const float* gscale, int64_t ldd,
                                             void* stream) {
  for (int j0 = 0; j0 < n_jobs; j0 += NCE_MAX_JOBS) {
    const int n = std::min(NCE_MAX_JOBS, n_jobs - j0);
    NcePosBwdJobs jobs;
    for (int j = 0; j < n; ++j) jobs.j[j] = {jobs_in[j0 + j].g0, jobs_in[j0 + j].q_hat, jobs_in[j0 + j].d_qhat};
    if (nce_pos_bwd_q_launch(jobs, n, t_hat, act_dtype, D, tok_b, tok_pos, T, LP, P, logit_scale, gscale, ldd, stream))
      return 1;
  }
  return 0;
}

// gather form: target row r = (b, pos_r) collects from queries at l = pos_r - 1 - p.  All jobs of the launch accumulate
// into the SAME target-gradient row, one after the other in job order inside the thread: deterministic, one launch.
template <typename TA>
__global__ void __launch_bounds__(256) nce_pos_bwd_t_kernel(const __grid_constant__ NcePosBwdJobs jobs, int n_jobs,
                                                            int64_t ldq, int D4, const int32_t* __restrict__ tok_index,
                                                            int LP, int P, const float* __restrict__ logit_scale,
                                                            const float* __restrict__ gscale,
                                                            float* __restrict__ d_that) {
  const int64_t r = blockIdx.x;
  const int b = (int)(r / LP), pos_r = (int)(r - (int64_t)b * LP);
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f)) * (gscale ? *gscale : 1.f);
  __shared__ int s_t[NCE_MAX_JOBS][NCE_MAXP];
  __shared__ float s_g[NCE_MAX_JOBS][NCE_MAXP];
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_jobs * P; idx += blockDim.x) {
    const int j = idx / P, p = idx - j * P;
    const int l = pos_r - 1 - p;
    const int t = l >= 0 ? tok_index[(int64_t)b * LP + l] : -1;
    const float g = t >= 0 ? jobs.j[j].g0[(int64_t)t * P + p] : 0.f;
    s_t[j][p] = g != 0.f ? t : -1;
    s_g[j][p] = g;
    if (g != 0.f) s_any = 1;
  }
  __syncthreads();
  if (!s_any) return;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    float acc[4];
    load4<float>(d_that + (r * D4 + c) * 4, acc);
    for (int j = 0; j < n_jobs; ++j) {
      const TA* __restrict__ q_hat = (const TA*)jobs.j[j].q_hat;
      for (int p = 0; p < P; ++p) {
        if (s_t[j][p] >= 0) {
          float v[4];
          load4<TA>(q_hat + (int64_t)s_t[j][p] * ldq + c * 4, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += tau * s_g[j][p] * v[k];
        }
      }
    }
    store4<float>(d_that + (r * D4 + c) * 4, acc);
  }
}

static int nce_pos_bwd_t_launch(const NcePosBwdJobs& jobs, int n_jobs, int64_t ldq, int act_dtype, int D,
                                const int32_t* tok_index, int B, int LP, int P, const float* logit_scale,
                                const float* gscale, float* d_that, void* stream) {
  if (B == 0) return 0;
  B200_CHECK_ARG(D % 4 == 0 && ldq % 4 == 0, "nce_pos_bwd_t: bad D/ldq");
  B200_CHECK_ARG(P <= NCE_MAXP, "nce_pos_bwd_t: pred_len too large");
  B200_CHECK_ARG(n_jobs >= 1 && n_jobs <= NCE_MAX_JOBS, "nce_pos_bwd_t: %d jobs (1..%d per launch)", n_jobs, NCE_MAX_JOBS);
  DISPATCH_ACT(act_dtype, TA, {
    nce_pos_bwd_t_kernel<TA><<<B * LP, 256, 0, (cudaStream_t)stream>>>(jobs, n_jobs, ldq, D / 4, tok_index, LP, P,
                                                                       logit_scale, gscale, d_that);
  });
  B200_LAUNCH_OK();
  return 0;
}

int b200rec_nce_pos_bwd_t(const float* g0, const void* q_hat, int64_t ldq, int act_dtype, int D,
                          const int32_t* tok_index, int B, int LP, int P, const float* logit_scale,
                          const float* gscale, float* d_that, void* stream) {
  NcePosBwdJobs jobs;
  jobs.j[0] = {g0, q_hat, nullptr};
  return nce_pos_bwd_t_launch(jobs, 1, ldq, act_dtype, D, tok_index, B, LP, P, logit_scale, gscale, d_that, stream);
}

extern "C" int b200rec_nce_pos_bwd_t_grouped(const b200rec_nce_pos_bwd_job* jobs_in, int n_jobs, int64_t ldq,
                                             int act_dtype, int D, const int32_t* tok_index, int B, int LP, int P,
                                             const float* logit_scale, const float* gscale, float* d_that,
                                             void* stream) {
  // chunks of NCE_MAX_JOBS run one after the other: the accumulation order stays the job order
  for (int j0 = 0; j0 < n_jobs; j0 += NCE_MAX_JOBS) {
    const int n = std::min(NCE_MAX_JOBS, n_jobs - j0);
    NcePosBwdJobs jobs;
    for (int j = 0; j < n; ++j) jobs.j[j] = {jobs_in[j0 + j].g0, jobs_in[j0 + j].q_hat, jobs_in[j0 + j].d_qhat};
    if (nce_pos_bwd_t_launch(jobs, n, ldq, act_dtype, D, tok_index, B, LP, P, logit_scale, gscale, d_that, stream))
      return 1;
  }
  return 0;
}


// =============================================================================================================
// Fused path (bf16 production mode): the logits GEMM writes softmax numerators E = 2^(tau2 cos - mref) and per-row
// partial sums (B200REC_EPI_NCE_EXP); nothing of size [T, Nneg] in fp32 ever reaches HBM and the old row kernel
// (logits read + G write) is gone.
//   nce_pos_ref  before the GEMM: positive cosines of every offset, the row's reference exponent mref and the p = 0
//                cosine (threshold of the top-k logging count)
//   nce_combine  after the GEMM: per-offset log-sum-exp / loss / positive-logit gradient / d(logit_scale) / rank from the
//                partial sums; offsets whose target filters negatives (hstu.py:613-614) subtract exactly those stored
//                numerators; the gradient w.r.t. the cosine logits is  G = row_scale[t] * E  (E patched at the few
//                filtered entries), consumed by the backward GEMMs as  dq = row_scale * (E n_hat),
//                dn = E^T (row_scale * q_hat)  -- the scaled query copy `qs` is written here.
// Range: mref = max over the row's valid offsets of the positive logit (log2 domain), floored at tau2 - 100, so
// E <= 2^100 (fp32 sums of 8192 terms cannot overflow) and the positive term never underflows against it; numerators
// more than 2^-126 below the reference flush to zero, i.e. softmax weights below 1e-38 are dropped.
// =============================================================================================================
// Grouped launches of the per-job row kernels (one job = one (negative set, head) pair of the loss): the 12 jobs of config
// B were 12 launches of ~5 us of work each, i.e. launch / tail bound; grid.y = job, the job table travels by value.
struct NcePosRefJob {
  const bf16* q_hat; uint32_t p_mask; int tok_ok_col; float* pos_cos; float* mref; float* thr;
};
struct NcePosRefJobs { NcePosRefJob j[NCE_MAX_JOBS]; };

template <int MAXV>
__global__ void __launch_bounds__(256) nce_pos_ref_kernel(const __grid_constant__ NcePosRefJobs jobs, int64_t ldq,
                                                          const bf16* __restrict__ t_hat, int D4,
                                                          const int32_t* __restrict__ tok_b,
                                                          const int32_t* __restrict__ tok_pos, int T, int LP, int P,
                                                          const uint8_t* __restrict__ tok_ok, int tok_ok_ld,
                                                          const float* __restrict__ logit_scale) {
  pdl_trigger();
  const NcePosRefJob& jb = jobs.j[blockIdx.y];
  const bf16* __restrict__ q_hat = jb.q_hat;
  const uint32_t p_mask = jb.p_mask;
  const int tok_ok_col = jb.tok_ok_col;
  float* __restrict__ pos_cos = jb.pos_cos;
  float* __restrict__ mref = jb.mref;
  float* __restrict__ thr = jb.thr;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const float tau2 = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f)) * 1.4426950408889634f;
  float q[MAXV][4];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int c = lane + 32 * u;
    if (c < D4) load4<bf16>(q_hat + (int64_t)t * ldq + c * 4, q[u]);
  }
  const int64_t r0 = (int64_t)tok_b[t] * LP + tok_pos[t] + 1;
  float best = -INFINITY, first = INFINITY;
  for (int p = 0; p < P; ++p) {
    float acc = 0.f;
    const bool ok = ((p_mask >> p) & 1u) && tok_ok[(r0 + p) * tok_ok_ld + tok_ok_col] != 0;   // warp-uniform
    if (ok) {
#pragma unroll
      for (int u = 0; u < MAXV; ++u) {
        const int c = lane + 32 * u;
        if (c < D4) {
          float b[4];
          load4<bf16>(t_hat + ((r0 + p) * D4 + c) * 4, b);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc = fmaf(q[u][k], b[k], acc);
        }
      }
      acc = warp_sum(acc);
      best = fmaxf(best, acc);
      if (p == 0) first = acc;
    }
    if (lane == 0) pos_cos[(int64_t)t * P + p] = ok ? acc : __int_as_float(0x7fc00000);  // NaN: offset not served
  }
  if (lane == 0) {
    // unused rows (no valid offset): reference = tau2, every numerator <= 1, the row is multiplied by row_scale = 0
    mref[t] = best == -INFINITY ? tau2 : fmaxf(best * tau2, tau2 - 100.f);
    thr[t] = first;
  }
}

static int nce_pos_ref_launch(const NcePosRefJobs& jobs, int n_jobs, int64_t ldq, const void* t_hat, int D,
                              const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, const uint8_t* tok_ok,
                              int tok_ok_ld, const float* logit_scale, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP, "nce_pos_ref: pred_len %d not in [1,%d]", P, NCE_MAXP);
  B200_CHECK_ARG(D % 4 == 0 && ldq % 4 == 0 && D <= 2048, "nce_pos_ref: D=%d must be a multiple of 4, <= 2048", D);
  B200_CHECK_ARG(n_jobs >= 1 && n_jobs <= NCE_MAX_JOBS, "nce_pos_ref: %d jobs (1..%d per launch)", n_jobs, NCE_MAX_JOBS);
  if (T == 0) return 0;
  const dim3 grid(ceil_div_i(T, 8), n_jobs);
  if (D <= 512)
    nce_pos_ref_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(jobs, ldq, (const bf16*)t_hat, D / 4, tok_b, tok_pos, T,
                                                                  LP, P, tok_ok, tok_ok_ld, logit_scale);
  else
    nce_pos_ref_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(jobs, ldq, (const bf16*)t_hat, D / 4, tok_b, tok_pos, T,
                                                                   LP, P, tok_ok, tok_ok_ld, logit_scale);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_nce_pos_ref(const void* q_hat, int64_t ldq, const void* t_hat, int D, const int32_t* tok_b,
                                   const int32_t* tok_pos, int T, int LP, int P, uint32_t p_mask, const uint8_t* tok_ok,
                                   int tok_ok_ld, int tok_ok_col, const float* logit_scale, float* pos_cos, float* mref,
                                   float* thr, void* stream) {
  NcePosRefJobs jobs;
  jobs.j[0] = {(const bf16*)q_hat, p_mask, tok_ok_col, pos_cos, mref, thr};
  return nce_pos_ref_launch(jobs, 1, ldq, t_hat, D, tok_b, tok_pos, T, LP, P, tok_ok, tok_ok_ld, logit_scale, stream);
}

extern "C" int b200rec_nce_pos_ref_grouped(const b200rec_nce_pos_ref_job* jobs_in, int n_jobs, int64_t ldq,
                                           const void* t_hat, int D, const int32_t* tok_b, const int32_t* tok_pos, int T,
                                           int LP, int P, const uint8_t* tok_ok, int tok_ok_ld, const float* logit_scale,
                                           void* stream) {
  for (int j0 = 0; j0 < n_jobs; j0 += NCE_MAX_JOBS) {
    const int n = std::min(NCE_MAX_JOBS, n_jobs - j0);
    NcePosRefJobs jobs;
    for (int j = 0; j < n; ++j) {
      const b200rec_nce_pos_ref_job& x = jobs_in[j0 + j];
      jobs.j[j] = {(const bf16*)x.q_hat, x.p_mask, x.tok_ok_col, x.pos_cos, x.mref, x.thr};
    }
    if (nce_pos_ref_launch(jobs, n, ldq, t_hat, D, tok_b, tok_pos, T, LP, P, tok_ok, tok_ok_ld, logit_scale, stream))
      return 1;
  }
  return 0;
}

struct NceCombineJob {
  const float* stats; bf16* E; const uint32_t* same_bits; const uint8_t* row_any; const float* pos_cos;
  const float* mref; const bf16* q_hat; const float* coef; float* loss; float* g0; float* dscale; int32_t* rank0;
  int32_t* nvalid; float* row_scale; bf16* qs;
};
struct NceCombineJobs { NceCombineJob j[NCE_MAX_JOBS]; };

// ld_out: row pitch of the [T, P] loss output (P for a separate tensor; n_jobs * P when the jobs' losses are column blocks
// of one [T, n_jobs * P] tensor, which lets ONE column-sum launch reduce every job's per-offset losses)
__global__ void __launch_bounds__(256) nce_combine_kernel(
    const __grid_constant__ NceCombineJobs jobs, int n_parts, int64_t lde, int n_neg, int64_t ldq, int D4,
    const int32_t* __restrict__ tok_b, const int32_t* __restrict__ tok_pos, int T, int LP, int P, int64_t ld_out,
    const float* __restrict__ logit_scale, int64_t ldqs) {
  pdl_trigger();
  const NceCombineJob& jb = jobs.j[blockIdx.y];
  const float* __restrict__ stats = jb.stats;
  bf16* __restrict__ E = jb.E;
  const uint32_t* __restrict__ same_bits = jb.same_bits;
  const uint8_t* __restrict__ row_any = jb.row_any;
  const float* __restrict__ pos_cos = jb.pos_cos;
  const float* __restrict__ mref_a = jb.mref;
  const bf16* __restrict__ q_hat = jb.q_hat;
  const float* __restrict__ coef = jb.coef;
  float* __restrict__ loss = jb.loss;
  float* __restrict__ g0 = jb.g0;
  float* __restrict__ dscale = jb.dscale;
  int32_t* __restrict__ rank0 = jb.rank0;
  int32_t* __restrict__ nvalid = jb.nvalid;
  float* __restrict__ row_scale = jb.row_scale;
  bf16* __restrict__ qs = jb.qs;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const float tau = __expf(fminf(fmaxf(*logit_scale, 0.f), 4.605170185988092f));
  const float tau2 = tau * LOG2E;
  const float mref = mref_a[t];
  const int64_t r0 = (int64_t)tok_b[t] * LP + tok_pos[t] + 1;
  const int n_words = (n_neg + 31) >> 5;
  // row totals of the stored numerators: fixed lane assignment + fixed shuffle tree -> deterministic
  float S = 0.f, Wr = 0.f, GT = 0.f;
  for (int q = lane; q < n_parts; q += 32) {
    float st4[4];
    load4<float>(stats + ((int64_t)t * n_parts + q) * 4, st4);
    S += st4[0]; Wr += st4[1]; GT += st4[2];
  }
  S = warp_sum(S); Wr = warp_sum(Wr); GT = warp_sum(GT);
  const float pc0 = pos_cos[(int64_t)t * P];            // NaN when offset 0 is not served
  float cp[NCE_MAXP];
  uint32_t masked_set = 0;
  float Csum = 0.f;
#pragma unroll
  for (int p = 0; p < NCE_MAXP; ++p) {
    cp[p] = 0.f;
    if (p >= P) continue;
    const float pc = pos_cos[(int64_t)t * P + p];
    const bool valid = pc == pc;                           // warp-uniform
    float l_ = 0.f, g_ = 0.f, d_ = 0.f;
    int rk = -1, nv = 0;
    if (valid) {
      float Sp = S, Wp = Wr, gtp = GT;
      int cnt = n_neg;
      const bool masked = row_any[r0 + p] != 0;
      if (masked) {
        // subtract exactly the stored numerators of the filtered negatives (rare: a negative that IS this target)
        const uint32_t* bits = same_bits + (r0 + p) * n_words;
        float ds = 0.f, dw = 0.f, dg = 0.f;
        int dc = 0;
        for (int w = lane; w < n_words; w += 32) {
          uint32_t b = bits[w];
          while (b) {
            const int j = (w << 5) + __ffs(b) - 1;
            b &= b - 1;
            if (j < n_neg) {
              const float e = __bfloat162float(E[(int64_t)t * lde + j]);
              dc += 1;
              if (e > 0.f) {
                const float cosj = (__log2f(e) + mref) / tau2;
                ds += e;
                dw = fmaf(e, cosj, dw);
                dg += (p == 0 && cosj > pc0) ? 1.f : 0.f;
              }
            }
          }
        }
        ds = warp_sum(ds); dw = warp_sum(dw); dg = warp_sum(dg);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dc += __shfl_xor_sync(0xffffffffu, dc, o);
        Sp = fmaxf(S - ds, 0.f); Wp = Wr - dw; gtp = GT - dg; cnt = n_neg - dc;
        masked_set |= 1u << p;
      }
      const float zp2 = tau2 * pc;
      const float a = Sp > 0.f ? mref + __log2f(Sp) : -INFINITY;
      const float M = fmaxf(a, zp2);
      const float lse2 = M + __log2f(exp2f(a - M) + exp2f(zp2 - M));
      const float c = coef[p];
      const float sm0 = exp2f(zp2 - lse2);
      const float wn = exp2f(mref - lse2);                 // weight of a unit numerator in this offset's softmax
      l_ = c * (lse2 - zp2) * LN2;
      g_ = c * (sm0 - 1.f);
      // d loss / d logit_scale = sum_k softmax_k z_k - z_0 with natural-log logits z = tau * cos
      d_ = c * (tau * Wp * wn + sm0 * tau * pc - tau * pc);
      rk = (p == 0 || masked) ? (int)(gtp + 0.5f) : -1;
      nv = cnt + 1;
      cp[p] = c * wn;
      Csum += cp[p];
    }
    if (lane == 0) {
      loss[(int64_t)t * ld_out + p] = l_;
      g0[(int64_t)t * P + p] = g_;
      dscale[(int64_t)t * P + p] = d_;
      rank0[(int64_t)t * P + p] = rk;
      nvalid[(int64_t)t * P + p] = nv;
    }
  }
  const float rs = tau * Csum;                             // G[t, j] = rs * E[t, j] for every unfiltered (t, j)
  if (lane == 0) row_scale[t] = rs;
  if (masked_set && Csum > 0.f) {
    // a filtered (offset, negative) pair takes no gradient from that offset: scale the stored numerator by the share of
    // the remaining offsets, so that rs * E'[t, j] is exactly the gradient w.r.t. logit (t, j)
    for (int w = lane; w < n_words; w += 32) {
      uint32_t wb[NCE_MAXP];
      uint32_t any = 0;
#pragma unroll
      for (int p = 0; p < NCE_MAXP; ++p) {
        wb[p] = ((masked_set >> p) & 1u) ? same_bits[(r0 + p) * n_words + w] : 0u;
        any |= wb[p];
      }
      while (any) {
        const int bit = __ffs(any) - 1;
        any &= any - 1;
        const int j = (w << 5) + bit;
        if (j >= n_neg) continue;
        float sc = 0.f;
#pragma unroll
        for (int p = 0; p < NCE_MAXP; ++p) sc += ((wb[p] >> bit) & 1u) ? cp[p] : 0.f;
        const float f = fmaxf(1.f - sc / Csum, 0.f);
        const int64_t at = (int64_t)t * lde + j;
        E[at] = __float2bfloat16_rn(__bfloat162float(E[at]) * f);
      }
    }
  }
  if (qs != nullptr) {
    for (int c = lane; c < D4; c += 32) {
      float v[4];
      load4<bf16>(q_hat + (int64_t)t * ldq + c * 4, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] *= rs;
      store4<bf16>(qs + (int64_t)t * ldqs + c * 4, v);
    }
  }
}

static int nce_combine_launch(const NceCombineJobs& jobs, int n_jobs, int n_parts, int64_t lde, int n_neg, int64_t ldq,
                              int D, const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int P, int64_t ld_out,
                              const float* logit_scale, int64_t ldqs, void* stream) {
  B200_CHECK_ARG(P >= 1 && P <= NCE_MAXP, "nce_combine: pred_len %d not in [1,%d]", P, NCE_MAXP);
  B200_CHECK_ARG(D % 4 == 0 && ldq % 4 == 0 && ldqs % 4 == 0 && n_parts >= 1 && ld_out >= P,
                 "nce_combine: bad D / ld / n_parts");
  B200_CHECK_ARG(n_jobs >= 1 && n_jobs <= NCE_MAX_JOBS, "nce_combine: %d jobs (1..%d per launch)", n_jobs, NCE_MAX_JOBS);
  if (T == 0) return 0;
  nce_combine_kernel<<<dim3(ceil_div_i(T, 8), n_jobs), 256, 0, (cudaStream_t)stream>>>(
      jobs, n_parts, lde, n_neg, ldq, D / 4, tok_b, tok_pos, T, LP, P, ld_out, logit_scale, ldqs);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_nce_combine(const float* stats, int n_parts, void* E, int64_t lde, int n_neg,
                                   const uint32_t* same_bits, const uint8_t* row_any, const float* pos_cos,
                                   const float* mref, const void* q_hat, int64_t ldq, int D, const int32_t* tok_b,
                                   const int32_t* tok_pos, int T, int LP, int P, const float* coef,
                                   const float* logit_scale, float* loss, float* g0, float* dscale, int32_t* rank0,
                                   int32_t* nvalid, float* row_scale, void* qs, int64_t ldqs, void* stream) {
  NceCombineJobs jobs;
  jobs.j[0] = {stats, (bf16*)E, same_bits, row_any, pos_cos, mref, (const bf16*)q_hat, coef, loss, g0, dscale, rank0,
               nvalid, row_scale, (bf16*)qs};
  return nce_combine_launch(jobs, 1, n_parts, lde, n_neg, ldq, D, tok_b, tok_pos, T, LP, P, P, logit_scale, ldqs, stream);
}

extern "C" int b200rec_nce_combine_grouped(const b200rec_nce_combine_job* jobs_in, int n_jobs, int n_parts, int64_t lde,
                                           int n_neg, int64_t ldq, int D, const int32_t* tok_b, const int32_t* tok_pos,
                                           int T, int LP, int P, int64_t ld_out, const float* logit_scale, int64_t ldqs,
                                           void* stream) {
  for (int j0 = 0; j0 < n_jobs; j0 += NCE_MAX_JOBS) {
    const int n = std::min(NCE_MAX_JOBS, n_jobs - j0);
    NceCombineJobs jobs;
    for (int j = 0; j < n; ++j) {
      const b200rec_nce_combine_job& x = jobs_in[j0 + j];
      jobs.j[j] = {x.stats, (bf16*)x.E, x.same_bits, x.row_any, x.pos_cos, x.mref, (const bf16*)x.q_hat, x.coef, x.loss,
                   x.g0, x.dscale, x.rank0, x.nvalid, x.row_scale, (bf16*)x.qs};
    }
    if (nce_combine_launch(jobs, n, n_parts, lde, n_neg, ldq, D, tok_b, tok_pos, T, LP, P, ld_out, logit_scale, ldqs,
                           stream))
      return 1;
  }
  return 0;
}

__global__ void __launch_bounds__(256) tail_norm_kernel(const bf16* __restrict__ x, int64_t n, int D4, int k04,
                                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  float ss = 0.f;
  for (int c = k04 + lane; c < D4; c += 32) {
    float v[4];
    load4<bf16>(x + (r * D4 + c) * 4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) ss = fmaf(v[k], v[k], ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) out[r] = sqrtf(ss) * (1.f + 1e-6f);     // never below the true norm of the rounded row
}

// Augmented prefix operand of the pruned filter: out[r, 0:k0] = x[r, 0:k0], out[r, k0] = |x[r, k0:]| rounded UP to bf16,
// out[r, k0+1 : k0+16] = 0.  The K = k0 + 16 GEMM of two such operands is  <prefixes> + |tail_a| |tail_b|  — the
// Cauchy-Schwarz upper bound of the full cosine — straight out of the tensor cores: the plain GT_BITS epilogue applies
// (a per-column factor in the epilogue cost more than the whole K = 64 GEMM).
__global__ void __launch_bounds__(256) prefix_aug_kernel(const bf16* __restrict__ x, int64_t n, int D4, int k04,
                                                         bf16* __restrict__ out) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  float ss = 0.f;
  for (int c = k04 + lane; c < D4; c += 32) {
    float v[4];
    load4<bf16>(x + (r * D4 + c) * 4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) ss = fmaf(v[k], v[k], ss);
  }
  ss = warp_sum(ss);
  const int ko4 = k04 + 4;                                  // output row: k0 + 16 elements
  for (int c = lane; c < ko4; c += 32) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < k04) load4<bf16>(x + (r * D4 + c) * 4, v);
    store4<bf16>(out + (r * ko4 + c) * 4, v);
  }
  if (lane == 0) {
    // round up: the smallest bf16 >= the fp32 norm (so the product of two rounded norms bounds the true product)
    const float nrm = sqrtf(ss) * (1.f + 1e-6f);
    uint32_t u = __float_as_uint(nrm);
    if (u & 0xffffu) u = (u | 0xffffu) + 1u;                // next multiple of 2^16 (positive floats: monotone)
    out[r * ko4 * 4 + (int64_t)k04 * 4] = __float2bfloat16_rn(__uint_as_float(u));
  }
}

extern "C" int b200rec_prefix_aug(const void* x_hat, int64_t n, int D, int k0, void* out, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && k0 % 16 == 0 && k0 >= 16 && k0 < D, "prefix_aug: bad D / k0");
  if (n == 0) return 0;
  prefix_aug_kernel<<<ceil_div_i(n, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)x_hat, n, D / 4, k0 / 4, (bf16*)out);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_tail_norm(const void* x_hat, int64_t n, int D, int k0, float* out, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && k0 % 4 == 0 && k0 >= 0 && k0 <= D, "tail_norm: bad D / k0");
  if (n == 0) return 0;
  tail_norm_kernel<<<ceil_div_i(n, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)x_hat, n, D / 4, k0 / 4, out);
  B200_LAUNCH_OK();
  return 0;
}

template <int MAXV>
__global__ void __launch_bounds__(256) gt_bits_verify_kernel(uint32_t* __restrict__ bits, int64_t M, int n_words, int N,
                                                             const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                             int D4, float thres, uint8_t* __restrict__ row_any,
                                                             int64_t bits_stride, int64_t b_stride, int64_t any_stride) {
  pdl_trigger();
  // grid.y = negative set: the sets share the row operand a, everything else is a strided block per set
  bits += (int64_t)blockIdx.y * bits_stride;
  b += (int64_t)blockIdx.y * b_stride;
  row_any += (int64_t)blockIdx.y * any_stride;
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  // cheap exit: most rows carry no candidate at all
  uint32_t anyw = 0;
  for (int w = lane; w < n_words; w += 32) anyw |= bits[m * n_words + w];
  if (__ballot_sync(0xffffffffu, anyw != 0) == 0) return;
  float av[MAXV][4];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int c = lane + 32 * u;
    if (c < D4) load4<bf16>(a + (m * D4 + c) * 4, av[u]);
  }
  bool keep_any = false;
  for (int w0 = 0; w0 < n_words; w0 += 32) {
    const int w = w0 + lane;
    uint32_t word = w < n_words ? bits[m * n_words + w] : 0u;
    uint32_t pending = __ballot_sync(0xffffffffu, word != 0);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      uint32_t cur = __shfl_sync(0xffffffffu, word, src);
      uint32_t kept = cur;
      while (cur) {
        const int bit = __ffs(cur) - 1;
        cur &= cur - 1;
        const int64_t j = (int64_t)(w0 + src) * 32 + bit;
        float acc = 0.f;
        if (j < N) {
#pragma unroll
          for (int u = 0; u < MAXV; ++u) {
            const int c = lane + 32 * u;
            if (c < D4) {
              float bv[4];
              load4<bf16>(b + (j * D4 + c) * 4, bv);
#pragma unroll
              for (int k = 0; k < 4; ++k) acc = fmaf(av[u][k], bv[k], acc);
            }
          }
        }
        acc = warp_sum(acc);
        if (!(j < N && acc > thres)) kept &= ~(1u << bit);
      }
      if (lane == src) word = kept;
    }
    if (w < n_words) bits[m * n_words + w] = word;
    keep_any |= __ballot_sync(0xffffffffu, word != 0) != 0;
  }
  if (lane == 0 && keep_any) row_any[m] = 1;
}

static int gt_bits_verify_launch(uint32_t* bits, int64_t M, int n_words, int N, const void* a_hat, const void* b_hat,
                                 int D, float thres, uint8_t* row_any, int n_sets, int64_t bits_stride,
                                 int64_t b_stride, int64_t any_stride, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && n_words * 32 >= N, "gt_bits_verify: bad D / n_words");
  if (M == 0 || n_sets == 0) return 0;
  const dim3 grid(ceil_div_i(M, 8), n_sets);
  if (D <= 512)
    gt_bits_verify_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(bits, M, n_words, N, (const bf16*)a_hat,
                                                                     (const bf16*)b_hat, D / 4, thres, row_any,
                                                                     bits_stride, b_stride, any_stride);
  else
    gt_bits_verify_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(bits, M, n_words, N, (const bf16*)a_hat,
                                                                      (const bf16*)b_hat, D / 4, thres, row_any,
                                                                      bits_stride, b_stride, any_stride);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_gt_bits_verify(uint32_t* bits, int64_t M, int n_words, int N, const void* a_hat,
                                      const void* b_hat, int D, float thres, uint8_t* row_any, void* stream) {
  return gt_bits_verify_launch(bits, M, n_words, N, a_hat, b_hat, D, thres, row_any, 1, 0, 0, 0, stream);
}

extern "C" int b200rec_gt_bits_verify_sets(uint32_t* bits, int64_t M, int n_words, int N, const void* a_hat,
                                           const void* b_hat, int D, float thres, uint8_t* row_any, int n_sets,
                                           int64_t bits_stride, int64_t b_stride, int64_t any_stride, void* stream) {
  return gt_bits_verify_launch(bits, M, n_words, N, a_hat, b_hat, D, thres, row_any, n_sets, bits_stride, b_stride,
                               any_stride, stream);
}
