// Fused AdamW (torch.optim.AdamW maths, trainer.py:296-299), one pass over p/m/v.  The embedding
// variant consumes the compact (unique id, row) gradient of the sorted-segment scatter-add and is
// dense-equivalent: rows without a gradient still decay and move with their momentum.
#include "common.cuh"

struct AdamCoef {
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt, gs;
  const float* dev;  // optional device override {lr, bc1, bc2_sqrt, step}: lets a captured CUDA graph advance the step
};

__device__ __forceinline__ AdamCoef adam_resolve(AdamCoef c) {
  if (c.dev) {
    c.lr = c.dev[0];
    c.bc1 = c.dev[1];
    c.bc2_sqrt = c.dev[2];
  }
  return c;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamCoef& c) {
  g *= c.gs;
  p *= (1.f - c.lr * c.wd);
  m = c.b1 * m + (1.f - c.b1) * g;
  v = c.b2 * v + (1.f - c.b2) * g * g;
  float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
  p -= (c.lr / c.bc1) * (m / denom);
}

// Table-row variant: the per-step scalars are folded once (decay, step size, 1/bias-correction) and sqrt / divide
// use the MUFU approximations (rel. error ~2^-22): ~9 instructions per element instead of ~25, which is what the
// lazy catch-up loop is bound by.  The dense-equivalent rows pass and the lazy pass share this function, so they stay
// bit-identical to each other; against torch.optim.AdamW the table differs by ~1e-7 relative per step.
struct RowStep {
  float decay, step_size, inv_bc2;
};
__device__ __forceinline__ RowStep row_step(float lr, float bc1, float bc2_sqrt, float wd) {
  RowStep r;
  r.decay = 1.f - lr * wd;
  r.step_size = lr / bc1;
  r.inv_bc2 = 1.f / bc2_sqrt;
  return r;
}
__device__ __forceinline__ void adam_update_row(float& p, float& m, float& v, float g, const RowStep& r, float b1,
                                                float b2, float eps) {
  p *= r.decay;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  float sq, rc;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));
  const float denom = fmaf(sq, r.inv_bc2, eps);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(denom));
  p = fmaf(-r.step_size * m, rc, p);
}

static AdamCoef make_coef(float lr, float b1, float b2, float eps, float wd, int step, float gs) {
  AdamCoef c;
  c.lr = lr; c.b1 = b1; c.b2 = b2; c.eps = eps; c.wd = wd; c.gs = gs;
  c.bc1 = 1.f - powf(b1, (float)step);
  c.bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  c.dev = nullptr;
  return c;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m,
                                                    float* __restrict__ v, const float* __restrict__ g, int64_t n,
                                                    AdamCoef c_in) {
  const AdamCoef c = adam_resolve(c_in);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float pp[4], mm[4], vv[4], gg[4];
    load4<float>(p + i * 4, pp);
    load4<float>(m + i * 4, mm);
    load4<float>(v + i * 4, vv);
    load4<float>(g + i * 4, gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) adam_update(pp[k], mm[k], vv[k], gg[k], c);
    store4<float>(p + i * 4, pp);
    store4<float>(m + i * 4, mm);
    store4<float>(v + i * 4, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = n4 * 4 + threadIdx.x;
    adam_update(p[i], m[i], v[i], g[i], c);
  }
}

__global__ void __launch_bounds__(256) adamw_scalar_kernel(float* __restrict__ p, float* __restrict__ m,
                                                           float* __restrict__ v, const float* __restrict__ g,
                                                           int64_t n, AdamCoef c_in) {
  const AdamCoef c = adam_resolve(c_in);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_update(p[i], m[i], v[i], g[i], c);
}

int b200rec_adamw(float* p, float* m, float* v, const float* g, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, const float* coef_dev, void* stream) {
  if (n == 0) return 0;
  AdamCoef c = make_coef(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  c.dev = coef_dev;
  if ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) != 0) {  // unaligned view: scalar path
    adamw_scalar_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(p, m, v, g,
                                                                                                           n, c);
    B200_LAUNCH_OK();
    return 0;
  }
  int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 148 * 16);
  adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, m, v, g, n, c);
  B200_LAUNCH_OK();
  return 0;
}

__global__ void row_slot_kernel(const int64_t* __restrict__ uniq_ids, const int32_t* __restrict__ n_uniq,
                                int32_t* __restrict__ row_slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *n_uniq) row_slot[uniq_ids[i]] = i;
}

__global__ void __launch_bounds__(256) adamw_rows_kernel(float* __restrict__ p, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n_vec, int D4,
                                                         const int32_t* __restrict__ row_slot,
                                                         const float* __restrict__ uniq_rows, AdamCoef c_in) {
  const AdamCoef c = adam_resolve(c_in);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    int64_t r = i / D4;
    int col = (int)(i - r * D4);
    int slot = __ldg(row_slot + r);
    float pp[4], mm[4], vv[4], gg[4] = {0.f, 0.f, 0.f, 0.f};
    load4<float>(p + i * 4, pp);
    load4<float>(m + i * 4, mm);
    load4<float>(v + i * 4, vv);
    if (slot >= 0) load4<float>(uniq_rows + ((int64_t)slot * D4 + col) * 4, gg);
    const RowStep rs = row_step(c.lr, c.bc1, c.bc2_sqrt, c.wd);
#pragma unroll
    for (int k = 0; k < 4; ++k) adam_update_row(pp[k], mm[k], vv[k], gg[k] * c.gs, rs, c.b1, c.b2, c.eps);
    store4<float>(p + i * 4, pp);
    store4<float>(m + i * 4, mm);
    store4<float>(v + i * 4, vv);
  }
}

int b200rec_adamw_rows(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* uniq_ids,
                       const float* uniq_rows, const int32_t* n_uniq, int32_t* row_slot_ws, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float grad_scale, const float* coef_dev,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(D % 4 == 0, "adamw_rows: D=%d must be a multiple of 4", D);
  if (n_rows == 0) return 0;
  AdamCoef c = make_coef(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  c.dev = coef_dev;
  B200_CUDA_OK(cudaMemsetAsync(row_slot_ws, 0xff, (size_t)n_rows * 4, st));
  // n_uniq lives on the device: launch for the worst case (every row touched)
  row_slot_kernel<<<ceil_div_i(n_rows, 256), 256, 0, st>>>(uniq_ids, n_uniq, row_slot_ws);
  int64_t n_vec = n_rows * (D / 4);
  int blocks = (int)std::min<int64_t>((n_vec + 255) / 256, 148 * 16);
  adamw_rows_kernel<<<blocks, 256, 0, st>>>(p, m, v, n_vec, D / 4, row_slot_ws, uniq_rows, c);
  B200_LAUNCH_OK();
  return 0;
}

// ---- device-side step counter: coef = {lr, bc1, bc2_sqrt, step}; a captured graph replays this tick ----------
__global__ void adamw_tick_kernel(float* coef, float b1, float b2) {
  float step = coef[3] + 1.f;
  coef[3] = step;
  coef[1] = 1.f - powf(b1, step);
  coef[2] = sqrtf(1.f - powf(b2, step));
}

extern "C" int b200rec_adamw_tick(float* coef_dev, float beta1, float beta2, void* stream) {
  adamw_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(coef_dev, beta1, beta2);
  B200_LAUNCH_OK();
  return 0;
}

// ---- multi-tensor AdamW: one launch for every dense parameter ---------------------------------------------
// table[t] = {p, m, v, g, n}; blocks[b] = {tensor index, first element of the 4096-element chunk}
struct AdamTensor {
  float *p, *m, *v;
  const float* g;
  bf16* shadow;   // optional compute-dtype copy of p, refreshed in the same pass (NULL: none)
  int64_t n;
};
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamTensor* __restrict__ table,
                                                          const int64_t* __restrict__ blocks, AdamCoef c_in) {
  const AdamCoef c = adam_resolve(c_in);
  const AdamTensor t = table[blocks[2 * blockIdx.x]];
  const int64_t beg = blocks[2 * blockIdx.x + 1];
  const int64_t end = min(t.n, beg + 4096);
  const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.m | (uintptr_t)t.v | (uintptr_t)t.g) & 15) == 0) &&
                   (((uintptr_t)t.shadow & 7) == 0);
  if (vec) {
    // 4 x 16-byte vectors per thread and array, every load issued before the first use
    const int64_t n4 = (end - beg) >> 2;
    float pp[4][4], mm[4][4], vv[4][4], gg[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = threadIdx.x + k * 256;
      if (i < n4) {
        load4<float>(t.p + beg + i * 4, pp[k]);
        load4<float>(t.m + beg + i * 4, mm[k]);
        load4<float>(t.v + beg + i * 4, vv[k]);
        load4<float>(t.g + beg + i * 4, gg[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = threadIdx.x + k * 256;
      if (i < n4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) adam_update(pp[k][e], mm[k][e], vv[k][e], gg[k][e], c);
        store4<float>(t.p + beg + i * 4, pp[k]);
        store4<float>(t.m + beg + i * 4, mm[k]);
        store4<float>(t.v + beg + i * 4, vv[k]);
        if (t.shadow) store4<bf16>(t.shadow + beg + i * 4, pp[k]);
      }
    }
    for (int64_t i = beg + n4 * 4 + threadIdx.x; i < end; i += 256) {
      adam_update(t.p[i], t.m[i], t.v[i], t.g[i], c);
      if (t.shadow) t.shadow[i] = __float2bfloat16_rn(t.p[i]);
    }
    return;
  }
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) {
    adam_update(t.p[i], t.m[i], t.v[i], t.g[i], c);
    if (t.shadow) t.shadow[i] = __float2bfloat16_rn(t.p[i]);
  }
}

extern "C" int b200rec_adamw_multi(const void* table_dev, const int64_t* blocks_dev, int n_blocks, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int step,
                                   float grad_scale, const float* coef_dev, void* stream) {
  if (n_blocks == 0) return 0;
  AdamCoef c = make_coef(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  c.dev = coef_dev;
  adamw_multi_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>((const AdamTensor*)table_dev, blocks_dev, c);
  B200_LAUNCH_OK();
  return 0;
}


// ---- lazy dense-equivalent AdamW on the item table ----------------------------------------------------------
// The dense-equivalent update touches all N rows every step (11 GB at N = 450 k, D = 1024) although a step reads
// ~80 k of them.  A row whose gradient is zero at step j only needs  m *= b1, v *= b2, p = f_j(p, m, v)  with the
// step's scalars: nothing but its own (p, m, v).  So the updates of rows nobody looks at are DEFERRED: last[row]
// is the last step applied to the row, hist[j] = {lr_j, bc1_j, bc2_sqrt_j} keeps every step's scalars, and a row is
// brought up to date (the same per-step formula, in the same order, with g = 0: bit-identical to the dense pass)
// when something is about to read it: the lookups of the next step, evaluation, a checkpoint.  Exact torch.optim.AdamW
// semantics, ~6x less HBM traffic per step.
__global__ void adamw_tick_hist_kernel(float* coef, float4* hist, int cap, float b1, float b2, float wd) {
  float step = coef[3] + 1.f;
  coef[3] = step;
  coef[1] = 1.f - powf(b1, step);
  coef[2] = sqrtf(1.f - powf(b2, step));
  int j = (int)step;
  const RowStep r = row_step(coef[0], coef[1], coef[2], wd);
  hist[j & (cap - 1)] = make_float4(r.decay, r.step_size, r.inv_bc2, coef[0]);   // ring: cap is a power of two
}

extern "C" int b200rec_adamw_tick_hist(float* coef_dev, void* hist, int cap, float beta1, float beta2,
                                       float weight_decay, void* stream) {
  adamw_tick_hist_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(coef_dev, (float4*)hist, cap, beta1, beta2, weight_decay);
  B200_LAUNCH_OK();
  return 0;
}

template <int MAXV>
__device__ __forceinline__ void lazy_catch_up(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                              int64_t row, int D4, int from, int to, const float4* __restrict__ hist,
                                              int hmask, AdamCoef c, const float* __restrict__ grow, bool with_grad) {
  // steps from+1 .. to with zero gradient; if with_grad, step to+1 follows with gradient row `grow`
  const int lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < D4; c0 += 32 * MAXV) {
    float pp[MAXV][4], mm[MAXV][4], vv[MAXV][4];
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      const int col = c0 + lane + 32 * u;
      if (col < D4) {
        load4<float>(p + (row * D4 + col) * 4, pp[u]);
        load4<float>(m + (row * D4 + col) * 4, mm[u]);
        load4<float>(v + (row * D4 + col) * 4, vv[u]);
      }
    }
    for (int j = from + 1; j <= to; ++j) {
      const float4 h = __ldg(hist + (j & hmask));
      RowStep rs;
      rs.decay = h.x; rs.step_size = h.y; rs.inv_bc2 = h.z;
#pragma unroll
      for (int u = 0; u < MAXV; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) adam_update_row(pp[u][k], mm[u][k], vv[u][k], 0.f, rs, c.b1, c.b2, c.eps);
    }
    if (with_grad) {
      const float4 h = __ldg(hist + ((to + 1) & hmask));
      RowStep rs;
      rs.decay = h.x; rs.step_size = h.y; rs.inv_bc2 = h.z;
#pragma unroll
      for (int u = 0; u < MAXV; ++u) {
        const int col = c0 + lane + 32 * u;
        if (col < D4) {
          float gg[4];
          load4<float>(grow + col * 4, gg);
#pragma unroll
          for (int k = 0; k < 4; ++k) adam_update_row(pp[u][k], mm[u][k], vv[u][k], gg[k] * c.gs, rs, c.b1, c.b2, c.eps);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      const int col = c0 + lane + 32 * u;
      if (col < D4) {
        store4<float>(p + (row * D4 + col) * 4, pp[u]);
        store4<float>(m + (row * D4 + col) * 4, mm[u]);
        store4<float>(v + (row * D4 + col) * 4, vv[u]);
      }
    }
  }
}

// one warp per requested id (duplicates allowed: the first claimant does the work, kernel boundaries order the reads)
__global__ void __launch_bounds__(256) adamw_rows_catchup_kernel(float* __restrict__ p, float* __restrict__ m,
                                                                 float* __restrict__ v, int64_t N, int D4,
                                                                 const int64_t* __restrict__ ids, int64_t n_ids,
                                                                 int32_t* __restrict__ last,
                                                                 const float4* __restrict__ hist, int hmask, AdamCoef c,
                                                                 int64_t id_stride, int64_t id_offset) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int S = (int)c.dev[3];                                  // completed optimizer steps
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_ids; i += warps) {
    int64_t row = ids ? ids[i] : i;
    if (ids && id_stride > 1) {                                 // global ids of a row-sharded table: keep my rows
      if (row < 0 || row % id_stride != id_offset) continue;
      row /= id_stride;
    }
    if (row < 0 || row >= N) continue;
    int old = 0;
    if (lane == 0) old = atomicExch(last + row, S);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old >= S) continue;
    lazy_catch_up<4>(p, m, v, row, D4, old, S, hist, hmask, c, nullptr, false);
  }
}

__global__ void __launch_bounds__(256) adamw_rows_lazy_kernel(float* __restrict__ p, float* __restrict__ m,
                                                              float* __restrict__ v, int D4,
                                                              const int64_t* __restrict__ uniq_ids,
                                                              const float* __restrict__ uniq_rows,
                                                              const int32_t* __restrict__ n_uniq,
                                                              int32_t* __restrict__ last,
                                                              const float4* __restrict__ hist, int hmask, AdamCoef c) {
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int S1 = (int)c.dev[3];                                 // the step being applied (tick already ran)
  const int nu = *n_uniq;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nu; i += warps) {
    const int64_t row = uniq_ids[i];                             // unique: no claim needed
    const int old = last[row];
    lazy_catch_up<4>(p, m, v, row, D4, min(old, S1 - 1), S1 - 1, hist, hmask, c, uniq_rows + i * D4 * 4, true);
    if ((threadIdx.x & 31) == 0) last[row] = S1;
  }
}

extern "C" int b200rec_adamw_rows_catchup(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* ids,
                                          int64_t n_ids, int32_t* last, const void* hist, int hist_cap,
                                          const float* coef_dev, float beta1, float beta2, float eps, float weight_decay,
                                          int64_t id_stride, int64_t id_offset, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && coef_dev != nullptr && hist != nullptr && last != nullptr && hist_cap > 0 &&
                     (hist_cap & (hist_cap - 1)) == 0, "adamw_rows_catchup: bad args (hist_cap must be a power of two)");
  if (n_ids == 0) return 0;
  AdamCoef c = make_coef(0.f, beta1, beta2, eps, weight_decay, 1, 1.f);
  c.dev = coef_dev;
  int blocks = (int)std::min<int64_t>((n_ids + 7) / 8, 148 * 8);
  adamw_rows_catchup_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, m, v, n_rows, D / 4, ids, n_ids, last,
                                                                      (const float4*)hist, hist_cap - 1, c,
                                                                      id_stride < 1 ? 1 : id_stride, id_offset);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_adamw_rows_lazy(float* p, float* m, float* v, int64_t n_rows, int D, const int64_t* uniq_ids,
                                       const float* uniq_rows, const int32_t* n_uniq, int64_t max_rows, int32_t* last,
                                       const void* hist, int hist_cap, const float* coef_dev, float beta1, float beta2,
                                       float eps, float weight_decay, float grad_scale, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && coef_dev != nullptr && hist != nullptr && last != nullptr && hist_cap > 0 &&
                     (hist_cap & (hist_cap - 1)) == 0, "adamw_rows_lazy: bad args (hist_cap must be a power of two)");
  if (max_rows == 0 || n_rows == 0) return 0;
  AdamCoef c = make_coef(0.f, beta1, beta2, eps, weight_decay, 1, grad_scale);
  c.dev = coef_dev;
  int blocks = (int)std::min<int64_t>((max_rows + 7) / 8, 148 * 8);
  adamw_rows_lazy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, m, v, D / 4, uniq_ids, uniq_rows, n_uniq, last,
                                                                   (const float4*)hist, hist_cap - 1, c);
  B200_LAUNCH_OK();
  return 0;
}
