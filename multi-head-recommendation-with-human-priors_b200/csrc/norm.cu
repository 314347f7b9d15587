// Row-wise normalisation kernels of the HSTU block: one warp per token row, the row
// is held in registers between the statistics pass and the write (single HBM read of each input).
#include "common.cuh"

// one warp per row; small blocks so that ~5 blocks (20 rows) are resident per SM and the last wave is short
#define ROWS_PER_BLOCK 4
#define ROW_THREADS (32 * ROWS_PER_BLOCK)

template <typename TY, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS) layernorm_fwd_kernel(const float* __restrict__ x, int T, int D4, float eps,
                                                            TY* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= T) return;
  float v[MAXV][4];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      load4<float>(x + ((int64_t)r * D4 + c) * 4, v[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) s += v[u][k];
    }
  }
  float D = (float)(D4 * 4);
  float mu = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = v[u][k] - mu;
        ss += d * d;
      }
    }
  }
  float rs = rsqrtf(warp_sum(ss) / D + eps);
  if (lane == 0) {
    mean[r] = mu;
    rstd[r] = rs;
  }
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] = (v[u][k] - mu) * rs;
      store4<TY>(y + ((int64_t)r * D4 + c) * 4, v[u]);
    }
  }
}

int b200rec_layernorm_fwd(const float* x, int T, int D, float eps, void* y, int y_dtype, float* mean, float* rstd,
                          void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048, "layernorm_fwd: D=%d must be a multiple of 4 and <= 2048", D);
  if (T == 0) return 0;
  int blocks = ceil_div_i(T, ROWS_PER_BLOCK);
  DISPATCH_ACT(y_dtype, TY, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(layernorm_fwd_kernel<TY, 4>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, x, T, D / 4, eps, (TY*)y, mean, rstd));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(layernorm_fwd_kernel<TY, 8>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, x, T, D / 4, eps, (TY*)y, mean, rstd));
    else
      B200_CUDA_OK(launch_pdl(layernorm_fwd_kernel<TY, 16>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, x, T, D / 4, eps, (TY*)y, mean, rstd));
  });
  B200_LAUNCH_OK();
  return 0;
}

template <typename TG, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS) layernorm_bwd_kernel(const TG* __restrict__ dy, int64_t ldy,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, int T, int D4,
                                                            const float* __restrict__ resid, float* __restrict__ dx,
                                                            TG* __restrict__ dx_act) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= T) return;
  float mu = mean[r], rs = rstd[r];
  float xh[MAXV][4], g[MAXV][4];
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      load4<float>(x + ((int64_t)r * D4 + c) * 4, xh[u]);
      load4<TG>(dy + (int64_t)r * ldy + c * 4, g[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xh[u][k] = (xh[u][k] - mu) * rs;
        sg += g[u][k];
        sgx += g[u][k] * xh[u][k];
      }
    }
  }
  float D = (float)(D4 * 4);
  sg = warp_sum(sg) / D;
  sgx = warp_sum(sgx) / D;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      float o[4];
      if (resid) load4<float>(resid + ((int64_t)r * D4 + c) * 4, o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = rs * (g[u][k] - sg - xh[u][k] * sgx);
        o[k] = resid ? o[k] + d : d;
      }
      store4<float>(dx + ((int64_t)r * D4 + c) * 4, o);
      if (dx_act) store4<TG>(dx_act + ((int64_t)r * D4 + c) * 4, o);
    }
  }
}

int b200rec_layernorm_bwd(const void* dy, int dy_dtype, int ldy, const float* x, const float* mean, const float* rstd,
                          int T, int D, const float* residual_grad, float* dx, void* dx_act, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && ldy % 4 == 0, "layernorm_bwd: bad D=%d ldy=%d", D, ldy);
  if (T == 0) return 0;
  int blocks = ceil_div_i(T, ROWS_PER_BLOCK);
  DISPATCH_ACT(dy_dtype, TG, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(layernorm_bwd_kernel<TG, 4>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TG*)dy, ldy, x, mean, rstd, T,
                                                                            D / 4, residual_grad, dx, (TG*)dx_act));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(layernorm_bwd_kernel<TG, 8>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TG*)dy, ldy, x, mean, rstd, T,
                                                                            D / 4, residual_grad, dx, (TG*)dx_act));
    else
      B200_CUDA_OK(launch_pdl(layernorm_bwd_kernel<TG, 16>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TG*)dy, ldy, x, mean, rstd, T,
                                                                            D / 4, residual_grad, dx, (TG*)dx_act));
  });
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------- gate: oin = u * LN(a)
template <typename TA, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS) gate_ln_fwd_kernel(const TA* __restrict__ u, int64_t ldu,
                                                          const float* __restrict__ a, int T, int D4, float eps,
                                                          TA* __restrict__ oin, float* __restrict__ mean,
                                                          float* __restrict__ rstd, DropCfg drop) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= T) return;
  float v[MAXV][4];
  float s = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < MAXV; ++k4) {
    int c = lane + k4 * 32;
    if (c < D4) {
      load4<float>(a + ((int64_t)r * D4 + c) * 4, v[k4]);
#pragma unroll
      for (int k = 0; k < 4; ++k) s += v[k4][k];
    }
  }
  float D = (float)(D4 * 4);
  float mu = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < MAXV; ++k4) {
    int c = lane + k4 * 32;
    if (c < D4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = v[k4][k] - mu;
        ss += d * d;
      }
    }
  }
  float rs = rsqrtf(warp_sum(ss) / D + eps);
  if (lane == 0) {
    mean[r] = mu;
    rstd[r] = rs;
  }
#pragma unroll
  for (int k4 = 0; k4 < MAXV; ++k4) {
    int c = lane + k4 * 32;
    if (c < D4) {
      float uu[4];
      load4<TA>(u + (int64_t)r * ldu + c * 4, uu);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k4][k] = (v[k4][k] - mu) * rs * uu[k];
      dropout4(drop, (uint64_t)r * D4 + c, v[k4]);   // F.dropout on the O-proj input (hstu.py:281-285)
      store4<TA>(oin + ((int64_t)r * D4 + c) * 4, v[k4]);
    }
  }
}

int b200rec_gate_ln_fwd(const void* u, int ldu, const float* a, int T, int D, float eps, void* oin, int act_dtype,
                        float* mean, float* rstd, float dropout_p, uint32_t seed, uint32_t layer,
                        const int64_t* rng_step_dev, void* stream) {
  DropCfg drop = {dropout_p, seed, layer, rng_step_dev};
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && ldu % 4 == 0, "gate_ln_fwd: bad D=%d ldu=%d", D, ldu);
  if (T == 0) return 0;
  int blocks = ceil_div_i(T, ROWS_PER_BLOCK);
  DISPATCH_ACT(act_dtype, TA, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(gate_ln_fwd_kernel<TA, 4>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TA*)u, ldu, a, T, D / 4, eps,
                                                                          (TA*)oin, mean, rstd, drop));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(gate_ln_fwd_kernel<TA, 8>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TA*)u, ldu, a, T, D / 4, eps,
                                                                          (TA*)oin, mean, rstd, drop));
    else
      B200_CUDA_OK(launch_pdl(gate_ln_fwd_kernel<TA, 16>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, (const TA*)u, ldu, a, T, D / 4, eps,
                                                                          (TA*)oin, mean, rstd, drop));
  });
  B200_LAUNCH_OK();
  return 0;
}

template <typename TA, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS) gate_ln_bwd_kernel(const TA* __restrict__ d_oin, const TA* __restrict__ u,
                                                          const TA* __restrict__ pre_u, int64_t ldu,
                                                          const float* __restrict__ a, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, int T, int D4,
                                                          TA* __restrict__ d_pre_u, TA* __restrict__ da,
                                                          DropCfg drop) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= T) return;
  float mu = mean[r], rs = rstd[r];
  float lna[MAXV][4], g[MAXV][4];  // g becomes d_lna
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < MAXV; ++k4) {
    int c = lane + k4 * 32;
    if (c < D4) {
      float go[4], uu[4], pu[4], dpu[4];
      load4<float>(a + ((int64_t)r * D4 + c) * 4, lna[k4]);
      load4<TA>(d_oin + ((int64_t)r * D4 + c) * 4, go);
      dropout4(drop, (uint64_t)r * D4 + c, go);       // same keep-mask as the forward
      load4<TA>(u + (int64_t)r * ldu + c * 4, uu);
      load4<TA>(pre_u + (int64_t)r * ldu + c * 4, pu);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lna[k4][k] = (lna[k4][k] - mu) * rs;
        dpu[k] = go[k] * lna[k4][k] * silu_grad_f(pu[k]);
        g[k4][k] = go[k] * uu[k];
        sg += g[k4][k];
        sgx += g[k4][k] * lna[k4][k];
      }
      store4<TA>(d_pre_u + (int64_t)r * ldu + c * 4, dpu);
    }
  }
  float D = (float)(D4 * 4);
  sg = warp_sum(sg) / D;
  sgx = warp_sum(sgx) / D;
#pragma unroll
  for (int k4 = 0; k4 < MAXV; ++k4) {
    int c = lane + k4 * 32;
    if (c < D4) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = rs * (g[k4][k] - sg - lna[k4][k] * sgx);
      store4<TA>(da + ((int64_t)r * D4 + c) * 4, o);
    }
  }
}

int b200rec_gate_ln_bwd(const void* d_oin, const void* u, const void* pre_u, int ldu, const float* a,
                        const float* mean, const float* rstd, int T, int D, void* d_pre_u, void* da, int act_dtype,
                        float dropout_p, uint32_t seed, uint32_t layer, const int64_t* rng_step_dev, void* stream) {
  DropCfg drop = {dropout_p, seed, layer, rng_step_dev};
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048 && ldu % 4 == 0, "gate_ln_bwd: bad D=%d ldu=%d", D, ldu);
  if (T == 0) return 0;
  int blocks = ceil_div_i(T, ROWS_PER_BLOCK);
  DISPATCH_ACT(act_dtype, TA, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(gate_ln_bwd_kernel<TA, 4>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, 
          (const TA*)d_oin, (const TA*)u, (const TA*)pre_u, ldu, a, mean, rstd, T, D / 4, (TA*)d_pre_u, (TA*)da, drop));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(gate_ln_bwd_kernel<TA, 8>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, 
          (const TA*)d_oin, (const TA*)u, (const TA*)pre_u, ldu, a, mean, rstd, T, D / 4, (TA*)d_pre_u, (TA*)da, drop));
    else
      B200_CUDA_OK(launch_pdl(gate_ln_bwd_kernel<TA, 16>, dim3(blocks), dim3(ROW_THREADS), 0, (cudaStream_t)stream, 
          (const TA*)d_oin, (const TA*)u, (const TA*)pre_u, ldu, a, mean, rstd, T, D / 4, (TA*)d_pre_u, (TA*)da, drop));
  });
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------- cast / column sums / reductions
template <typename TY>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ x, int64_t n4, TY* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float v[4];
    load4<float>(x + i * 4, v);
    store4<TY>(y + i * 4, v);
  }
}
template <typename TY>
__global__ void cast_tail_kernel(const float* __restrict__ x, int64_t beg, int64_t n, TY* __restrict__ y) {
  int64_t i = beg + threadIdx.x;
  if (i < n) y[i] = from_f32<TY>(x[i]);
}

int b200rec_cast(const float* x, int64_t n, void* y, int y_dtype, void* stream) {
  if (n == 0) return 0;
  int64_t n4 = n / 4;
  DISPATCH_ACT(y_dtype, TY, {
    if (n4 > 0) {
      int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 16);
      cast_kernel<TY><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n4, (TY*)y);
    }
    if (n4 * 4 < n) cast_tail_kernel<TY><<<1, 4, 0, (cudaStream_t)stream>>>(x, n4 * 4, n, (TY*)y);
  });
  B200_LAUNCH_OK();
  return 0;
}

// Column sums, deterministic, ONE launch: (1) each block sums a 256-row slab of 32 columns into ws[slab][col]; (2) the
// block that finishes LAST for its column tile (a counter per tile) adds the slabs in ascending order -- the order is fixed
// whichever block that is.  Workspace = [counters: one uint32 per 32-column tile, zero on entry, restored to zero by the
// kernel] ++ [partials].  (The separate final-stage kernel was 18 extra launches per training step.)
#define CS_SLAB 256
#define CS_MAX_COLS 32768
#define CS_CNT_BYTES(cols) ((size_t)(CS_MAX_COLS / 32) * sizeof(uint32_t))   // FIXED size: calls with different widths share the buffer
template <typename TX>
__global__ void __launch_bounds__(256) colsum_kernel(const TX* __restrict__ x, int64_t ldx, int rows, int cols,
                                                     uint32_t* __restrict__ counters, float* __restrict__ ws,
                                                     float* __restrict__ out, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[8][33];
  __shared__ uint32_t s_last;
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + cx;
  int r0 = blockIdx.y * CS_SLAB, r1 = min(rows, r0 + CS_SLAB);
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + ry; r < r1; r += 8) s += to_f32(x[(int64_t)r * ldx + c]);
  part[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][cx];
    ws[(int64_t)blockIdx.y * cols + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (ry == 0 && c < cols) {
    const volatile float* w = ws;
    float t = 0.f;
    for (int sl = 0; sl < (int)gridDim.y; ++sl) t += w[(int64_t)sl * cols + c];
    out[c] = accumulate ? out[c] + t : t;
  }
  if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

size_t b200rec_colsum_workspace_bytes(int rows, int cols) {
  return CS_CNT_BYTES(std::max(cols, 1)) +
         (size_t)std::max(1, ceil_div_i(rows, CS_SLAB)) * (size_t)std::max(cols, 1) * sizeof(float);
}

int b200rec_colsum(const void* x, int x_dtype, int ldx, int rows, int cols, float* out, int accumulate,
                   void* workspace, size_t workspace_bytes, void* stream) {
  if (cols == 0) return 0;
  B200_CHECK_ARG(workspace_bytes >= b200rec_colsum_workspace_bytes(rows, cols), "colsum: workspace too small");
  B200_CHECK_ARG(cols <= CS_MAX_COLS, "colsum: at most %d columns per call (got %d)", CS_MAX_COLS, cols);
  B200_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "colsum: workspace must be 256-byte aligned");
  int slabs = std::max(1, ceil_div_i(rows, CS_SLAB));
  dim3 grid(ceil_div_i(cols, 32), slabs);
  uint32_t* counters = (uint32_t*)workspace;
  float* ws = (float*)((char*)workspace + CS_CNT_BYTES(cols));
  DISPATCH_ACT(x_dtype, TX, {
    B200_CUDA_OK(launch_pdl(colsum_kernel<TX>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const TX*)x, ldx, rows, cols, counters, ws, out, accumulate));
  });
  B200_LAUNCH_OK();
  return 0;
}

// single block, fixed-order tree: deterministic scalar reduction (loss sums, dscale)
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                          float* __restrict__ out, int accumulate) {
  __shared__ float red[40];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + scale * s : scale * s;
}

int b200rec_reduce_sum(const float* x, int64_t n, float scale, float* out, int accumulate, void* stream) {
  reduce_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out, accumulate);
  B200_LAUNCH_OK();
  return 0;
}

__global__ void counter_add_kernel(int64_t* c, int64_t v) { *c += v; }
extern "C" int b200rec_counter_add(int64_t* counter_dev, int64_t v, void* stream) {
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, v);
  B200_LAUNCH_OK();
  return 0;
}
