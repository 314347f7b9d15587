// HSTU pointwise-SiLU causal attention over jagged sequences (hstu.py:137-160):
//   A = silu(q k^T) / n_pad * [key valid & j <= i],  out = A v       (no softmax, no 1/sqrt(d))
// Round-1 implementation: fp32 CUDA-core tiles (64 queries x 64 keys) held in shared memory, the
// [L, L] score matrix never touches HBM, causal key tiles above the diagonal are skipped.
// Backward recomputes S in two deterministic passes (dQ by query tile, dK/dV by
// key tile) and fuses the silu'(pre) factor of the uvqk activation into its stores.
#include "common.cuh"

#define AT 64  // tile edge (queries and keys)

template <typename TA, int DH>
__device__ __forceinline__ void load_tile(float (*dst)[DH + 1], const TA* __restrict__ src, int64_t ld, int row0,
                                          int n_rows) {
  // 64 x DH tile; rows beyond n_rows are zero
  constexpr int V = DH / 4;
  for (int idx = threadIdx.x; idx < AT * V; idx += blockDim.x) {
    int r = idx / V, c = idx - r * V;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n_rows) load4<TA>(src + (int64_t)(row0 + r) * ld + c * 4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[r][c * 4 + k] = v[k];
  }
}

// P[i][j] = sum_d X[i][d] * Y[j][d]   (thread (ty,tx): i = ty+16a, j = tx+16b)
template <int DH>
__device__ __forceinline__ void tile_xyT(const float (*X)[DH + 1], const float (*Y)[DH + 1], float (&acc)[4][4]) {
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
  for (int d = 0; d < DH; ++d) {
    float x[4], y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) x[a] = X[ty + 16 * a][d];
#pragma unroll
    for (int b = 0; b < 4; ++b) y[b] = Y[tx + 16 * b][d];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(x[a], y[b], acc[a][b]);
  }
}

// O[i][c] += sum_j S[i][j] * V[j][c]      (i = ty+16a, c = tx+16e)
template <int DH>
__device__ __forceinline__ void tile_sv(const float (*S)[AT + 1], const float (*V)[DH + 1], float (&o)[4][DH / 16]) {
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 8
  for (int j = 0; j < AT; ++j) {
    float s[4], v[DH / 16];
#pragma unroll
    for (int a = 0; a < 4; ++a) s[a] = S[ty + 16 * a][j];
#pragma unroll
    for (int e = 0; e < DH / 16; ++e) v[e] = V[j][tx + 16 * e];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < DH / 16; ++e) o[a][e] = fmaf(s[a], v[e], o[a][e]);
  }
}

// O[j][c] += sum_i S[i][j] * X[i][c]      (transposed use of S; j = ty+16a, c = tx+16e)
template <int DH>
__device__ __forceinline__ void tile_sTx(const float (*S)[AT + 1], const float (*X)[DH + 1],
                                         float (&o)[4][DH / 16]) {
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 8
  for (int i = 0; i < AT; ++i) {
    float s[4], x[DH / 16];
#pragma unroll
    for (int a = 0; a < 4; ++a) s[a] = S[i][ty + 16 * a];
#pragma unroll
    for (int e = 0; e < DH / 16; ++e) x[e] = X[i][tx + 16 * e];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < DH / 16; ++e) o[a][e] = fmaf(s[a], x[e], o[a][e]);
  }
}

template <int DH>
struct AttnSmem {
  float q[AT][DH + 1];
  float k[AT][DH + 1];
  float v[AT][DH + 1];
  float dout[AT][DH + 1];
  float s[AT][AT + 1];
  float ds[AT][AT + 1];
  uint8_t kvalid[AT];
};

template <typename TA, int DH>
__global__ void __launch_bounds__(256) hstu_attn_fwd_kernel(const TA* __restrict__ q, const TA* __restrict__ k,
                                                            const TA* __restrict__ v, int64_t ld,
                                                            const int32_t* __restrict__ seq_off,
                                                            const uint8_t* __restrict__ key_valid, float inv_n,
                                                            float* __restrict__ out, int D,
                                                            const float* __restrict__ bias_d) {
  // bias_d (nullable): relative position bias by query-key distance, A = silu(q k^T + bias_d[i - j]) / n_pad
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem<DH>& sm = *reinterpret_cast<AttnSmem<DH>*>(smem_raw);
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
  const int q0 = qt * AT;
  if (q0 >= len) return;
  const int nq = min(AT, len - q0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  load_tile<TA, DH>(sm.q, q + h * DH, ld, t0 + q0, nq);
  float o[4][DH / 16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < DH / 16; ++e) o[a][e] = 0.f;
  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    const int nk = min(AT, len - k0);
    __syncthreads();
    load_tile<TA, DH>(sm.k, k + h * DH, ld, t0 + k0, nk);
    load_tile<TA, DH>(sm.v, v + h * DH, ld, t0 + k0, nk);
    if (threadIdx.x < AT) sm.kvalid[threadIdx.x] = threadIdx.x < nk ? key_valid[t0 + k0 + threadIdx.x] : 0;
    __syncthreads();
    float acc[4][4];
    tile_xyT<DH>(sm.q, sm.k, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        int i = ty + 16 * a, j = tx + 16 * bb;
        bool keep = (k0 + j <= q0 + i) && sm.kvalid[j];
        const float sb = (keep && bias_d) ? acc[a][bb] + __ldg(bias_d + (q0 + i) - (k0 + j)) : acc[a][bb];
        sm.s[i][j] = keep ? silu_f(sb) * inv_n : 0.f;
      }
    __syncthreads();
    tile_sv<DH>(sm.s, sm.v, o);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int i = ty + 16 * a;
    if (i < nq) {
#pragma unroll
      for (int e = 0; e < DH / 16; ++e) out[(int64_t)(t0 + q0 + i) * D + h * DH + tx + 16 * e] = o[a][e];
    }
  }
}

// dQ pass: one block per (sequence, head, query tile)
template <typename TA, int DH>
__global__ void __launch_bounds__(256)
hstu_attn_bwd_dq_kernel(const TA* __restrict__ q, const TA* __restrict__ k, const TA* __restrict__ v,
                        const TA* __restrict__ pre_q, int64_t ld, const int32_t* __restrict__ seq_off,
                        const uint8_t* __restrict__ key_valid, float inv_n, const TA* __restrict__ d_out, int D,
                        TA* __restrict__ d_pre_q, const float* __restrict__ bias_d, float* __restrict__ dbias_part,
                        int max_len) {
  // dbias_part (with bias_d): [gridDim.z * gridDim.y * gridDim.x, max_len] partial sums of dS by distance i - j, one
  // row per block (summed afterwards in fixed order: deterministic)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem<DH>& sm = *reinterpret_cast<AttnSmem<DH>*>(smem_raw);
  const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
  const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
  const int q0 = qt * AT;
  if (q0 >= len) return;
  const int nq = min(AT, len - q0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  load_tile<TA, DH>(sm.q, q + h * DH, ld, t0 + q0, nq);
  load_tile<TA, DH>(sm.dout, d_out + h * DH, D, t0 + q0, nq);
  float dq[4][DH / 16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < DH / 16; ++e) dq[a][e] = 0.f;
  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    const int nk = min(AT, len - k0);
    __syncthreads();
    load_tile<TA, DH>(sm.k, k + h * DH, ld, t0 + k0, nk);
    load_tile<TA, DH>(sm.v, v + h * DH, ld, t0 + k0, nk);
    if (threadIdx.x < AT) sm.kvalid[threadIdx.x] = threadIdx.x < nk ? key_valid[t0 + k0 + threadIdx.x] : 0;
    __syncthreads();
    float s[4][4], da[4][4];
    tile_xyT<DH>(sm.q, sm.k, s);
    tile_xyT<DH>(sm.dout, sm.v, da);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        int i = ty + 16 * a, j = tx + 16 * bb;
        bool keep = (k0 + j <= q0 + i) && sm.kvalid[j];
        const float sb = (keep && bias_d) ? s[a][bb] + __ldg(bias_d + (q0 + i) - (k0 + j)) : s[a][bb];
        sm.ds[i][j] = keep ? da[a][bb] * inv_n * silu_grad_f(sb) : 0.f;
      }
    __syncthreads();
    tile_sv<DH>(sm.ds, sm.k, dq);
    if (dbias_part != nullptr) {
      // distances of this tile pair: d = (q0 - k0) + (i - j) in [q0 - k0 - 63, q0 - k0 + 63]; thread x sums one diagonal
      float* mine = dbias_part + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * max_len;
      if (threadIdx.x < 2 * AT - 1) {
        const int rel = (int)threadIdx.x - (AT - 1);           // i - j
        const int d = q0 - k0 + rel;
        if (d >= 0 && d < max_len) {
          float acc_d = 0.f;
          for (int i = max(0, rel); i < AT && i - rel < AT; ++i) acc_d += sm.ds[i][i - rel];
          mine[d] += acc_d;                                    // each (block, d) is touched by one thread per key tile
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int i = ty + 16 * a;
    if (i < nq) {
#pragma unroll
      for (int e = 0; e < DH / 16; ++e) {
        int64_t off = (int64_t)(t0 + q0 + i) * ld + h * DH + tx + 16 * e;
        d_pre_q[off] = from_f32<TA>(dq[a][e] * silu_grad_f(to_f32(pre_q[off])));
      }
    }
  }
}

// dK/dV pass: one block per (sequence, head, key tile); loops over query tiles at or below it
template <typename TA, int DH>
__global__ void __launch_bounds__(256)
hstu_attn_bwd_dkv_kernel(const TA* __restrict__ q, const TA* __restrict__ k, const TA* __restrict__ v,
                         const TA* __restrict__ pre_k, const TA* __restrict__ pre_v, int64_t ld,
                         const int32_t* __restrict__ seq_off, const uint8_t* __restrict__ key_valid, float inv_n,
                         const TA* __restrict__ d_out, int D, TA* __restrict__ d_pre_k, TA* __restrict__ d_pre_v,
                         const float* __restrict__ bias_d) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem<DH>& sm = *reinterpret_cast<AttnSmem<DH>*>(smem_raw);
  const int b = blockIdx.z, h = blockIdx.y, kt = blockIdx.x;
  const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
  const int k0 = kt * AT;
  if (k0 >= len) return;
  const int nk = min(AT, len - k0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  load_tile<TA, DH>(sm.k, k + h * DH, ld, t0 + k0, nk);
  load_tile<TA, DH>(sm.v, v + h * DH, ld, t0 + k0, nk);
  if (threadIdx.x < AT) sm.kvalid[threadIdx.x] = threadIdx.x < nk ? key_valid[t0 + k0 + threadIdx.x] : 0;
  float dk[4][DH / 16], dv[4][DH / 16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < DH / 16; ++e) dk[a][e] = dv[a][e] = 0.f;
  const int n_qt = (len + AT - 1) / AT;
  for (int qt = kt; qt < n_qt; ++qt) {
    const int q0 = qt * AT;
    const int nq = min(AT, len - q0);
    __syncthreads();
    load_tile<TA, DH>(sm.q, q + h * DH, ld, t0 + q0, nq);
    load_tile<TA, DH>(sm.dout, d_out + h * DH, D, t0 + q0, nq);
    __syncthreads();
    float s[4][4], da[4][4];
    tile_xyT<DH>(sm.q, sm.k, s);
    tile_xyT<DH>(sm.dout, sm.v, da);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        int i = ty + 16 * a, j = tx + 16 * bb;
        bool keep = (k0 + j <= q0 + i) && sm.kvalid[j] && i < nq;
        const float sb = (keep && bias_d) ? s[a][bb] + __ldg(bias_d + (q0 + i) - (k0 + j)) : s[a][bb];
        sm.s[i][j] = keep ? silu_f(sb) * inv_n : 0.f;
        sm.ds[i][j] = keep ? da[a][bb] * inv_n * silu_grad_f(sb) : 0.f;
      }
    __syncthreads();
    tile_sTx<DH>(sm.s, sm.dout, dv);
    tile_sTx<DH>(sm.ds, sm.q, dk);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int j = ty + 16 * a;
    if (j < nk) {
#pragma unroll
      for (int e = 0; e < DH / 16; ++e) {
        int64_t off = (int64_t)(t0 + k0 + j) * ld + h * DH + tx + 16 * e;
        d_pre_k[off] = from_f32<TA>(dk[a][e] * silu_grad_f(to_f32(pre_k[off])));
        d_pre_v[off] = from_f32<TA>(dv[a][e] * silu_grad_f(to_f32(pre_v[off])));
      }
    }
  }
}

template <typename TA, int DH>
static int attn_fwd_launch(const void* q, const void* k, const void* v, int ld, const int32_t* seq_off,
                           const uint8_t* key_valid, int B, int n_heads, float inv_n, int max_len, float* out,
                           const float* bias_d, cudaStream_t st) {
  size_t smem = sizeof(AttnSmem<DH>);
  { static bool once_1 = false; if (!once_1) { B200_CUDA_OK(cudaFuncSetAttribute(hstu_attn_fwd_kernel<TA, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem)); once_1 = true; } }
  dim3 grid(ceil_div_i(max_len, AT), n_heads, B);
  hstu_attn_fwd_kernel<TA, DH><<<grid, 256, smem, st>>>((const TA*)q, (const TA*)k, (const TA*)v, ld, seq_off,
                                                        key_valid, inv_n, out, n_heads * DH, bias_d);
  B200_LAUNCH_OK();
  return 0;
}

template <typename TA, int DH>
static int attn_bwd_launch(const void* q, const void* k, const void* v, const void* pre_q, const void* pre_k,
                           const void* pre_v, int ld, const int32_t* seq_off, const uint8_t* key_valid, int B,
                           int n_heads, float inv_n, int max_len, const void* d_out, void* d_pre_q, void* d_pre_k,
                           void* d_pre_v, const float* bias_d, float* dbias_part, cudaStream_t st) {
  size_t smem = sizeof(AttnSmem<DH>);
  { static bool once_2 = false; if (!once_2) { B200_CUDA_OK(cudaFuncSetAttribute(hstu_attn_bwd_dq_kernel<TA, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem)); once_2 = true; } }
  { static bool once_3 = false; if (!once_3) { B200_CUDA_OK(cudaFuncSetAttribute(hstu_attn_bwd_dkv_kernel<TA, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem)); once_3 = true; } }
  dim3 grid(ceil_div_i(max_len, AT), n_heads, B);
  hstu_attn_bwd_dq_kernel<TA, DH><<<grid, 256, smem, st>>>((const TA*)q, (const TA*)k, (const TA*)v,
                                                           (const TA*)pre_q, ld, seq_off, key_valid, inv_n,
                                                           (const TA*)d_out, n_heads * DH, (TA*)d_pre_q, bias_d,
                                                           dbias_part, max_len);
  hstu_attn_bwd_dkv_kernel<TA, DH><<<grid, 256, smem, st>>>((const TA*)q, (const TA*)k, (const TA*)v,
                                                            (const TA*)pre_k, (const TA*)pre_v, ld, seq_off,
                                                            key_valid, inv_n, (const TA*)d_out, n_heads * DH,
                                                            (TA*)d_pre_k, (TA*)d_pre_v, bias_d);
  B200_LAUNCH_OK();
  return 0;
}

#define DISPATCH_DH(dh, DH, ...)                                              \
  do {                                                                        \
    if ((dh) == 16) { constexpr int DH = 16; __VA_ARGS__ }                    \
    else if ((dh) == 32) { constexpr int DH = 32; __VA_ARGS__ }               \
    else if ((dh) == 64) { constexpr int DH = 64; __VA_ARGS__ }               \
    else if ((dh) == 128) { constexpr int DH = 128; __VA_ARGS__ }             \
    else { b200rec_set_error("attention: unsupported head dim %d", (dh)); return 1; } \
  } while (0)

int b200rec_hstu_attn_fwd(const void* q, const void* k, const void* v, int ld, int act_dtype,
                          const int32_t* seq_off, const uint8_t* key_valid, int B, int T, int n_heads, int dh,
                          float inv_n, int max_len, float* out, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld % 4 == 0, "attention: ld=%d must be a multiple of 4", ld);
  DISPATCH_ACT(act_dtype, TA, {
    DISPATCH_DH(dh, DH, {
      return attn_fwd_launch<TA, DH>(q, k, v, ld, seq_off, key_valid, B, n_heads, inv_n, max_len, out, nullptr,
                                     (cudaStream_t)stream);
    });
  });
  return 0;
}

extern "C" int b200rec_hstu_attn_bias_fwd(const void* q, const void* k, const void* v, int ld, int act_dtype,
                                          const int32_t* seq_off, const uint8_t* key_valid, int B, int T, int n_heads,
                                          int dh, float inv_n, int max_len, const float* bias_d, float* out,
                                          void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld % 4 == 0 && bias_d != nullptr, "attention(bias): ld=%d must be a multiple of 4, bias_d non-null", ld);
  DISPATCH_ACT(act_dtype, TA, {
    DISPATCH_DH(dh, DH, {
      return attn_fwd_launch<TA, DH>(q, k, v, ld, seq_off, key_valid, B, n_heads, inv_n, max_len, out, bias_d,
                                     (cudaStream_t)stream);
    });
  });
  return 0;
}

extern "C" size_t b200rec_hstu_attn_bias_ws_floats(int B, int n_heads, int max_len) {
  return (size_t)B * n_heads * ((max_len + AT - 1) / AT) * max_len;
}

// backward with the relative position bias: also returns dbias_part [B * n_heads * ceil(max_len / 64), max_len]
// (ZERO-FILLED by the caller); d bias_d[d] = column sums of it.
extern "C" int b200rec_hstu_attn_bias_bwd(const void* q, const void* k, const void* v, const void* pre_q,
                                          const void* pre_k, const void* pre_v, int ld, int act_dtype,
                                          const int32_t* seq_off, const uint8_t* key_valid, int B, int T, int n_heads,
                                          int dh, float inv_n, int max_len, const float* bias_d, const void* d_out,
                                          void* d_pre_q, void* d_pre_k, void* d_pre_v, float* dbias_part, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld % 4 == 0 && bias_d != nullptr && dbias_part != nullptr, "attention(bias) bwd: bad args");
  DISPATCH_ACT(act_dtype, TA, {
    DISPATCH_DH(dh, DH, {
      return attn_bwd_launch<TA, DH>(q, k, v, pre_q, pre_k, pre_v, ld, seq_off, key_valid, B, n_heads, inv_n, max_len,
                                     d_out, d_pre_q, d_pre_k, d_pre_v, bias_d, dbias_part, (cudaStream_t)stream);
    });
  });
  return 0;
}

int b200rec_hstu_attn_bwd(const void* q, const void* k, const void* v, const void* pre_q, const void* pre_k,
                          const void* pre_v, int ld, int act_dtype, const int32_t* seq_off,
                          const uint8_t* key_valid, int B, int T, int n_heads, int dh, float inv_n, int max_len,
                          const void* d_out, void* d_pre_q, void* d_pre_k, void* d_pre_v, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld % 4 == 0, "attention: ld=%d must be a multiple of 4", ld);
  DISPATCH_ACT(act_dtype, TA, {
    DISPATCH_DH(dh, DH, {
      return attn_bwd_launch<TA, DH>(q, k, v, pre_q, pre_k, pre_v, ld, seq_off, key_valid, B, n_heads, inv_n,
                                     max_len, d_out, d_pre_q, d_pre_k, d_pre_v, nullptr, nullptr, (cudaStream_t)stream);
    });
  });
  return 0;
}
