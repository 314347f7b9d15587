// Item-embedding gather and its deterministic sorted-segment gradient.
// HBM-bound: 16-byte vector accesses, coalesced along the row, several rows in flight per thread.
#include <cub/cub.cuh>

#include "common.cuh"

// ------------------------------------------------------------------------------------ gather
template <typename TO>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, int D4,
                                                          const int64_t* __restrict__ ids, int64_t n_vec,
                                                          TO* __restrict__ out) {
  // flat index over (row, 4-float vector); 4 independent vectors per thread per iteration
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n_vec; i += 4 * stride) {
    float v[4][4];
    int64_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = i + u * stride;
      if (idx[u] < n_vec) {
        int64_t r = idx[u] / D4;
        int c = (int)(idx[u] - r * D4);
        int64_t id = __ldg(ids + r);
        load4<float>(table + (id * D4 + c) * 4, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (idx[u] < n_vec) store4<TO>(out + idx[u] * 4, v[u]);
  }
}

int b200rec_gather_rows(const float* table, int D, const int64_t* ids, int64_t n_ids, void* out,
                        int out_dtype, void* stream) {
  B200_CHECK_ARG(D % 4 == 0, "gather_rows: D=%d must be a multiple of 4", D);
  if (n_ids == 0) return 0;
  int64_t n_vec = n_ids * (D / 4);
  int blocks = (int)std::min<int64_t>((n_vec + 256 * 4 - 1) / (256 * 4), 148 * 16);
  DISPATCH_ACT(out_dtype, TO, {
    gather_rows_kernel<TO><<<blocks, 256, 0, (cudaStream_t)stream>>>(table, D / 4, ids, n_vec, (TO*)out);
  });
  B200_LAUNCH_OK();
  return 0;
}

__global__ void __launch_bounds__(256) embed_tokens_kernel(const float* __restrict__ table,
                                                           const float* __restrict__ pos_table,
                                                           const int64_t* __restrict__ items,
                                                           const int32_t* __restrict__ tok_b,
                                                           const int32_t* __restrict__ tok_pos, int64_t n_vec,
                                                           int LP, int D4, float* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    int64_t t = i / D4;
    int c = (int)(i - t * D4);
    int pos = __ldg(tok_pos + t);
    int64_t id = __ldg(items + (int64_t)__ldg(tok_b + t) * LP + pos);
    float a[4], b[4];
    load4<float>(table + (id * D4 + c) * 4, a);
    load4<float>(pos_table + ((int64_t)pos * D4 + c) * 4, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] += b[k];
    store4<float>(x + i * 4, a);
  }
}

int b200rec_embed_tokens(const float* table, const float* pos_table, const int64_t* items,
                         const int32_t* tok_b, const int32_t* tok_pos, int T, int LP, int D, float* x,
                         void* stream) {
  B200_CHECK_ARG(D % 4 == 0, "embed_tokens: D=%d must be a multiple of 4", D);
  if (T == 0) return 0;
  int64_t n_vec = (int64_t)T * (D / 4);
  int blocks = (int)std::min<int64_t>((n_vec + 255) / 256, 148 * 16);
  embed_tokens_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(table, pos_table, items, tok_b, tok_pos,
                                                                n_vec, LP, D / 4, x);
  B200_LAUNCH_OK();
  return 0;
}

// one warp per row: row stays in registers between the norm reduction and the scaled store
template <typename TO, int MAXV>
__global__ void __launch_bounds__(256) gather_l2norm_kernel(const float* __restrict__ table,
                                                            const float* __restrict__ rows_in, int D4,
                                                            const int64_t* __restrict__ ids, int64_t n,
                                                            TO* __restrict__ out, float* __restrict__ inv_norm) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* src = table ? table + __ldg(ids + r) * (int64_t)D4 * 4 : rows_in + r * (int64_t)D4 * 4;
  float v[MAXV][4];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      load4<float>(src + c * 4, v[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) ss += v[u][k] * v[u][k];
    }
  }
  ss = warp_sum(ss);
  float inv = 1.f / sqrtf(ss);
  if (lane == 0) inv_norm[r] = inv;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] *= inv;
      store4<TO>(out + (r * D4 + c) * 4, v[u]);
    }
  }
}

int b200rec_gather_l2norm(const float* table, const float* rows_in, int D, const int64_t* ids, int64_t n,
                          void* out_hat, int out_dtype, float* inv_norm, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048, "gather_l2norm: D=%d must be a multiple of 4 and <= 2048", D);
  B200_CHECK_ARG(table != nullptr || rows_in != nullptr, "gather_l2norm: no source");
  if (n == 0) return 0;
  int blocks = ceil_div_i(n, 8);
  DISPATCH_ACT(out_dtype, TO, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(gather_l2norm_kernel<TO, 4>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, table, rows_in, D / 4, ids, n,
                                                                            (TO*)out_hat, inv_norm));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(gather_l2norm_kernel<TO, 8>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, table, rows_in, D / 4, ids, n,
                                                                            (TO*)out_hat, inv_norm));
    else
      B200_CUDA_OK(launch_pdl(gather_l2norm_kernel<TO, 16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, table, rows_in, D / 4, ids, n,
                                                                            (TO*)out_hat, inv_norm));
  });
  B200_LAUNCH_OK();
  return 0;
}

template <typename TA, int MAXV>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const TA* __restrict__ xh, const float* __restrict__ inv_norm,
                                                         const float* __restrict__ dxh, int64_t n, int D4,
                                                         float* __restrict__ dx, int accumulate) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  float a[MAXV][4], g[MAXV][4];
  float dot = 0.f;
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      load4<TA>(xh + (r * D4 + c) * 4, a[u]);
      load4<float>(dxh + (r * D4 + c) * 4, g[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) dot += a[u][k] * g[u][k];
    }
  }
  dot = warp_sum(dot);
  float inv = inv_norm[r];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    int c = lane + u * 32;
    if (c < D4) {
      float o[4];
      if (accumulate) load4<float>(dx + (r * D4 + c) * 4, o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = (g[u][k] - a[u][k] * dot) * inv;
        o[k] = accumulate ? o[k] + d : d;
      }
      store4<float>(dx + (r * D4 + c) * 4, o);
    }
  }
}

int b200rec_l2norm_bwd(const void* x_hat, int act_dtype, const float* inv_norm, const float* d_xhat, int64_t n,
                       int D, float* dx, int accumulate, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && D <= 2048, "l2norm_bwd: D=%d must be a multiple of 4 and <= 2048", D);
  if (n == 0) return 0;
  int blocks = ceil_div_i(n, 8);
  DISPATCH_ACT(act_dtype, TA, {
    if (D <= 512)
      B200_CUDA_OK(launch_pdl(l2norm_bwd_kernel<TA, 4>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const TA*)x_hat, inv_norm, d_xhat, n,
                                                                         D / 4, dx, accumulate));
    else if (D <= 1024)
      B200_CUDA_OK(launch_pdl(l2norm_bwd_kernel<TA, 8>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const TA*)x_hat, inv_norm, d_xhat, n,
                                                                         D / 4, dx, accumulate));
    else
      B200_CUDA_OK(launch_pdl(l2norm_bwd_kernel<TA, 16>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const TA*)x_hat, inv_norm, d_xhat, n,
                                                                         D / 4, dx, accumulate));
  });
  B200_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------- sorted-segment scatter-add
// workspace layout (all 256-byte aligned): keys_in, keys_out (u32[n]); pos_in, pos_out (i32[n]);
// seg (i32[n]); seg_start (i32[n+1]); cub temp.
struct ScatterWs {
  uint32_t *keys_in, *keys_out;
  int32_t *pos_in, *pos_out, *seg, *seg_start;
  int32_t* long_list;
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static ScatterWs scatter_layout(int64_t n, void* base) {
  ScatterWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += align256(bytes);
    return q;
  };
  w.keys_in = (uint32_t*)take(n * 4);
  w.keys_out = (uint32_t*)take(n * 4);
  w.pos_in = (int32_t*)take(n * 4);
  w.pos_out = (int32_t*)take(n * 4);
  w.seg = (int32_t*)take(n * 4);
  w.seg_start = (int32_t*)take((n + 1) * 4);
  w.long_list = (int32_t*)take((n / 32 + 2) * 4);   // [0] = count, then ids of segments longer than 32 rows
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  w.cub_bytes = std::max(sort_bytes, scan_bytes);
  w.cub_tmp = take(w.cub_bytes);
  w.total = off;
  return w;
}

size_t b200rec_scatter_add_workspace_bytes(int64_t n_ids) {
  if (n_ids <= 0) return 256;
  return scatter_layout(n_ids, nullptr).total;
}

__global__ void scatter_prep_kernel(const int64_t* __restrict__ ids, int n, uint32_t* keys, int32_t* pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t id = ids[i];
  // padding_idx 0 and negative ids carry no gradient: park them behind every real key
  keys[i] = (id <= 0 || id >= 0xffffffffll) ? 0xffffffffu : (uint32_t)id;
  pos[i] = i;
}

__global__ void scatter_flag_kernel(const uint32_t* __restrict__ keys, int n, int32_t* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  flag[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// seg[i] = inclusive count of segment heads up to i  ->  segment index = seg[i]-1 for valid keys
__global__ void scatter_bounds_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ seg, int n,
                                      int32_t* seg_start, int64_t* uniq_ids, int32_t* n_uniq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  bool valid = k != 0xffffffffu;
  bool head = valid && (i == 0 || keys[i - 1] != k);
  if (head) {
    seg_start[seg[i] - 1] = i;
    uniq_ids[seg[i] - 1] = (int64_t)k;
  }
  // end sentinel: first invalid position (or n)
  bool last_valid = valid && (i == n - 1 || keys[i + 1] == 0xffffffffu);
  if (last_valid) {
    seg_start[seg[i]] = i + 1;
    *n_uniq = seg[i];
  }
  if (i == 0 && !valid) *n_uniq = 0;
}

// Deterministic segment sums.  Summation order (part of the contract, tests replicate it):
//   * a segment of <= 32 rows: rows added one by one in ascending input position;
//   * a longer segment: cut into chunks of 32 consecutive rows; chunk j is summed row by row; warp w (0..7) adds
//     the chunk sums j = w, w+8, w+16, ... in ascending j; the 8 warp totals are added in order w = 0..7.
// Short segments take one warp each (a lane owns every 32nd float4 column, 4 rows of loads in flight); long ones
// (hot items of a Zipf catalogue) take a whole block so that one id cannot serialise the kernel.
#define SEG_CH 32

// Where gradient row `pos` lives: one local array, or (peer variant) W equally long arrays -- one per rank, mapped
// through CUDA IPC -- addressed src-major: pos = src * n_per + index.
struct RowSrc {
  const float* base;
  const float* const* srcs;
  int n_per;
  __device__ __forceinline__ const float* row(int pos, int D4) const {
    if (srcs == nullptr) return base + (int64_t)pos * D4 * 4;
    const int s = pos / n_per;
    return srcs[s] + (int64_t)(pos - s * n_per) * D4 * 4;
  }
};

template <int MAXV>
__device__ __forceinline__ void seg_sum_rows(const RowSrc grad_rows, const int32_t* __restrict__ pos_sorted,
                                             int beg, int end, int D4, int c0, float (&acc)[MAXV][4]) {
  const int lane = threadIdx.x & 31;
  int i = beg;
  for (; i + 4 <= end; i += 4) {
    float v[4][MAXV][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* row = grad_rows.row(pos_sorted[i + r], D4);
#pragma unroll
      for (int u = 0; u < MAXV; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < D4) load4<float>(row + c * 4, v[r][u]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int u = 0; u < MAXV; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[u][k] += v[r][u][k];
  }
  for (; i < end; ++i) {
    const float* row = grad_rows.row(pos_sorted[i], D4);
    float v[MAXV][4];
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      const int c = c0 + lane + 32 * u;
      if (c < D4) load4<float>(row + c * 4, v[u]);
    }
#pragma unroll
    for (int u = 0; u < MAXV; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[u][k] += v[u][k];
  }
}

#define SEG_V 4   // float4 columns per lane and pass: 4 rows x 4 vectors x 16 B in flight per lane

__global__ void __launch_bounds__(256) scatter_reduce_short_kernel(const int32_t* __restrict__ seg_start,
                                                                   const int32_t* __restrict__ pos_sorted,
                                                                   const int32_t* __restrict__ n_uniq,
                                                                   const RowSrc grad_rows, int D4,
                                                                   float* __restrict__ uniq_rows,
                                                                   int32_t* __restrict__ long_list) {
  const int nu = *n_uniq;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nu; s += warps) {
    const int beg = seg_start[s], end = seg_start[s + 1];
    if (end - beg > SEG_CH) {
      if (lane == 0) long_list[1 + atomicAdd(&long_list[0], 1)] = s;   // list order does not affect any sum
      continue;
    }
    for (int c0 = 0; c0 < D4; c0 += 32 * SEG_V) {
      float acc[SEG_V][4];
#pragma unroll
      for (int u = 0; u < SEG_V; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[u][k] = 0.f;
      seg_sum_rows<SEG_V>(grad_rows, pos_sorted, beg, end, D4, c0, acc);
#pragma unroll
      for (int u = 0; u < SEG_V; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < D4) store4<float>(uniq_rows + ((int64_t)s * D4 + c) * 4, acc[u]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) scatter_reduce_long_kernel(const int32_t* __restrict__ seg_start,
                                                                  const int32_t* __restrict__ pos_sorted,
                                                                  const RowSrc grad_rows, int D4,
                                                                  float* __restrict__ uniq_rows,
                                                                  const int32_t* __restrict__ long_list) {
  __shared__ float part[8][32 * SEG_V * 4];
  const int n_long = long_list[0];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = blockIdx.x; q < n_long; q += gridDim.x) {
    const int s = long_list[1 + q];
    const int beg = seg_start[s], end = seg_start[s + 1];
    const int n_chunks = (end - beg + SEG_CH - 1) / SEG_CH;
    for (int c0 = 0; c0 < D4; c0 += 32 * SEG_V) {
      float tot[SEG_V][4];
#pragma unroll
      for (int u = 0; u < SEG_V; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) tot[u][k] = 0.f;
      for (int j = w; j < n_chunks; j += 8) {
        float acc[SEG_V][4];
#pragma unroll
        for (int u = 0; u < SEG_V; ++u)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[u][k] = 0.f;
        const int cb = beg + j * SEG_CH;
        seg_sum_rows<SEG_V>(grad_rows, pos_sorted, cb, min(end, cb + SEG_CH), D4, c0, acc);
#pragma unroll
        for (int u = 0; u < SEG_V; ++u)
#pragma unroll
          for (int k = 0; k < 4; ++k) tot[u][k] += acc[u][k];
      }
#pragma unroll
      for (int u = 0; u < SEG_V; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) part[w][(lane + 32 * u) * 4 + k] = tot[u][k];
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * SEG_V * 4; e += blockDim.x) {
        const int col = c0 * 4 + e;
        if (col < D4 * 4) {
          float sum = part[0][e];
#pragma unroll
          for (int ww = 1; ww < 8; ++ww) sum += part[ww][e];
          uniq_rows[(int64_t)s * D4 * 4 + col] = sum;
        }
      }
      __syncthreads();
    }
  }
}

// keys of the sharded variant: global id g owned by `rank` (g % W == rank, g != 0) -> local row g / W; everything
// else (other ranks' ids, padding id 0, negative fillers) is parked behind the real keys
__global__ void scatter_prep_sharded_kernel(const int64_t* __restrict__ ids, int n, int W, int rank, uint32_t* keys,
                                            int32_t* pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t id = ids[i];
  const bool mine = id > 0 && (int)(id % W) == rank && id / W < 0xffffffffll;
  keys[i] = mine ? (uint32_t)(id / W) : 0xffffffffu;
  pos[i] = i;
}

static int scatter_add_run(const int64_t* ids, int64_t n_ids, RowSrc src, int W, int rank, int D, int64_t* uniq_ids,
                           float* uniq_rows, int32_t* n_uniq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  B200_CHECK_ARG(D % 4 == 0, "scatter_add_sorted: D=%d must be a multiple of 4", D);
  B200_CHECK_ARG(n_ids < (1ll << 31), "scatter_add_sorted: too many ids");
  if (n_ids == 0) {
    B200_CUDA_OK(cudaMemsetAsync(n_uniq, 0, 4, st));
    return 0;
  }
  ScatterWs w = scatter_layout(n_ids, workspace);
  B200_CHECK_ARG(workspace_bytes >= w.total, "scatter_add_sorted: workspace %zu < %zu", workspace_bytes, w.total);
  int n = (int)n_ids;
  int blocks = ceil_div_i(n, 256);
  if (W > 0) scatter_prep_sharded_kernel<<<blocks, 256, 0, st>>>(ids, n, W, rank, w.keys_in, w.pos_in);
  else scatter_prep_kernel<<<blocks, 256, 0, st>>>(ids, n, w.keys_in, w.pos_in);
  size_t tmp = w.cub_bytes;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.pos_in, w.pos_out, n, 0, 32,
                                               st));
  scatter_flag_kernel<<<blocks, 256, 0, st>>>(w.keys_out, n, w.pos_in /* reuse as flag */);
  tmp = w.cub_bytes;
  B200_CUDA_OK(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, w.pos_in, w.seg, n, st));
  scatter_bounds_kernel<<<blocks, 256, 0, st>>>(w.keys_out, w.seg, n, w.seg_start, uniq_ids, n_uniq);
  B200_CUDA_OK(cudaMemsetAsync(w.long_list, 0, 4, st));
  int rblocks = std::min(ceil_div_i(n, 8), 148 * 8);
  scatter_reduce_short_kernel<<<rblocks, 256, 0, st>>>(w.seg_start, w.pos_out, n_uniq, src, D / 4, uniq_rows,
                                                       w.long_list);
  scatter_reduce_long_kernel<<<std::min(n / SEG_CH + 1, 148 * 2), 256, 0, st>>>(w.seg_start, w.pos_out, src, D / 4,
                                                                               uniq_rows, w.long_list);
  B200_LAUNCH_OK();
  return 0;
}

int b200rec_scatter_add_sorted(const int64_t* ids, int64_t n_ids, const float* grad_rows, int D, int64_t* uniq_ids,
                               float* uniq_rows, int32_t* n_uniq, void* workspace, size_t workspace_bytes,
                               void* stream) {
  RowSrc src;
  src.base = grad_rows; src.srcs = nullptr; src.n_per = 0;
  return scatter_add_run(ids, n_ids, src, 0, 0, D, uniq_ids, uniq_rows, n_uniq, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

// Owner-side reduction of a row-sharded table's gradient: ids = the global ids behind the W x n_per gradient rows of
// all ranks (rank-major), src_ptrs_dev[r] = rank r's gradient-row buffer [n_per, D] mapped through CUDA IPC.  Keeps the
// ids this rank owns (id % W == rank, id != 0), returns LOCAL row indices (id / W) in uniq_ids and sums the rows of
// each in ascending (rank, index) order -- the rows are read over NVLink while they are summed.
int b200rec_scatter_add_sorted_peer(const int64_t* ids, int64_t n_per, int W, int rank, const void* src_ptrs_dev, int D,
                                    int64_t* uniq_ids, float* uniq_rows, int32_t* n_uniq, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  B200_CHECK_ARG(W >= 1 && rank >= 0 && rank < W && n_per * W < (1ll << 31), "scatter_add_sorted_peer: bad W / rank / n");
  RowSrc src;
  src.base = nullptr; src.srcs = (const float* const*)src_ptrs_dev; src.n_per = (int)n_per;
  return scatter_add_run(ids, n_per * W, src, W, rank, D, uniq_ids, uniq_rows, n_uniq, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

__global__ void __launch_bounds__(256) rows_to_dense_kernel(const int64_t* __restrict__ uniq_ids,
                                                            const float* __restrict__ uniq_rows,
                                                            const int32_t* __restrict__ n_uniq, int D4,
                                                            float* __restrict__ dense, int accumulate) {
  int64_t n_vec = (int64_t)(*n_uniq) * D4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    int64_t r = i / D4;
    int c = (int)(i - r * D4);
    float v[4];
    load4<float>(uniq_rows + i * 4, v);
    float* dst = dense + (uniq_ids[r] * D4 + c) * 4;
    if (accumulate) {
      float o[4];
      load4<float>(dst, o);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] += o[k];
    }
    store4<float>(dst, v);
  }
}

int b200rec_rows_to_dense(const int64_t* uniq_ids, const float* uniq_rows, const int32_t* n_uniq, int64_t max_rows,
                          int D, float* dense, int accumulate, void* stream) {
  B200_CHECK_ARG(D % 4 == 0, "rows_to_dense: D=%d must be a multiple of 4", D);
  if (max_rows == 0) return 0;
  int blocks = (int)std::min<int64_t>((max_rows * (D / 4) + 255) / 256, 148 * 16);
  rows_to_dense_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(uniq_ids, uniq_rows, n_uniq, D / 4, dense,
                                                                 accumulate);
  B200_LAUNCH_OK();
  return 0;
}

// ---- position-embedding gradient: d_pos[pos,:] = sum_b dx0[tok_index[b*LP+pos],:]  (b ascending) -----
__global__ void __launch_bounds__(256) pos_emb_grad_kernel(const float* __restrict__ dx0,
                                                           const int32_t* __restrict__ tok_index, int B, int LP,
                                                           int D4, float* __restrict__ d_pos) {
  const int pos = blockIdx.x;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < B; ++b) {
      int t = tok_index[(int64_t)b * LP + pos];
      if (t >= 0) {
        float v[4];
        load4<float>(dx0 + ((int64_t)t * D4 + c) * 4, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += v[k];
      }
    }
    store4<float>(d_pos + ((int64_t)pos * D4 + c) * 4, acc);
  }
}

extern "C" int b200rec_pos_emb_grad(const float* dx0, const int32_t* tok_index, int B, int LP, int L, int D,
                                    float* d_pos, void* stream) {
  B200_CHECK_ARG(D % 4 == 0, "pos_emb_grad: D must be a multiple of 4");
  if (L == 0) return 0;
  pos_emb_grad_kernel<<<L, 256, 0, (cudaStream_t)stream>>>(dx0, tok_index, B, LP, D / 4, d_pos);
  B200_LAUNCH_OK();
  return 0;
}

// ---- decode-head ResBlock backward (llm_heads.py:26-40): dz = d_hd * silu'(z); dy = sum_h d_hd[:, h, :]
template <typename TA>
__global__ void __launch_bounds__(256) resblock_bwd_kernel(const float* __restrict__ d_hd, const TA* __restrict__ z,
                                                           int64_t T, int H, int D4, TA* __restrict__ dz,
                                                           float* __restrict__ dy) {
  pdl_trigger();
  const int64_t n_vec = T * D4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    int64_t t = i / D4;
    int c = (int)(i - t * D4);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < H; ++h) {
      int64_t off = ((t * H + h) * D4 + c) * 4;
      float g[4];
      load4<float>(d_hd + off, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += g[k];
      if (z) {
        float zz[4];
        load4<TA>(z + off, zz);
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] *= silu_grad_f(zz[k]);
        store4<TA>(dz + off, g);
      }
    }
    store4<float>(dy + i * 4, acc);
  }
}

extern "C" int b200rec_resblock_bwd(const float* d_hd, const void* z, int act_dtype, int64_t T, int H, int D,
                                    void* dz, float* dy, void* stream) {
  B200_CHECK_ARG(D % 4 == 0, "resblock_bwd: D must be a multiple of 4");
  if (T == 0) return 0;
  int blocks = (int)std::min<int64_t>((T * (D / 4) + 255) / 256, 148 * 16);
  DISPATCH_ACT(act_dtype, TA, {
    resblock_bwd_kernel<TA><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_hd, (const TA*)z, T, H, D / 4, (TA*)dz, dy);
  });
  B200_LAUNCH_OK();
  return 0;
}
