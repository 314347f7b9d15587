// Full-catalogue eval: prior / id-0 / history masks, cross-head merge and top-K.
// The reference takes a per-head top-K, sorts the H*K candidates and dedupes sequentially
// (collector.py:241-275); that equals top-K of the max over heads, so one pass folds
// the H masked scores of an item into (max, argmax head) and a per-user radix select + bitonic
// sort produces the list.  Tie rule: value desc, item id asc, head asc.
#include "common.cuh"

__device__ __forceinline__ uint32_t order_key(float v) {  // ascending-order-preserving map
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- pass 1: fold heads ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fold_heads_kernel(const float* __restrict__ scores, int64_t ld, int H, int64_t N,
                  const int32_t* __restrict__ head_cat, const uint32_t* __restrict__ item_tag_bits,
                  const uint8_t* __restrict__ head_on, int split_mode, int64_t id_offset, int64_t id_stride,
                  float* __restrict__ fval, uint8_t* __restrict__ fhead) {
  const int b = blockIdx.y;
  __shared__ int s_cat[64];
  __shared__ uint8_t s_on[64];
  if (threadIdx.x < H) {
    s_cat[threadIdx.x] = head_cat ? head_cat[threadIdx.x] : -1;
    s_on[threadIdx.x] = head_on ? head_on[b * H + threadIdx.x] : 1;
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float best = -INFINITY;
    int bh = 0;
    float sum = 0.f;
    int nfin = 0;
    if (i * id_stride + id_offset != 0) {  // trainer.py:724  scores[:, :, 0] = -inf (global item id 0)
      uint32_t tags = item_tag_bits ? item_tag_bits[i] : 0xffffffffu;
      for (int h = 0; h < H; ++h) {
        float v = scores[((int64_t)b * H + h) * ld + i];
        bool on = s_on[h] && (s_cat[h] < 0 || ((tags >> s_cat[h]) & 1u));
        v = on ? v : -INFINITY;
        if (v > best) { best = v; bh = h; }  // strict > keeps the lowest head on ties
        if (v > -INFINITY && v < INFINITY) { sum += v; ++nfin; }
      }
    }
    if (split_mode == 1 && H > 1) {  // 'average': mean over finite heads (collector.py:227-230)
      best = sum / ((float)nfin + 1e-8f);
      bh = 0;
    }
    fval[(int64_t)b * N + i] = best;
    fhead[(int64_t)b * N + i] = (uint8_t)bh;
  }
}

// history suppression (trainer.py:725-726): fval[b, item] = -inf
__global__ void suppress_history_kernel(const int32_t* __restrict__ hist_off, const int64_t* __restrict__ hist_items,
                                        int B, int64_t N, int64_t ld, float* __restrict__ fval, int split_mode,
                                        int64_t id_offset, int64_t id_stride) {
  int b = blockIdx.x;
  for (int i = hist_off[b] + threadIdx.x; i < hist_off[b + 1]; i += blockDim.x) {
    int64_t g = hist_items[i] - id_offset;  // global id -> row of this shard (if it lives here)
    if (g < 0 || g % id_stride != 0) continue;
    int64_t it = g / id_stride;
    if (it < N) fval[(int64_t)b * ld + it] = split_mode == 1 ? 0.f : -INFINITY;
  }
}

// ---- pass 2: per-user radix select + sort ---------------------------------------------------------
#define SEL_THREADS 1024
#define SEL_MAXK 1024
#define SEL_BINS 2048      // first pass: top 11 bits of the order key
#define SEL_CAND2 4096     // capacity for the elements of the threshold bin

// visits every element of the row once; 16-byte loads when the row allows it
template <typename F>
__device__ __forceinline__ void sel_for_each(const float* __restrict__ row, int64_t N, F f) {
  const int tid = threadIdx.x;
  if ((((uintptr_t)row) & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int64_t i = tid; i < (N >> 2); i += SEL_THREADS) {
      float4 v = __ldg(r4 + i);
      f(v.x, i * 4); f(v.y, i * 4 + 1); f(v.z, i * 4 + 2); f(v.w, i * 4 + 3);
    }
    for (int64_t i = (N & ~(int64_t)3) + tid; i < N; i += SEL_THREADS) f(row[i], i);
  } else {
    for (int64_t i = tid; i < N; i += SEL_THREADS) f(row[i], i);
  }
}

// block-wide bitonic sort (ascending) of n_pad (power of two) 64-bit keys in shared memory
__device__ __forceinline__ void sel_bitonic(unsigned long long* a, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += SEL_THREADS) {
        int j = i ^ stride;
        if (j > i) {
          bool up = (i & size) == 0;
          unsigned long long x = a[i], y = a[j];
          if ((x > y) == up) { a[i] = y; a[j] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(SEL_THREADS)
select_topk_kernel(const float* __restrict__ fval, const uint8_t* __restrict__ fhead, int64_t N, int64_t ld, int K,
                   int64_t id_offset, int64_t id_stride, int64_t* __restrict__ topk_idx, float* __restrict__ topk_val, int32_t* __restrict__ topk_head) {
  // cand2 holds the elements of the threshold bin; the 11-bit histogram of the first pass aliases it
  __shared__ unsigned long long cand2[SEL_CAND2];
  __shared__ unsigned long long cand[SEL_MAXK];
  __shared__ unsigned int s_prefix, s_remaining, s_count, s_count2, s_scan[SEL_THREADS / 32 + 1];
  unsigned int* hist = reinterpret_cast<unsigned int*>(cand2);
  const int b = blockIdx.x;
  const float* row = fval + (int64_t)b * ld;
  const int tid = threadIdx.x;

  // ---- read 1: 11-bit histogram -> threshold bin d (the bin holding the K-th largest), rem = how many of its
  //      elements belong to the top-K
  for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
  __syncthreads();
  sel_for_each(row, N, [&](float v, int64_t) { atomicAdd(&hist[order_key(v) >> 21], 1u); });
  __syncthreads();
  if (tid == 0) {
    uint32_t rem = (uint32_t)K, d = SEL_BINS - 1;
    for (;; --d) {
      uint32_t c = hist[d];
      if (c >= rem || d == 0) break;
      rem -= c;
    }
    s_prefix = d;
    s_remaining = rem;
    s_count2 = hist[d];
    s_count = 0;
  }
  __syncthreads();
  const uint32_t bin = s_prefix, rem_bin = s_remaining, n_bin = s_count2;
  __syncthreads();   // hist (aliasing cand2) is dead from here

  if (n_bin <= SEL_CAND2) {
    // ---- read 2: elements above the bin are in the top-K; elements of the bin go to cand2
    if (tid == 0) s_count2 = 0;
    __syncthreads();
    sel_for_each(row, N, [&](float v, int64_t i) {
      const uint32_t k = order_key(v), top = k >> 21;
      if (top > bin) {
        unsigned int slot = atomicAdd(&s_count, 1u);
        cand[slot] = ((unsigned long long)(~k) << 32) | (unsigned long long)(uint32_t)i;
      } else if (top == bin) {
        unsigned int slot = atomicAdd(&s_count2, 1u);
        cand2[slot] = ((unsigned long long)(~k) << 32) | (unsigned long long)(uint32_t)i;
      }
    });
    __syncthreads();
    const uint32_t base = s_count;
    // order the bin by (value desc, id asc) and take its first rem_bin elements
    int n2 = 1;
    while (n2 < (int)n_bin) n2 <<= 1;
    for (int i = n_bin + tid; i < n2; i += SEL_THREADS) cand2[i] = 0xffffffffffffffffull;
    __syncthreads();
    sel_bitonic(cand2, n2);
    for (uint32_t i = tid; i < rem_bin && base + i < (uint32_t)K; i += SEL_THREADS) cand[base + i] = cand2[i];
    __syncthreads();
  } else {
    // ---- crowded threshold bin (heavily tied / quantised scores): exact 4 x 8-bit MSD radix passes
    uint32_t prefix = 0, mask = 0;
    uint32_t remaining = (uint32_t)K;
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      for (int64_t i = tid; i < N; i += SEL_THREADS) {
        uint32_t k = order_key(row[i]);
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t rem = remaining, d = 255;
        for (;; --d) {  // walk digits from the largest
          uint32_t c = hist[d];
          if (c >= rem || d == 0) break;
          rem -= c;
        }
        s_prefix = prefix | (d << shift);
        s_remaining = rem;
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      mask |= 255u << shift;
      __syncthreads();
    }
    const uint32_t kth = prefix;       // key of the K-th largest
    const uint32_t quota = remaining;  // how many elements equal to kth belong to the top-K
    // collect strictly-greater elements (unordered) ...
    for (int64_t i = tid; i < N; i += SEL_THREADS) {
      uint32_t k = order_key(row[i]);
      if (k > kth) {
        unsigned int slot = atomicAdd(&s_count, 1u);
        if (slot < SEL_MAXK) cand[slot] = ((unsigned long long)(~k) << 32) | (unsigned long long)(uint32_t)i;
      }
    }
    __syncthreads();
    // ... then the `quota` smallest ids among the ties at kth (ordered compaction, id ascending)
    uint32_t base = s_count;
    uint32_t taken = 0;
    for (int64_t i0 = 0; i0 < N && taken < quota; i0 += SEL_THREADS) {
      int64_t i = i0 + tid;
      uint32_t flag = (i < N && order_key(row[i]) == kth) ? 1u : 0u;
      uint32_t ball = __ballot_sync(0xffffffffu, flag);
      int lane = tid & 31, w = tid >> 5;
      if (lane == 0) s_scan[w] = __popc(ball);
      __syncthreads();
      if (tid == 0) {
        uint32_t run = 0;
        for (int x = 0; x < SEL_THREADS / 32; ++x) {
          uint32_t c = s_scan[x];
          s_scan[x] = run;
          run += c;
        }
        s_scan[SEL_THREADS / 32] = run;
      }
      __syncthreads();
      uint32_t rank = taken + s_scan[w] + __popc(ball & ((1u << lane) - 1u));
      if (flag && rank < quota) cand[base + rank] = ((unsigned long long)(~kth) << 32) | (unsigned long long)(uint32_t)i;
      taken += s_scan[SEL_THREADS / 32];
      __syncthreads();
    }
  }
  // pad and bitonic sort ascending on (~key, id): value desc, id asc
  int n_pad = 1;
  while (n_pad < K) n_pad <<= 1;
  for (int i = tid; i < n_pad; i += SEL_THREADS)
    if (i >= K) cand[i] = 0xffffffffffffffffull;
  __syncthreads();
  sel_bitonic(cand, n_pad);
  for (int i = tid; i < K; i += SEL_THREADS) {
    unsigned long long c = cand[i];
    uint32_t id = (uint32_t)(c & 0xffffffffull);
    topk_idx[(int64_t)b * K + i] = (int64_t)id * id_stride + id_offset;
    topk_val[(int64_t)b * K + i] = row[id];
    topk_head[(int64_t)b * K + i] = fhead[(int64_t)b * ld + id];
  }
}

size_t b200rec_topk_workspace_bytes(int B, int64_t N) {
  size_t f = ((size_t)B * N * 4 + 255) & ~(size_t)255;
  size_t h = ((size_t)B * N + 255) & ~(size_t)255;
  return f + h;
}

int b200rec_score_mask_topk(const float* scores, int64_t ld_scores, int B, int H, int64_t N, int K,
                            const int32_t* head_cat, const uint32_t* item_tag_bits, const uint8_t* head_on,
                            const int32_t* hist_off, const int64_t* hist_items, int split_mode, int64_t id_offset,
                            int64_t id_stride, int64_t* topk_idx, float* topk_val, int32_t* topk_head,
                            void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(H >= 1 && H <= 64, "score_mask_topk: H=%d not in [1,64]", H);
  B200_CHECK_ARG(K >= 1 && K <= SEL_MAXK && K <= N, "score_mask_topk: K=%d not in [1,%d] or > N", K, SEL_MAXK);
  B200_CHECK_ARG(N < (1ll << 32), "score_mask_topk: N too large");
  B200_CHECK_ARG(id_stride >= 1 && id_offset >= 0, "score_mask_topk: bad id mapping");
  B200_CHECK_ARG(workspace_bytes >= b200rec_topk_workspace_bytes(B, N), "score_mask_topk: workspace too small");
  if (B == 0) return 0;
  float* fval = (float*)workspace;
  uint8_t* fhead = (uint8_t*)workspace + (((size_t)B * N * 4 + 255) & ~(size_t)255);
  dim3 grid((unsigned)std::min<int64_t>((N + 255) / 256, 148 * 8), B);
  fold_heads_kernel<<<grid, 256, 0, st>>>(scores, ld_scores, H, N, head_cat, item_tag_bits, head_on, split_mode,
                                          id_offset, id_stride, fval, fhead);
  if (hist_off && hist_items) suppress_history_kernel<<<B, 128, 0, st>>>(hist_off, hist_items, B, N, N, fval, split_mode, id_offset, id_stride);
  select_topk_kernel<<<B, SEL_THREADS, 0, st>>>(fval, fhead, N, N, K, id_offset, id_stride, topk_idx, topk_val, topk_head);
  B200_LAUNCH_OK();
  return 0;
}

int b200rec_topk_select(float* fval, const uint8_t* fhead, int B, int64_t N, int64_t ld, int K, const int32_t* hist_off,
                        const int64_t* hist_items, int64_t id_offset, int64_t id_stride, int64_t* topk_idx,
                        float* topk_val, int32_t* topk_head, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(K >= 1 && K <= SEL_MAXK && K <= N, "topk_select: K=%d not in [1,%d] or > N", K, SEL_MAXK);
  B200_CHECK_ARG(N < (1ll << 32) && id_stride >= 1 && id_offset >= 0 && ld >= N, "topk_select: bad N / ld / id mapping");
  if (B == 0) return 0;
  if (hist_off && hist_items) suppress_history_kernel<<<B, 128, 0, st>>>(hist_off, hist_items, B, N, ld, fval, 0, id_offset, id_stride);
  select_topk_kernel<<<B, SEL_THREADS, 0, st>>>(fval, fhead, N, ld, K, id_offset, id_stride, topk_idx, topk_val, topk_head);
  B200_LAUNCH_OK();
  return 0;
}

// ---- streamed eval: final order of the candidates appended by the scoring GEMM's epilogue ----------------------
// key = (~order_key(score) << 32) | (local item << 5) | head: ascending key order = value desc, item asc, head asc.
// The sort covers the next power of two above THIS user's count, not the capacity.  dedupe (several head groups per
// user): a first sort by (item, value desc, head) puts the copies of an item next to each other, all but the first of
// each run are dropped, the second sort restores the final order.
__global__ void __launch_bounds__(SEL_THREADS)
topk_from_candidates_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ cnt, int cap,
                            int K, int dedupe, const int32_t* __restrict__ hist_off,
                            const int64_t* __restrict__ hist_items, int64_t id_offset, int64_t id_stride,
                            int64_t* __restrict__ topk_idx, float* __restrict__ topk_val,
                            int32_t* __restrict__ topk_head, int32_t* __restrict__ overflow) {
  extern __shared__ unsigned long long sc[];             // [next pow2 >= cap]
  __shared__ int s_valid;
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint32_t n_all = cnt[b];
  const int n = (int)min(n_all, (uint32_t)cap);
  int n_pad = 1;
  while (n_pad < n || n_pad < K) n_pad <<= 1;
  if (tid == 0) {
    s_valid = 0;
    if (n_all > (uint32_t)cap) overflow[0] = 1;
  }
  __syncthreads();
  const unsigned long long* kb = keys + (int64_t)b * cap;
  const int h0 = hist_off ? hist_off[b] : 0, h1 = hist_off ? hist_off[b + 1] : 0;
  for (int i = tid; i < n_pad; i += SEL_THREADS) {
    unsigned long long k = 0xffffffffffffffffull;
    if (i < n) {
      const unsigned long long c = kb[i];
      const int64_t gid = (int64_t)((uint32_t)(c & 0xffffffffull) >> 5) * id_stride + id_offset;
      bool drop = gid == 0;                                // trainer.py:724
      for (int h = h0; h < h1 && !drop; ++h) drop = hist_items[h] == gid;   // trainer.py:725-726
      if (!drop) {
        // dedupe: first order by (item, ~value key, head) = swap the two 27 / 32-bit fields
        k = dedupe ? ((unsigned long long)((uint32_t)(c & 0xffffffffull) >> 5) << 37) | ((c >> 32) << 5) | (c & 31ull) : c;
      }
    }
    sc[i] = k;
  }
  __syncthreads();
  if (dedupe) {
    sel_bitonic(sc, n_pad);
    // keep the first (= best) entry of each item, convert back to the final key
    unsigned long long mine[8];                            // n_pad <= 16384 = 16 x 1024: two rounds of 8 would be needed
    for (int base = 0; base < n_pad; base += 8 * SEL_THREADS) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * SEL_THREADS + tid;
        unsigned long long k = 0xffffffffffffffffull;
        if (i < n_pad) {
          const unsigned long long c = sc[i];
          if (c != 0xffffffffffffffffull) {
            const bool first = i == 0 || (sc[i - 1] >> 37) != (c >> 37);
            if (first) k = (((c >> 5) & 0xffffffffull) << 32) | ((c >> 37) << 5) | (c & 31ull);
          }
        }
        mine[u] = k;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * SEL_THREADS + tid;
        if (i < n_pad) sc[i] = mine[u];
      }
      __syncthreads();
    }
  }
  int cntv = 0;
  for (int i = tid; i < n_pad; i += SEL_THREADS) cntv += sc[i] != 0xffffffffffffffffull;
  if (cntv) atomicAdd(&s_valid, cntv);
  sel_bitonic(sc, n_pad);
  if (tid == 0 && s_valid < K) overflow[0] = 1;
  for (int i = tid; i < K; i += SEL_THREADS) {
    const unsigned long long c = i < n_pad ? sc[i] : 0xffffffffffffffffull;
    const bool ok = c != 0xffffffffffffffffull;
    const uint32_t low = (uint32_t)(c & 0xffffffffull);
    uint32_t u = ~(uint32_t)(c >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;        // inverse of order_key
    topk_idx[(int64_t)b * K + i] = ok ? (int64_t)(low >> 5) * id_stride + id_offset : 0;
    topk_val[(int64_t)b * K + i] = ok ? __uint_as_float(u) : -INFINITY;
    topk_head[(int64_t)b * K + i] = ok ? (int)(low & 31u) : 0;
  }
}

int b200rec_topk_from_candidates(const uint64_t* keys, const uint32_t* cnt, int cap, int B, int K, int dedupe,
                                 const int32_t* hist_off, const int64_t* hist_items, int64_t id_offset,
                                 int64_t id_stride, int64_t* topk_idx, float* topk_val, int32_t* topk_head,
                                 int32_t* overflow, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK_ARG(K >= 1 && K <= cap && cap <= 16384 && id_stride >= 1 && id_offset >= 0,
                 "topk_from_candidates: K=%d / cap=%d (<= 16384) / id mapping", K, cap);
  if (B == 0) return 0;
  int n_pad = 1;
  while (n_pad < cap) n_pad <<= 1;
  const size_t smem = (size_t)n_pad * 8;
  static bool once = false;
  if (!once) {
    B200_CUDA_OK(cudaFuncSetAttribute(topk_from_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    once = true;
  }
  topk_from_candidates_kernel<<<B, SEL_THREADS, smem, st>>>((const unsigned long long*)keys, cnt, cap, K, dedupe, hist_off,
                                                             hist_items, id_offset, id_stride, topk_idx, topk_val,
                                                             topk_head, overflow);
  B200_LAUNCH_OK();
  return 0;
}

// ---- hit matrix (collector.py:300-316) -----------------------------------------------------------
__global__ void hit_matrix_kernel(const int64_t* __restrict__ topk_idx, const int64_t* __restrict__ positive_i, int B,
                                  int K, int Pe, int p, int32_t* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ int64_t tgt[64];
  __shared__ int pos_len;
  if (threadIdx.x == 0) {
    // pos_len quirk: #distinct ids among the p+1 smallest-id targets of the FULL row
    int64_t s[64];
    for (int i = 0; i < Pe; ++i) s[i] = positive_i[(int64_t)b * Pe + i];
    for (int i = 1; i < Pe; ++i) {
      int64_t v = s[i];
      int j = i - 1;
      while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
      s[j + 1] = v;
    }
    int c = 0;
    for (int i = 0; i <= p; ++i) c += (i == 0 || s[i] != s[i - 1]);
    pos_len = c;
    for (int i = 0; i <= p; ++i) tgt[i] = positive_i[(int64_t)b * Pe + i];  // cumulative slice [0, p]
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int64_t id = topk_idx[(int64_t)b * K + k];
    int hit = 0;
    for (int i = 0; i <= p; ++i) hit |= (tgt[i] == id);
    out[(int64_t)b * (K + 1) + k] = hit;
  }
  if (threadIdx.x == 0) out[(int64_t)b * (K + 1) + K] = pos_len;
}

int b200rec_hit_matrix(const int64_t* topk_idx, const int64_t* positive_i, int B, int K, int Pe,
                       const int32_t* pred_list_host, int n_p, int32_t* out, void* stream) {
  B200_CHECK_ARG(Pe >= 1 && Pe <= 64, "hit_matrix: eval_pred_len %d not in [1,64]", Pe);
  if (B == 0) return 0;
  for (int q = 0; q < n_p; ++q) {
    int p = pred_list_host[q];
    B200_CHECK_ARG(p >= 0 && p < Pe, "hit_matrix: pred index %d out of range", p);
    hit_matrix_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(topk_idx, positive_i, B, K, Pe, p,
                                                           out + (int64_t)q * B * (K + 1));
  }
  B200_LAUNCH_OK();
  return 0;
}

// ---- in-place masks for the reference-compatible predict() that returns [B, H, N] scores -------------
// hstu.py:983-999: heads switched off per user (prior_given_at_test) and items outside the head's category.
__global__ void __launch_bounds__(256)
apply_score_masks_kernel(float* __restrict__ scores, int64_t ld, int H, int64_t N, const int32_t* __restrict__ head_cat,
                         const uint32_t* __restrict__ item_tag_bits, const uint8_t* __restrict__ head_on) {
  const int row = blockIdx.y;  // b*H + h
  const int h = row % H;
  const int cat = head_cat ? head_cat[h] : -1;
  const bool on = head_on ? head_on[row] != 0 : true;
  if (on && cat < 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    bool keep = on && ((item_tag_bits[i] >> cat) & 1u);
    if (!keep) scores[(int64_t)row * ld + i] = -INFINITY;
  }
}

extern "C" int b200rec_apply_score_masks(float* scores, int64_t ld_scores, int B, int H, int64_t N,
                                         const int32_t* head_cat, const uint32_t* item_tag_bits,
                                         const uint8_t* head_on, void* stream) {
  if (B == 0) return 0;
  dim3 grid((unsigned)std::min<int64_t>((N + 255) / 256, 148 * 4), B * H);
  apply_score_masks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, ld_scores, H, N, head_cat, item_tag_bits,
                                                                   head_on);
  B200_LAUNCH_OK();
  return 0;
}
