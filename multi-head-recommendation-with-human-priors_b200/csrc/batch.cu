// GPU-side batch construction: what the reference does per sample in 8 Python DataLoader
// workers (data/dataset/trainset.py:70-177, evalset.py:81-155, collate_fn.py:59-90) as one kernel per batch.
//
// Train row (trainset.py:155-177): window [pad x (L - ctx) | user_seq[start : end + pred] | pad x (P - pred)],
// pads = random items outside the row's own items (pad_random_sample) or 0, mask = 1 on real positions,
// negatives (trainset.py:70-97,126-137): for every category pool c (neg_sample_by_cat) and then the global pool,
// n DISTINCT items drawn uniformly from the pool, none of them in the (padded) row; tags = item_tags[item]
// (category_by item) or the one-hot event type at real positions (category_by event).
//
// Randomness is a counter-based Philox4x32-10 stream keyed on (seed, step, row, slot): reproducible, independent
// of the launch shape, and necessarily a different stream from the reference's numpy generators — parity is by
// the sampling LAW (uniform without replacement outside the blacklist), checked property by property.
#include "common.cuh"

namespace {

struct Rng {
  uint32_t seed_lo, seed_hi, step_lo, step_hi;
  // 32 random bits for (row, slot, draw)
  __device__ __forceinline__ uint32_t bits(uint32_t row, uint32_t slot, uint32_t draw) const {
    uint4 r = philox4x32_10(make_uint4(row, slot, draw >> 2, step_lo), make_uint2(seed_lo, seed_hi ^ step_hi));
    const uint32_t k = draw & 3u;
    return k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w));
  }
  // uniform integer in [0, n): 64-bit multiply-high of two 32-bit draws (bias < 2^-32 * n)
  __device__ __forceinline__ int64_t below(uint32_t row, uint32_t slot, uint32_t draw, int64_t n) const {
    const uint64_t x = ((uint64_t)bits(row, slot, 2 * draw) << 32) | bits(row, slot, 2 * draw + 1);
    return (int64_t)__umul64hi(x, (uint64_t)n);
  }
};

#define BATCH_MAX_LP 1024   // longest window held in shared memory

__global__ void __launch_bounds__(128) build_train_batch_kernel(
    const int64_t* __restrict__ user_seq, const int64_t* __restrict__ user_off, const int32_t* __restrict__ train_len,
    const int32_t* __restrict__ event_seq, const int64_t* __restrict__ sample_uid,
    const int32_t* __restrict__ sample_end, const int64_t* __restrict__ batch_index, int L, int P, int pad_random,
    int64_t item_num, int n_sets, int n_neg, const int64_t* __restrict__ cat_items,
    const int64_t* __restrict__ cat_off, int n_pools, float mix_ratio, const uint8_t* __restrict__ item_tags, int C,
    Rng rng, int64_t* __restrict__ items, int64_t* __restrict__ neg, int64_t* __restrict__ mask,
    int64_t* __restrict__ tags) {
  __shared__ int64_t row[BATCH_MAX_LP];
  extern __shared__ int64_t acc_all[];            // [warps][n_neg] accepted negatives of the set a warp works on
  const int b = blockIdx.x, LP = L + P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const int64_t smp = batch_index[b];
  const int64_t uid = sample_uid[smp];
  const int end = sample_end[smp];
  const int start = max(0, end - L);
  const int ctx_pad = L - (end - start);
  const int pred = min(train_len[uid] - end, P);
  const int n_real = (end - start) + pred;
  const int64_t* seq = user_seq + user_off[uid] + start;
  // ---- real items, mask
  for (int i = tid; i < LP; i += blockDim.x) {
    const int j = i - ctx_pad;
    const bool real = j >= 0 && j < n_real;
    row[i] = real ? seq[j] : 0;
    mask[(int64_t)b * LP + i] = real ? 1 : 0;
  }
  __syncthreads();
  // ---- pads: distinct random items outside the row's real items (one warp, position by position)
  if (pad_random && warp == 0) {
    uint32_t draw = 0;
    for (int i = 0; i < LP; ++i) {
      const int j = i - ctx_pad;
      if (j >= 0 && j < n_real) continue;                       // warp-uniform
      for (int tries = 0;; ++tries) {
        const int64_t cand = 1 + rng.below(b, 0, draw++, item_num - 1);
        bool clash = false;
        for (int q = lane; q < LP; q += 32) {
          const int jq = q - ctx_pad;
          const bool filled = (jq >= 0 && jq < n_real) || q < i;  // real items and the pads already placed
          clash |= filled && row[q] == cand;
        }
        if (!__any_sync(0xffffffffu, clash) || tries >= 64) {    // (a catalogue smaller than the window cannot comply)
          if (lane == 0) row[i] = cand;
          break;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < LP; i += blockDim.x) items[(int64_t)b * LP + i] = row[i];
  // ---- tags
  if (tags != nullptr && C > 0) {
    for (int i = tid; i < LP * C; i += blockDim.x) {
      const int pos = i / C, c = i - pos * C;
      int64_t v;
      if (item_tags != nullptr) {
        v = item_tags[row[pos] * C + c];                         // trainset.py:165-167: pads included
      } else {
        const int j = pos - ctx_pad;
        v = (j >= 0 && j < n_real && event_seq != nullptr) ? (event_seq[user_off[uid] + start + j] == c) : 0;
      }
      tags[((int64_t)b * LP + pos) * C + c] = v;
    }
  }
  // ---- negatives: one warp per set, 32 candidates per round
  for (int s = warp; s < n_sets; s += n_warps) {
    int64_t* acc = acc_all + (int64_t)warp * n_neg;
    const int64_t* pool = nullptr;
    int64_t pool_n = item_num - 1;                               // global pool = ids 1 .. item_num-1
    if (s < n_pools) {
      const bool use_cat = mix_ratio <= 0.f || (rng.bits(b, 1 + s, 0xfffffff0u) * 2.3283064365386963e-10f) > mix_ratio;
      if (use_cat) {
        pool = cat_items + cat_off[s];
        pool_n = cat_off[s + 1] - cat_off[s];
      }
    }
    int n_acc = 0;
    uint32_t draw = 0;
    // a pool smaller than n_neg + blacklist cannot deliver n distinct items: fall back to "with replacement"
    const bool distinct = pool_n >= (int64_t)n_neg + LP;
    for (int rounds = 0; n_acc < n_neg; ++rounds) {
      const int64_t r = rng.below(b, 1 + s, draw + lane, pool_n);
      draw += 32;
      const int64_t cand = pool ? pool[r] : 1 + r;
      bool ok = true;
      for (int q = 0; q < LP && ok; ++q) ok = row[q] != cand;    // blacklist = the padded row (trainset.py:128-133)
      if (rounds >= 256) ok = true;                              // pool (almost) inside the blacklist: give up filtering
      if (distinct) {
        for (int q = 0; q < n_acc && ok; ++q) ok = acc[q] != cand;
        for (int o = 0; o < 32; ++o) {                           // first occurrence inside this round wins
          const int64_t other = __shfl_sync(0xffffffffu, cand, o);
          const bool other_ok = __shfl_sync(0xffffffffu, (int)ok, o) != 0;
          if (o < lane && other_ok && other == cand) ok = false;
        }
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, ok);
      const int slot = n_acc + __popc(ball & ((1u << lane) - 1u));
      if (ok && slot < n_neg) {
        acc[slot] = cand;
        neg[((int64_t)b * n_sets + s) * n_neg + slot] = cand;
      }
      n_acc = min(n_neg, n_acc + __popc(ball));
      __syncwarp();
    }
  }
}

// Eval rows (evalset.py:81-155 + collate_fn.py:59-90).  phase 0 = valid: history = seq[:train_len], targets the
// next Pe items; phase 1 = test: history = seq[:-Pe], targets the last Pe.  item_seq = last L history items,
// left-padded with 0; history pairs (u = row, i = every history item) land at hist_off[row] (host prefix sum).
__global__ void __launch_bounds__(128) build_eval_batch_kernel(
    const int64_t* __restrict__ user_seq, const int64_t* __restrict__ user_off, const int32_t* __restrict__ train_len,
    const int32_t* __restrict__ event_seq, const int64_t* __restrict__ uids, int L, int Pe, int phase,
    const uint8_t* __restrict__ item_tags, int C, const int64_t* __restrict__ hist_off, int64_t* __restrict__ item_seq,
    int64_t* __restrict__ item_target, int64_t* __restrict__ target_tags, int64_t* __restrict__ hist_u,
    int64_t* __restrict__ hist_i) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t uid = uids[b];
  const int64_t base = user_off[uid];
  const int total = (int)(user_off[uid + 1] - base);
  const int n_hist = phase == 0 ? train_len[uid] : total - Pe;
  for (int i = tid; i < L; i += blockDim.x) {
    const int j = n_hist - L + i;                                // right-aligned window
    item_seq[(int64_t)b * L + i] = j >= 0 ? user_seq[base + j] : 0;
  }
  for (int i = tid; i < Pe; i += blockDim.x) {
    const int j = n_hist + i;
    const int64_t it = j < total ? user_seq[base + j] : 0;
    item_target[(int64_t)b * Pe + i] = it;
    for (int c = 0; c < C; ++c) {
      int64_t v = 0;
      if (item_tags != nullptr) v = item_tags[it * C + c];
      else if (event_seq != nullptr && j < total) v = event_seq[base + j] == c;
      if (target_tags != nullptr) target_tags[((int64_t)b * Pe + i) * C + c] = v;
    }
  }
  const int64_t h0 = hist_off[b];
  for (int i = tid; i < n_hist; i += blockDim.x) {
    hist_u[h0 + i] = b;
    hist_i[h0 + i] = user_seq[base + i];
  }
}

}  // namespace

extern "C" int b200rec_build_train_batch(const int64_t* user_seq, const int64_t* user_off, const int32_t* train_len,
                                         const int32_t* event_seq, const int64_t* sample_uid,
                                         const int32_t* sample_end, const int64_t* batch_index, int B, int L, int P,
                                         int pad_random, int64_t item_num, int n_sets, int n_neg,
                                         const int64_t* cat_items, const int64_t* cat_off, int n_pools,
                                         float neg_sample_mix_ratio, const uint8_t* item_tags, int C, uint64_t seed,
                                         uint64_t step, int64_t* items, int64_t* neg_items, int64_t* mask,
                                         int64_t* tags, void* stream) {
  B200_CHECK_ARG(L + P <= BATCH_MAX_LP, "build_train_batch: L+P=%d exceeds %d", L + P, BATCH_MAX_LP);
  B200_CHECK_ARG(n_pools >= 0 && n_pools <= n_sets && n_neg >= 0 && item_num >= 2, "build_train_batch: bad sizes");
  B200_CHECK_ARG(n_pools == 0 || (cat_items != nullptr && cat_off != nullptr), "build_train_batch: category pools missing");
  B200_CHECK_ARG((size_t)4 * n_neg * 8 <= 160 * 1024, "build_train_batch: n_neg=%d too large", n_neg);
  if (B == 0) return 0;
  Rng rng = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
  const size_t smem = (size_t)4 * std::max(n_neg, 1) * sizeof(int64_t);
  if (smem > 40 * 1024)
    B200_CUDA_OK(cudaFuncSetAttribute(build_train_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  build_train_batch_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(
      user_seq, user_off, train_len, event_seq, sample_uid, sample_end, batch_index, L, P, pad_random, item_num, n_sets,
      n_neg, cat_items, cat_off, n_pools, neg_sample_mix_ratio, item_tags, C, rng, items, neg_items, mask, tags);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_build_eval_batch(const int64_t* user_seq, const int64_t* user_off, const int32_t* train_len,
                                        const int32_t* event_seq, const int64_t* uids, int B, int L, int Pe, int phase,
                                        const uint8_t* item_tags, int C, const int64_t* hist_off, int64_t* item_seq,
                                        int64_t* item_target, int64_t* target_tags, int64_t* hist_u, int64_t* hist_i,
                                        void* stream) {
  B200_CHECK_ARG(phase == 0 || phase == 1, "build_eval_batch: phase must be 0 (valid) or 1 (test)");
  if (B == 0) return 0;
  build_eval_batch_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(user_seq, user_off, train_len, event_seq, uids, L, Pe,
                                                               phase, item_tags, C, hist_off, item_seq, item_target,
                                                               target_tags, hist_u, hist_i);
  B200_LAUNCH_OK();
  return 0;
}
