// Prior construction on the GPU: the two graph steps of the reference's clustering scripts.
//
//  (1) co-occurrence edges (item-clustering.py:152-162, user-clustering.py:218-233 / 268-290): every group (a user's
//      training window of items, or the users of one item) contributes all unordered pairs of its DISTINCT members; the
//      graph is the set of distinct pairs.  The reference builds a Python set of itertools.combinations; here a group's
//      members arrive sorted and de-duplicated (one 64-bit radix sort of (group, member) keys on the host side), one
//      CTA per group emits its n(n-1)/2 pairs as 64-bit keys (a << 32 | b, a < b) with coalesced 8-byte stores, and a
//      second radix sort + unique gives the edge set.  HBM-bound integer work: 8 B written per pair, 16 B per pair of
//      sort traffic per pass.
//  (2) community detection: the reference calls igraph's Leiden (community_leiden, objective "modularity",
//      item-clustering.py:227-250).  igraph is a third-party dependency that is not part of the reference tree; what is
//      built here is the modularity local-moving phase that Louvain and Leiden share, in a deterministic synchronous
//      form (below), applied level by level with graph contraction on the host side.  All weights are integers, so the
//      community totals are exact and the run is bit-reproducible; gains are compared in IEEE double without
//      contraction so that the CPU oracle (oracle/graph_oracle.py) reproduces every decision.
//
// Local moving, one sweep: for node i with degree k_i in community A, every neighbouring community C (and A itself) is
// scored  s(C) = w(i -> C) - gamma * k_i * (tot(C) - [C == A] k_i) / 2m ; the node proposes the best-scoring community
// (ties -> smaller community id) if it beats staying strictly.  Only nodes whose parity matches the sweep move (odd /
// even sweeps alternate), and a node alone in its community never moves to another singleton with a larger id (the
// two classic guards against oscillation of synchronous moves).
#include "common.cuh"

__global__ void __launch_bounds__(256) group_pairs_count_kernel(const int64_t* __restrict__ group_off, int64_t G, int cap,
                                                                int64_t* __restrict__ counts) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  int64_t n = group_off[g + 1] - group_off[g];
  if (cap > 0 && n > cap) n = cap;
  counts[g] = n * (n - 1) / 2;
}

// one CTA per group; row i of the upper triangle starts at i * n - i * (i + 1) / 2
__global__ void __launch_bounds__(256) group_pairs_emit_kernel(const int32_t* __restrict__ members,
                                                               const int64_t* __restrict__ group_off,
                                                               const int64_t* __restrict__ pair_off, int64_t G, int cap,
                                                               unsigned long long* __restrict__ keys) {
  for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
    const int64_t beg = group_off[g];
    int64_t n = group_off[g + 1] - beg;
    if (cap > 0 && n > cap) n = cap;
    if (n < 2) continue;
    unsigned long long* out = keys + pair_off[g];
    const int32_t* m = members + beg;
    if (n <= 64) {
      // small group: flat index over the pairs, (i, j) decoded by walking the rows (n <= 64: at most 63 steps)
      const int64_t np = n * (n - 1) / 2;
      for (int64_t p = threadIdx.x; p < np; p += blockDim.x) {
        int i = 0;
        int64_t rem = p;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        const int j = i + 1 + (int)rem;
        out[p] = ((unsigned long long)(uint32_t)m[i] << 32) | (unsigned long long)(uint32_t)m[j];
      }
    } else {
      for (int64_t i = 0; i + 1 < n; ++i) {
        const unsigned long long hi = (unsigned long long)(uint32_t)m[i] << 32;
        unsigned long long* row = out + (i * n - i * (i + 1) / 2) - (i + 1);     // row[j] for j in (i, n)
        for (int64_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x) row[j] = hi | (unsigned long long)(uint32_t)m[j];
      }
    }
  }
}

extern "C" int b200rec_group_pairs_count(const int64_t* group_off, int64_t G, int cap, int64_t* counts, void* stream) {
  if (G == 0) return 0;
  group_pairs_count_kernel<<<(unsigned)ceil_div_i(G, 256), 256, 0, (cudaStream_t)stream>>>(group_off, G, cap, counts);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_group_pairs_emit(const int32_t* members, const int64_t* group_off, const int64_t* pair_off,
                                        int64_t G, int cap, uint64_t* keys, void* stream) {
  if (G == 0) return 0;
  const int blocks = (int)std::min<int64_t>(G, 148 * 16);
  group_pairs_emit_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(members, group_off, pair_off, G, cap,
                                                                   (unsigned long long*)keys);
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- local moving
// Input: for every node the list of (neighbouring community, summed edge weight) runs, sorted by community id
// (node_off[i] .. node_off[i+1]); self loops are not part of the lists (they only count in deg).  One thread per node.
__global__ void __launch_bounds__(256) louvain_best_move_kernel(const int64_t* __restrict__ node_off,
                                                                const int32_t* __restrict__ run_comm,
                                                                const int64_t* __restrict__ run_w,
                                                                const int32_t* __restrict__ comm,
                                                                const int64_t* __restrict__ deg,
                                                                const int64_t* __restrict__ tot,
                                                                const int32_t* __restrict__ csize, int64_t n,
                                                                int64_t two_m, double gamma, int parity,
                                                                int32_t* __restrict__ best) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t own = comm[i];
  int32_t choice = own;
  if ((int)(i & 1) == parity) {
    const double ki = (double)deg[i];
    const double scale = __ddiv_rn(__dmul_rn(gamma, ki), (double)two_m);        // gamma * k_i / 2m
    // staying: edges into the own community (the run may be absent -> 0)
    double stay_w = 0.0;
    for (int64_t r = node_off[i]; r < node_off[i + 1]; ++r)
      if (run_comm[r] == own) stay_w = (double)run_w[r];
    const double stay = __dsub_rn(stay_w, __dmul_rn(scale, (double)(tot[own] - deg[i])));
    double best_s = stay;
    for (int64_t r = node_off[i]; r < node_off[i + 1]; ++r) {
      const int32_t c = run_comm[r];
      if (c == own) continue;
      if (csize[own] == 1 && csize[c] == 1 && c > own) continue;               // singleton swap guard
      const double s = __dsub_rn((double)run_w[r], __dmul_rn(scale, (double)tot[c]));
      if (s > best_s) { best_s = s; choice = c; }                              // runs ascend in c: ties keep the smaller id
    }
  }
  best[i] = choice;
}

__global__ void __launch_bounds__(256) louvain_apply_kernel(const int32_t* __restrict__ best, int32_t* __restrict__ comm,
                                                            const int64_t* __restrict__ deg, int64_t* __restrict__ tot,
                                                            int32_t* __restrict__ csize, int64_t n,
                                                            unsigned long long* __restrict__ moved) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t a = comm[i], b = best[i];
  if (a == b) return;
  comm[i] = b;
  // integer atomics: the totals are exact whatever the order
  atomicAdd((unsigned long long*)(tot + a), (unsigned long long)(-deg[i]));
  atomicAdd((unsigned long long*)(tot + b), (unsigned long long)deg[i]);
  atomicSub(csize + a, 1);
  atomicAdd(csize + b, 1);
  atomicAdd(moved, 1ull);
}

extern "C" int b200rec_louvain_best_move(const int64_t* node_off, const int32_t* run_comm, const int64_t* run_w,
                                         const int32_t* comm, const int64_t* deg, const int64_t* tot,
                                         const int32_t* csize, int64_t n, int64_t two_m, double gamma, int parity,
                                         int32_t* best, void* stream) {
  if (n == 0) return 0;
  B200_CHECK_ARG(two_m > 0, "louvain: graph has no edges");
  louvain_best_move_kernel<<<(unsigned)ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(
      node_off, run_comm, run_w, comm, deg, tot, csize, n, two_m, gamma, parity, best);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_louvain_apply(const int32_t* best, int32_t* comm, const int64_t* deg, int64_t* tot, int32_t* csize,
                                     int64_t n, uint64_t* moved, void* stream) {
  if (n == 0) return 0;
  louvain_apply_kernel<<<(unsigned)ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(
      best, comm, deg, tot, csize, n, (unsigned long long*)moved);
  B200_LAUNCH_OK();
  return 0;
}
