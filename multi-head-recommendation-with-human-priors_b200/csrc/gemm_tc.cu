// Persistent warp-specialised bf16 GEMM on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring -> tcgen05.mma kind::f16 (1 thread)
//   -> fp32 accumulators in TMEM (2 stages) -> tcgen05.ld -> fused epilogue -> global.
// C[M,N] = epi(A[M,K] * B[N,K]^T); each operand may be K-major or MN-major in global memory
// (MN-major avoids every explicit transpose in the backward pass: dW = X^T dY, dX = dY W).
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..9 = epilogue (two warps per TMEM lane
// quarter).  Epilogue = TMEM -> registers (lane = row) -> swizzled smem transpose -> lanes along the row
// -> fused math + fully coalesced 16-byte global accesses (8 lanes cover 128 contiguous bytes).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "gemm_epilogue.cuh"
#include "tc_ptx.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_THREADS 320
#define TC_EPI_WARPS 8
#define TC_STAGE_BYTES (TC_EPI_WARPS * 4096)
#define TC_TMEM_COLS 512
#define TC_SMEM_LIMIT (227 * 1024)

struct TcParams {
  int M, N, K;
  int BN;          // 128 or 256
  int stages;
  int a_mn, b_mn;  // 1 = MN-major operand
  int num_m, num_n;
  int split_k, kb_per_split;      // split-K: work item = (tile, k-range); partial tiles go to a workspace
  int64_t split_stride;           // elements between consecutive partial outputs
  int groups;                     // grouped launch: `groups` independent problems of identical shape
  int n_fastest;                  // tile order: consecutive tiles share the A (row) tile instead of the B tile
  // tail split: the tiles of the last, partly filled wave (linear tile index >= tail_first) are cut into tail_split
  // k-ranges of tail_kb k-blocks each, one per otherwise idle unit; their fp32 partial accumulators go to tail_ws
  // ([tile][split][128 * CTAS rows][BN]) and tail_fixup_kernel sums them in split order and applies the epilogue
  int tail_first, tail_split, tail_kb;
  float* tail_ws;
};

// One unit of work of the persistent loop: an output tile and the k-blocks [kb_beg, kb_end) of it.
struct TcWork {
  int gtile, split, kb_beg, kb_end;
  int tail_slot;   // >= 0: partial accumulators of a tail tile -> tail_ws slot
};
__device__ __forceinline__ TcWork tc_decode(const TcParams& p, int item, int num_tiles, int num_k) {
  TcWork w;
  if (p.tail_split > 0 && item >= p.tail_first) {
    const int j = item - p.tail_first;
    const int tt = j / p.tail_split;
    w.split = j - tt * p.tail_split;
    w.gtile = p.tail_first + tt;
    w.kb_beg = w.split * p.tail_kb;
    w.kb_end = min(num_k, w.kb_beg + p.tail_kb);
    w.tail_slot = j;
  } else {
    w.gtile = item % num_tiles;
    w.split = item / num_tiles;
    w.kb_beg = w.split * p.kb_per_split;
    w.kb_end = min(num_k, w.kb_beg + p.kb_per_split);
    w.tail_slot = -1;
  }
  return w;
}

// Grouped launch (one persistent kernel over several same-shape problems, so that small problems share
// waves instead of each paying its own tail): per-group operand maps and output pointers, by value.
#define TC_MAX_GROUPS 16
struct TcGroupArgs {
  CUtensorMap a[TC_MAX_GROUPS];
  CUtensorMap b[TC_MAX_GROUPS];
  void* C[TC_MAX_GROUPS];
  const float* row_scale[TC_MAX_GROUPS];   // STORE / ACCUM (nullable)
  const float* nce_mref[TC_MAX_GROUPS];    // NCE_EXP
  const float* nce_thr[TC_MAX_GROUPS];
  float* nce_stats[TC_MAX_GROUPS];
};

// B200REC_EPI_FOLD_ITEMS, one epilogue warp: lane = item row m, the warp's `cols` accumulator columns starting at column
// nc0 (a multiple of HP) hold whole users of HP heads each.  NCH 32-column chunks are loaded per batch (NCH * 32 % HP == 0)
// and every user's masked max / arg-max is taken over registers of ONE thread: no shuffles, no shared-memory transpose.
template <int HP, int NCH>
__device__ __forceinline__ void fold_items_warp(const EpiParams& ep, int cols, uint32_t t_addr, int m, int nc0) {
  constexpr int UPB = NCH * 32 / HP;                       // users per batch
  const int n_users = ep.N / HP;
  const int64_t id = (int64_t)m * ep.fold_id_stride + ep.fold_id_offset;
  const bool row_ok = m < ep.M;
  uint32_t hmask = 0;                                      // bit h: head h may score this item
  if (row_ok && id != 0) {
    const uint32_t tg = ep.fold_item_tags ? __ldg(ep.fold_item_tags + m) : 0xffffffffu;
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      const int cat = ep.fold_head_cat ? __ldg(ep.fold_head_cat + h) : -1;
      hmask |= ((cat < 0 || ((tg >> cat) & 1u)) ? 1u : 0u) << h;
    }
  }
#pragma unroll 1
  for (int b0 = 0; b0 < cols; b0 += NCH * 32) {
    uint32_t rr[NCH][32];
#pragma unroll
    for (int q = 0; q < NCH; ++q) tmem_ld_32x32_nowait(t_addr + (uint32_t)(b0 + q * 32), rr[q]);
    tmem_ld_wait();
    const int user0 = (nc0 + b0) / HP;
    if (ep.fold_thr != nullptr) {
      // streamed eval: only scores that can still enter the user's top-K leave the SM.  Two phases: (A) every user's
      // masked max against its threshold with NO side effects, so the per-user threshold / head-mask loads of all UPB
      // users are in flight together (one loop with the append inside serialised them behind its atomics: the H = 1
      // sweep ran at 0.08 of the HBM rate); (B) the rare survivors append to the candidate lists.
      unsigned long long pass = 0ull;
#pragma unroll
      for (int u = 0; u < UPB; ++u) {
        const int user = user0 + u;
        const bool uok = user < n_users;
        const uint32_t bits = uok ? (hmask & __ldg(ep.fold_on_bits + user)) : 0u;
        const float thr = uok ? __ldg(ep.fold_thr + user) : INFINITY;
        float best = -INFINITY;
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          const float x = __uint_as_float(rr[(u * HP + h) >> 5][(u * HP + h) & 31]);
          best = fmaxf(best, ((bits >> h) & 1u) ? x : -INFINITY);
        }
        pass |= (unsigned long long)((best >= thr && best > -INFINITY) ? 1u : 0u) << u;
      }
      if (pass != 0ull) {
#pragma unroll
        for (int u = 0; u < UPB; ++u) {
          if (!((pass >> u) & 1ull)) continue;
          const int user = user0 + u;
          const uint32_t bits = hmask & __ldg(ep.fold_on_bits + user);
          float best = -INFINITY;
#pragma unroll
          for (int h = 0; h < HP; ++h) {
            const float x = __uint_as_float(rr[(u * HP + h) >> 5][(u * HP + h) & 31]);
            best = fmaxf(best, ((bits >> h) & 1u) ? x : -INFINITY);
          }
          int bh = 0;
#pragma unroll
          for (int h = HP - 1; h >= 0; --h) {
            const float x = __uint_as_float(rr[(u * HP + h) >> 5][(u * HP + h) & 31]);
            if (((bits >> h) & 1u) && x == best) bh = h;   // lowest head wins ties
          }
          const uint32_t slot = atomicAdd(ep.fold_cnt + user, 1u);
          if (slot < (uint32_t)ep.fold_cap) {
            uint32_t k = __float_as_uint(best);
            k = (k & 0x80000000u) ? ~k : (k | 0x80000000u);                      // ascending-order-preserving key
            ep.fold_keys[(int64_t)user * ep.fold_cap + slot] =
                ((unsigned long long)(~k) << 32) | (unsigned long long)(((uint32_t)m << 5) | (uint32_t)bh);
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < UPB; ++u) {
      const int user = user0 + u;
      if (user >= n_users) break;                          // warp-uniform
      const uint32_t bits = hmask & __ldg(ep.fold_on_bits + user);
      float best = -INFINITY;
#pragma unroll
      for (int h = 0; h < HP; ++h) {
        const float x = __uint_as_float(rr[(u * HP + h) >> 5][(u * HP + h) & 31]);
        best = fmaxf(best, ((bits >> h) & 1u) ? x : -INFINITY);
      }
      if (row_ok) {
        int bh = 0;
#pragma unroll
        for (int h = HP - 1; h >= 0; --h) {
          const float x = __uint_as_float(rr[(u * HP + h) >> 5][(u * HP + h) & 31]);
          if (((bits >> h) & 1u) && x == best) bh = h;
        }
        ((float*)ep.C)[(int64_t)user * ep.ldc + m] = best;                       // lanes = consecutive items: coalesced
        ((uint8_t*)ep.C2)[(int64_t)user * ep.ldc2 + m] = (uint8_t)bh;
      }
    }
  }
}

// CTAS = 1: one CTA per 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN
// tile: each CTA stages its own 128 A rows and HALF of the B tile, so per-CTA shared-memory and L2 operand
// traffic per flop drop by a third and the epilogue's transpose traffic fits beside the mainloop.
template <int MODE, int CTAS>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap* maps_a, const CUtensorMap* maps_b, const TcParams& p,
                                             const EpiParams& ep_in, const TcGroupArgs* ga) {
  EpiParams ep = ep_in;
  if (ep.alpha_dev && (ep.mode == B200REC_EPI_STORE || ep.mode == B200REC_EPI_ACCUM)) ep.alpha *= *ep.alpha_dev;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int bn_cta = p.BN / CTAS;                       // B rows staged by this CTA
  const uint32_t a_bytes = TC_BM * 128u;
  const uint32_t b_bytes = (uint32_t)bn_cta * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t epi_stage_base = bar_base + 256u;  // TC_EPI_WARPS x 4 KB transpose buffers

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(maps_a);
    tma_prefetch_desc(maps_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS * CTAS);   // pair: the peer's epilogue warps arrive remotely
    }
    mbar_fence_init();
  }
  if (CTAS == 2) cluster_sync_all();                   // peer barriers exist before anyone signals them
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_2sm(tmem_slot, TC_TMEM_COLS); else tmem_alloc(tmem_slot, TC_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                   // both halves of the pair own their TMEM columns
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: everything above (barriers, tensor-map prefetch, TMEM allocation, cluster syncs) ran under the previous
  // kernel's tail; nothing below may touch global memory before that kernel has completed
  pdl_trigger();
  pdl_wait();

  const int tiles_pg = p.num_m * p.num_n;              // p.num_m counts (128*CTAS)-row tiles
  const int num_tiles = tiles_pg * p.groups;
  const int num_k = (p.K + TC_BK - 1) / TC_BK;
  const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
  const int num_items = p.tail_split > 0 ? p.tail_first + (num_tiles - p.tail_first) * p.tail_split
                                          : num_tiles * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = tile0; item < num_items; item += tile_step) {
        const TcWork wk = tc_decode(p, item, num_tiles, num_k);
        const int gtile = wk.gtile;
        const int grp = gtile / tiles_pg, tile = gtile - grp * tiles_pg;
        const CUtensorMap* map_a = maps_a + grp;
        const CUtensorMap* map_b = maps_b + grp;
        const int tm = p.n_fastest ? tile / p.num_n : tile % p.num_m;
        const int tn = p.n_fastest ? tile % p.num_n : tile / p.num_m;
        const int m0 = tm * (TC_BM * CTAS) + (int)rank * TC_BM;
        const int n0 = tn * p.BN + (int)rank * bn_cta;
        const int kb_beg = wk.kb_beg, kb_end = wk.kb_end;
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + a_bytes;
          // pair: both CTAs credit the LEADER's full barrier (it gates the leader's MMA issue)
          const uint32_t fb = CTAS == 2 ? mapa_shared(full_bar(s), 0) : full_bar(s);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), stage_bytes * CTAS);
          const int k0 = kb * TC_BK;
          auto load = [&](uint32_t dst, const CUtensorMap* mp, int c0, int c1) {
            if (CTAS == 2) tma_load_2d_2sm(dst, mp, fb, c0, c1); else tma_load_2d(dst, mp, fb, c0, c1);
          };
          if (!p.a_mn) {
            load(sa, map_a, k0, m0);                       // box {64 k, 128 m}
          } else {
            for (int j = 0; j < TC_BM / 64; ++j)          // boxes {64 m, 64 k}
              load(sa + j * 8192u, map_a, m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            load(sb, map_b, k0, n0);                       // box {64 k, BN/CTAS n}
          } else {
            for (int j = 0; j < bn_cta / 64; ++j)
              load(sb + j * 8192u, map_b, n0 + 64 * j, k0);
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM * CTAS, p.BN, p.a_mn, p.b_mn);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int item = tile0; item < num_items; item += tile_step) {
        const TcWork wk = tc_decode(p, item, num_tiles, num_k);
        const int kb_beg = wk.kb_beg, kb_end = wk.kb_end;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * (uint32_t)p.BN;
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // K-major: 16 k-elements = 32 bytes inside the 128B swizzle span; 8-row groups 1024B apart.
            // MN-major: 16 k-rows = 2 x 1024B atoms; next 64-wide mn block 8192B (64 k-rows x 128B) away.
            const uint64_t da = p.a_mn ? umma_smem_desc(sa + k * 2048u, 8192u, 1024u)
                                       : umma_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = p.b_mn ? umma_smem_desc(sb + k * 2048u, 8192u, 1024u)
                                       : umma_smem_desc(sb + k * 32u, 16u, 1024u);
            const uint32_t accum = (kb > kb_beg || k > 0) ? 1u : 0u;
            if (CTAS == 2) umma_bf16_2sm(d_tmem, da, db, idesc, accum);
            else umma_bf16(d_tmem, da, db, idesc, accum);
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (CTAS == 2) umma_commit_2sm(empty_bar(s)); else umma_commit(empty_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CTAS == 2) umma_commit_2sm(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;        // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    const int e = warp - 2;              // 0..7
    const int half = e >> 2;             // the two warps of a quarter take alternate 32-column chunks
    const uint32_t stg = epi_stage_base + (uint32_t)e * 4096u;
    const int cg = lane & 7, sub = lane >> 3;
    int acc = 0;
    uint32_t acc_ph = 0;
    const EpiParams ep_base = ep;
    for (int item = tile0; item < num_items; item += tile_step) {
      const TcWork wk = tc_decode(p, item, num_tiles, num_k);
      const int gtile = wk.gtile, split = wk.split;
      const int grp = gtile / tiles_pg, tile = gtile - grp * tiles_pg;
      const int tm = p.n_fastest ? tile / p.num_n : tile % p.num_m;
      const int tn = p.n_fastest ? tile % p.num_n : tile / p.num_m;
      const int m0 = tm * (TC_BM * CTAS) + (int)rank * TC_BM;
      const int n0 = tn * p.BN;
      // tail item: raw fp32 partial accumulators of this CTA's 128 rows -> workspace slot (row pitch BN)
      float* const tail_dst = wk.tail_slot < 0 ? nullptr
          : p.tail_ws + ((int64_t)wk.tail_slot * (TC_BM * CTAS) + (int64_t)rank * TC_BM + quarter * 32) * p.BN;
      if (ga) {
        ep.C = ga->C[grp];
        ep.row_scale = ga->row_scale[grp];
        ep.nce_mref = ga->nce_mref[grp]; ep.nce_thr = ga->nce_thr[grp]; ep.nce_stats = ga->nce_stats[grp];
      }
      // NCE_EXP: this lane's row constants and the partial sums of the columns this warp converts in this tile
      float nce_mref = 0.f, nce_thr = INFINITY, nce_tau2 = 0.f, nce_s = 0.f, nce_w = 0.f, nce_gt = 0.f;
      if (MODE == B200REC_EPI_NCE_EXP) {
        const int m = m0 + quarter * 32 + lane;
        if (m < ep.M) { nce_mref = __ldg(ep.nce_mref + m); nce_thr = __ldg(ep.nce_thr + m); }
        nce_tau2 = __expf(fminf(fmaxf(__ldg(ep.nce_logit_scale), 0.f), 4.605170185988092f)) * 1.4426950408889634f;
      }
      if (p.split_k > 1) ep.C = (float*)ep_base.C + (int64_t)split * p.split_stride;   // partial tile of this k-range
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * (uint32_t)p.BN;
      // GT_BITS: ONE warp per lane quarter converts all chunks of its rows, so that a lane owns BN / 32 consecutive
      // words (32 bytes at BN = 256 = one full sector); alternating chunks between the two warps of a quarter made every
      // 4-byte word a partial-sector write 1 KB away from its neighbours and the K = 64 filter GEMM store-bound
      if (MODE == B200REC_EPI_GT_BITS) {
        // each warp converts BN / 64 CONSECUTIVE 32-column chunks of its rows: all TMEM loads are issued before one wait
        // (the K = 64 filter GEMM is epilogue-latency bound), and a lane ends up with 2-4 consecutive words = one 8 / 16
        // byte store instead of scattered 4-byte words
        const int nch = p.BN / 64;                                   // 2 (BN = 128) or 4 (BN = 256)
        const int m = m0 + quarter * 32 + lane;
        uint32_t rr[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nch) tmem_ld_32x32_nowait(t_row + (uint32_t)(half * nch + q) * 32u, rr[q]);
        tmem_ld_wait();
        const float ra = (ep.gt_row != nullptr && m < ep.M) ? __ldg(ep.gt_row + m) : 0.f;
        uint32_t words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q >= nch) continue;
          const int nn = n0 + (half * nch + q) * 32;
          uint32_t wbits = 0;
          if (nn < ep.N) {
            if (ep.gt_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float cb = nn + i < ep.N ? __ldg(ep.gt_col + nn + i) : 0.f;
                wbits |= (fmaf(ra, cb, __uint_as_float(rr[q][i])) > ep.alpha) ? (1u << i) : 0u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) wbits |= (__uint_as_float(rr[q][i]) > ep.alpha) ? (1u << i) : 0u;
            }
            if (nn + 32 > ep.N) wbits &= (1u << (ep.N - nn)) - 1u;
          }
          words[q] = wbits;
        }
        if (m < ep.M) {
          const int w0 = (n0 >> 5) + half * nch;
          const int n_words_row = (ep.N + 31) >> 5;
          uint32_t* dst = (uint32_t*)ep.C + (int64_t)m * ep.ldc + w0;
          if (nch == 4 && w0 + 4 <= n_words_row && (ep.ldc & 3) == 0 && (((uintptr_t)ep.C) & 15) == 0) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nch && w0 + q < n_words_row) dst[q] = words[q];
          }
          if ((words[0] | words[1] | words[2] | words[3]) != 0u && ep.C2) ((uint8_t*)ep.C2)[m] = 1;
        }
      }
      if (MODE == B200REC_EPI_FOLD_ITEMS) {
        // rows = items (lane = item), columns = (user, head): the fold over a user's heads is register-local
        const int m = m0 + quarter * 32 + lane;
        const uint32_t t_half = t_row + (uint32_t)(half * (p.BN / 2));
        const int nc0 = n0 + half * (p.BN / 2);
        switch (ep.fold_hp) {
          case 1: fold_items_warp<1, 2>(ep, p.BN / 2, t_half, m, nc0); break;
          case 2: fold_items_warp<2, 2>(ep, p.BN / 2, t_half, m, nc0); break;
          case 4: fold_items_warp<4, 2>(ep, p.BN / 2, t_half, m, nc0); break;
          case 8: fold_items_warp<8, 2>(ep, p.BN / 2, t_half, m, nc0); break;
          case 12: fold_items_warp<12, 3>(ep, p.BN / 2, t_half, m, nc0); break;
          default: fold_items_warp<16, 2>(ep, p.BN / 2, t_half, m, nc0); break;
        }
      }
      for (int c = half; c < p.BN / 32 && MODE != B200REC_EPI_GT_BITS && MODE != B200REC_EPI_FOLD_ITEMS; c += 2) {
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)c * 32u, v);
        if (MODE == B200REC_EPI_NCE_EXP) {
          // softmax numerators relative to the row's reference: e = 2^(tau2 * cos - mref), rounded to bf16 BEFORE it is
          // summed (the stored tile and the partial sums stay consistent); columns beyond N contribute nothing
          const int nn = n0 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float cosv = v[i];
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(cosv, nce_tau2, -nce_mref)));
            e = (nn + i < ep.N) ? __bfloat162float(__float2bfloat16_rn(e)) : 0.f;
            nce_s += e;
            nce_w = fmaf(e, cosv, nce_w);
            nce_gt += (nn + i < ep.N && cosv > nce_thr) ? 1.f : 0.f;
            v[i] = e;
          }
        }
        // phase 1: lane owns row `lane`; 16-byte chunk q goes to physical chunk q ^ (lane & 7)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                       "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                       : "memory");
        }
        __syncwarp();
        // phase 2: 8 lanes cover one 128-byte row segment, 4 rows per pass
        const int ncol = n0 + c * 32 + cg * 4;
        if (tail_dst != nullptr) {
#pragma unroll 2
          for (int j = 0; j < 8; ++j) {
            const int row = 4 * j + sub;
            const uint32_t addr = stg + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
            float x[4];
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                         : "r"(addr)
                         : "memory");
            store4<float>(tail_dst + (int64_t)row * p.BN + c * 32 + cg * 4, x);
          }
          __syncwarp();
          continue;
        }
        if (MODE == B200REC_EPI_FOLD_HEADS) {
          // rows of this warp = 32/hp users x hp heads.  Lane (cg, sub) folds rows [8*sub, 8*sub+8) of its
          // 4 columns; partial results of one user are combined across sub-lanes with shuffles.
          const int hp = ep.fold_hp;
          const int mrow0 = m0 + quarter * 32;
          uint32_t tg[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
          bool colok[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = ncol + i;
            colok[i] = n < ep.N && ((int64_t)n * ep.fold_id_stride + ep.fold_id_offset != 0);
            if (ep.fold_item_tags && n < ep.N) tg[i] = __ldg(ep.fold_item_tags + n);
          }
          const int grp = hp >= 8 ? 8 : hp;            // rows folded sequentially by one lane
#pragma unroll 1
          for (int g = 0; g < 8 / grp; ++g) {
            float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int bh[4] = {0, 0, 0, 0};
            for (int r = 0; r < grp; ++r) {
              const int row = sub * 8 + g * grp + r;
              const int m = mrow0 + row;
              const int h = row & (hp - 1);
              const uint32_t addr = stg + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
              float x[4];
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                           : "r"(addr)
                           : "memory");
              const bool on = m < ep.M && __ldg(ep.fold_head_on + m) != 0;
              // streamed variant with G groups of hp heads per user: row m = (user, group, h), head = group * hp + h
              const int hfull = ep.fold_groups > 1 ? ((m / hp) % ep.fold_groups) * hp + h : h;
              const int cat = ep.fold_head_cat ? __ldg(ep.fold_head_cat + hfull) : -1;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const bool keep = on && colok[i] && (cat < 0 || ((tg[i] >> cat) & 1u));
                const float v = keep ? x[i] : -INFINITY;
                if (v > best[i]) { best[i] = v; bh[i] = h; }   // strict >: lowest head wins ties
              }
            }
            // hp > 8: the user's heads are spread over hp/8 sub-lanes (lane stride 8)
            for (int o = 8; o < hp; o <<= 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float ov = __shfl_xor_sync(0xffffffffu, best[i], o);
                const int oh = __shfl_xor_sync(0xffffffffu, bh[i], o);
                if (ov > best[i] || (ov == best[i] && oh < bh[i])) { best[i] = ov; bh[i] = oh; }
              }
            }
            const int subs_per_user = hp >= 8 ? hp / 8 : 1;
            if ((sub % subs_per_user) == 0) {
              const int row_first = sub * 8 + g * grp;
              const int64_t user = (int64_t)(mrow0 + row_first) / hp;
              if (ep.fold_thr != nullptr) {
                // streamed eval: only scores that can still enter the user's top-K leave the SM
                if (mrow0 + row_first < ep.M && ncol < ep.N) {
                  const int grp = ep.fold_groups > 1 ? (int)(user % ep.fold_groups) : 0;
                  const int64_t vuser = user;
                  const int64_t user = ep.fold_groups > 1 ? vuser / ep.fold_groups : vuser;
                  const float thr = __ldg(ep.fold_thr + user);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    if (ncol + i < ep.N && best[i] >= thr && best[i] > -INFINITY) {
                      const uint32_t slot = atomicAdd(ep.fold_cnt + user, 1u);
                      if (slot < (uint32_t)ep.fold_cap) {
                        uint32_t u = __float_as_uint(best[i]);
                        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // ascending-order-preserving key
                        ep.fold_keys[user * ep.fold_cap + slot] =
                            ((unsigned long long)(~u) << 32) |
                            (unsigned long long)(((uint32_t)(ncol + i) << 5) | (uint32_t)(grp * hp + bh[i]));  // item < 2^27, head < 32
                      }
                    }
                  }
                }
              } else if (mrow0 + row_first < ep.M && ncol < ep.N) {
                float* fv = (float*)ep.C + user * ep.ldc + ncol;
                uint8_t* fh = (uint8_t*)ep.C2 + user * ep.ldc2 + ncol;
                if (ncol + 4 <= ep.N) {
                  store4<float>(fv, best);
                  *reinterpret_cast<uint32_t*>(fh) = (uint32_t)bh[0] | ((uint32_t)bh[1] << 8) | ((uint32_t)bh[2] << 16) |
                                                     ((uint32_t)bh[3] << 24);
                } else {
                  for (int i = 0; i < 4 && ncol + i < ep.N; ++i) { fv[i] = best[i]; fh[i] = (uint8_t)bh[i]; }
                }
              }
            }
          }
          __syncwarp();
          continue;
        }
        if (epi_has_fast_chunk<MODE>() && ep.vec_ok && n0 + c * 32 + 32 <= ep.N) {
          epi_chunk_fast<MODE>(ep, stg, m0 + quarter * 32, ncol, sub, cg);   // warp-uniform condition
          __syncwarp();
          continue;
        }
        if (epi_needs_prefetch<MODE>()) {
          // issue every row-dependent global load of the chunk before consuming any (latency overlap)
          float pre[8][4];
          bool okv[8];
          float bias4[4] = {0.f, 0.f, 0.f, 0.f};
          if (MODE != B200REC_EPI_ACCUM && ep.bias && ep.vec_ok && ncol + 4 <= ep.N) load4<float>(ep.bias + ncol, bias4);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            okv[j] = epi_prefetch_vec4<MODE>(ep, m0 + quarter * 32 + 4 * j + sub, ncol, pre[j]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 4 * j + sub;
            const uint32_t addr = stg + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
            float x[4];
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                         : "r"(addr)
                         : "memory");
            const int m = m0 + quarter * 32 + row;
            if (okv[j]) epi_finish_vec4<MODE>(ep, m, ncol, x, pre[j], bias4);
            else if (m < ep.M && ncol < ep.N) epi_apply_tail4(ep, m, ncol, x[0], x[1], x[2], x[3]);
          }
        } else {
#pragma unroll 2
          for (int j = 0; j < 8; ++j) {
            const int row = 4 * j + sub;
            const uint32_t addr = stg + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
            float x[4];
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                         : "r"(addr)
                         : "memory");
            epi_apply_vec4<MODE>(ep, m0 + quarter * 32 + row, ncol, x);
          }
        }
        __syncwarp();
      }
      if (MODE == B200REC_EPI_NCE_EXP) {
        const int m = m0 + quarter * 32 + lane;
        if (m < ep.M) {
          const int part = tn * 2 + half;
          float st4[4] = {nce_s, nce_w, nce_gt, 0.f};
          store4<float>(ep.nce_stats + ((int64_t)m * ep.nce_parts + part) * 4, st4);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();                   // the peer may still be reading operands / TMEM
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2sm(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

template <int MODE, int CTAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const TcParams p, const EpiParams ep_in) {
  gemm_tc_body<MODE, CTAS>(&map_a, &map_b, p, ep_in, nullptr);
}

template <int MODE, int CTAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_grouped_kernel(const __grid_constant__ TcGroupArgs ga, const TcParams p, const EpiParams ep_in) {
  gemm_tc_body<MODE, CTAS>(ga.a, ga.b, p, ep_in, &ga);
}

// sums the split-K partial outputs in ascending split order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int split_k, int64_t stride,
                                                            int64_t n4, float* __restrict__ out, float alpha,
                                                            const float* __restrict__ alpha_dev) {
  if (alpha_dev) alpha *= *alpha_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < split_k; ++s) {
      float v[4];
      load4<float>(ws + (int64_t)s * stride + i * 4, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] *= alpha;
    store4<float>(out + i * 4, acc);
  }
}

// Tail split, second half: sums the k-range partials of every tail tile in ascending split order (deterministic) and
// applies the real epilogue.  One thread per 4 consecutive columns of a row; grid = (tile chunks, tail tiles).
template <int MODE>
__global__ void __launch_bounds__(256) tail_fixup_kernel(const TcParams p, const EpiParams ep_in, int rows_per_tile) {
  pdl_trigger();
  EpiParams ep = ep_in;
  if (ep.alpha_dev && (ep.mode == B200REC_EPI_STORE || ep.mode == B200REC_EPI_ACCUM)) ep.alpha *= *ep.alpha_dev;
  const int tt = blockIdx.y;
  const int tile = p.tail_first + tt;
  const int tm = p.n_fastest ? tile / p.num_n : tile % p.num_m;
  const int tn = p.n_fastest ? tile % p.num_n : tile / p.num_m;
  const int bn4 = p.BN / 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = idx / bn4, c4 = idx - row * bn4;
  if (row >= rows_per_tile) return;
  const int m = tm * rows_per_tile + row, n = tn * p.BN + c4 * 4;
  if (m >= ep.M || n >= ep.N) return;
  const int64_t tile_elems = (int64_t)rows_per_tile * p.BN;
  const float* src = p.tail_ws + (int64_t)tt * p.tail_split * tile_elems + (int64_t)row * p.BN + c4 * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int sidx = 0; sidx < p.tail_split; ++sidx) {
    float v[4];
    load4<float>(src + (int64_t)sidx * tile_elems, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += v[k];
  }
  epi_apply_vec4<MODE>(ep, m, n, acc);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

// 2D bf16 tensor map: inner (contiguous) extent d0, outer extent d1, row pitch ld elements;
// swizzle_bytes in {32, 64, 128} must equal the inner box width in bytes.
int b200_make_map_bf16(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box0,
                       uint32_t box1, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  B200_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): ptr=%p d0=%llu d1=%llu ld=%llu box=%u,%u",
                 (int)r, ptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, box0, box1);
  return 0;
}

static int make_map(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box0,
                    uint32_t box1) {
  return b200_make_map_bf16(m, ptr, d0, d1, ld, box0, box1, 128);
}

static int g_num_sms = 0;
static int g_force_bn = 0;    // test hook
static int g_force_ctas = 0;  // test hook: 1 = never use CTA pairs
static int g_tail_split = 1;  // tail split of the last partial wave (B200REC_TAIL_SPLIT=0 / b200rec_gemm_use_tail_split(0))
static int g_use_pdl = 1;     // programmatic dependent launch of the GEMM kernels (B200REC_PDL=0 / b200rec_gemm_use_pdl(0))

extern "C" void b200rec_gemm_use_pdl(int on) { g_use_pdl = on ? 1 : 0; }
int b200rec_pdl_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B200REC_PDL");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return env && g_use_pdl;
}

extern "C" void b200rec_gemm_use_tail_split(int on) { g_tail_split = on ? 1 : 0; }
extern "C" void b200rec_gemm_force_bn(int bn) { g_force_bn = bn; }
// partial-sum slots per row of the NCE_EXP epilogue: two epilogue warps per BN-wide column tile
extern "C" int b200rec_gemm_nce_parts(int N) { return 2 * ceil_div_i(N, N > 128 ? 256 : 128); }
extern "C" void b200rec_gemm_force_ctas(int ctas) { g_force_ctas = ctas; }

int gemm_tc_launch(const b200rec_gemm_args* a, const EpiParams& ep_in, cudaStream_t st) {
  B200_CHECK_ARG(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->B & 15) == 0, "gemm: A/B must be 16-byte aligned");
  B200_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 (16-byte pitch)");
  if (g_num_sms == 0) {
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    B200_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    const char* e = getenv("B200REC_PDL");
    if (e != nullptr && e[0] == '0') g_use_pdl = 0;
    e = getenv("B200REC_TAIL_SPLIT");
    if (e != nullptr && e[0] == '0') g_tail_split = 0;
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
  }
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major; p.b_mn = a->b_major;
  // CTA pairs whenever there are at least two 128-row tiles (and the test hook does not forbid it)
  const int ctas = (a->M > TC_BM && g_force_ctas != 1) ? 2 : 1;
  const int units = g_num_sms / ctas;                   // concurrently resident tiles
  {
    // measured on B200 (scripts/gemm_probe.py, SWEEP=1): the 256-wide tile wins on every shape of the path
    // (the 128-wide one is shared-memory-port bound) unless it leaves more than half of the SMs without a tile
    const int64_t tiles256 = (int64_t)ceil_div_i(a->M, TC_BM * ctas) * ceil_div_i(a->N, 256);
    p.BN = (a->N > 128 && tiles256 * 2 >= units) ? 256 : 128;
    if (ep_in.mode == B200REC_EPI_NCE_EXP) p.BN = a->N > 128 ? 256 : 128;   // nce_parts is a function of N alone
  }
  if ((g_force_bn == 128 || g_force_bn == 256) && ep_in.mode != B200REC_EPI_NCE_EXP) p.BN = g_force_bn;
  p.n_fastest = 0;
  if (ep_in.mode == B200REC_EPI_FOLD_ITEMS) {
    // each epilogue warp owns BN / 2 columns = whole users of fold_hp heads: 12 heads -> 192-wide tiles (8 users per
    // warp), powers of two -> 256 (or 128 when there are few users).  All user tiles of one item tile run back to back,
    // so the item table streams from HBM once and the (small) user-head operand stays in L2.
    p.BN = ep_in.fold_hp == 12 ? 192 : (a->N > 128 ? 256 : 128);
    p.n_fastest = 1;
  }
  p.split_k = 1;
  p.groups = 1;
  const int num_k_host = ceil_div_i(a->K, TC_BK);
  if (a->splitk_ws != nullptr && ep_in.mode == B200REC_EPI_STORE && a->c_dtype == B200REC_F32 && a->n_split == 0 &&
      g_force_bn == 0 && a->N > 128) {
    // few output tiles but a long K (weight gradients): spread the k-range of each tile over idle SMs
    const int64_t tiles256 = (int64_t)ceil_div_i(a->M, TC_BM * ctas) * ceil_div_i(a->N, 256);
    int sk = (int)std::min<int64_t>(std::min<int64_t>(units / std::max<int64_t>(tiles256, 1), num_k_host / 8), 8);
    if (sk >= 2 && (size_t)sk * a->M * a->ldc * sizeof(float) <= a->splitk_ws_bytes) {
      p.BN = 256;
      p.kb_per_split = ceil_div_i(num_k_host, sk);
      p.split_k = ceil_div_i(num_k_host, p.kb_per_split);
      p.split_stride = (int64_t)a->M * a->ldc;
    }
  }
  if (p.split_k == 1) p.kb_per_split = num_k_host;
  if (a->n_split > 0)
    B200_CHECK_ARG(a->n_split % 32 == 0, "gemm: n_split must be a multiple of 32");
  const uint32_t stage_bytes = TC_BM * 128u + (uint32_t)(p.BN / ctas) * 128u;
  p.stages = (TC_SMEM_LIMIT - 1024 - 256 - TC_STAGE_BYTES) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.num_m = ceil_div_i(a->M, TC_BM * ctas);
  p.num_n = ceil_div_i(a->N, p.BN);
  CUtensorMap ma, mb;
  if (!p.a_mn) {
    if (make_map(&ma, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, 64, TC_BM)) return 1;
  } else {
    if (make_map(&ma, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, 64)) return 1;
  }
  if (!p.b_mn) {
    if (make_map(&mb, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, 64, (uint32_t)(p.BN / ctas))) return 1;
  } else {
    if (make_map(&mb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, 64)) return 1;
  }
  // Tail split: with a few tiles more than a whole number of waves (O-proj, d_oin, dn of config B: 84 tiles on 74 CTA
  // pairs) the last wave keeps most SMs idle for a full tile time.  Its r tiles are cut into s = units / r k-ranges so
  // that the wave lasts 1/s of a tile; partials meet in the workspace and tail_fixup_kernel applies the epilogue.
  p.tail_first = 0; p.tail_split = 0; p.tail_kb = 0; p.tail_ws = nullptr;
  {
    const int tiles_total = p.num_m * p.num_n;
    const int r = tiles_total % units;
    if (g_tail_split && a->splitk_ws != nullptr && p.split_k == 1 && ep_in.mode <= B200REC_EPI_RESBLOCK &&
        tiles_total > units && r > 0 && 2 * r <= units && num_k_host >= 32) {
      // (measured: at K = 1024 the fix-up launch costs more than the saved 13 k-blocks; K = 4096: -9 us, K = 12288: -48 us)
      int sp = std::min(units / r, num_k_host / 2);
      if (sp >= 2) {
        const int kb = ceil_div_i(num_k_host, sp);
        sp = ceil_div_i(num_k_host, kb);
        const size_t need = (size_t)r * sp * (TC_BM * ctas) * p.BN * sizeof(float);
        if (sp >= 2 && need <= a->splitk_ws_bytes && ((uintptr_t)a->splitk_ws & 15) == 0) {
          p.tail_first = tiles_total - r; p.tail_split = sp; p.tail_kb = kb; p.tail_ws = (float*)a->splitk_ws;
        }
      }
    }
  }
  int tiles = p.tail_split > 0 ? p.tail_first + (p.num_m * p.num_n - p.tail_first) * p.tail_split
                               : p.num_m * p.num_n * p.split_k;
  int grid = (tiles < units ? tiles : units) * ctas;
  EpiParams ep = ep_in;
  if (p.split_k > 1) {        // partial sums go to the workspace (alpha applied by the final reduction)
    ep.C = a->splitk_ws;
    ep.alpha = 1.f;
    ep.alpha_dev = nullptr;
    ep.vec_ok = 1;
  }
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + TC_STAGE_BYTES;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
#define TC_LAUNCH(MODE_)                                                                                   \
  do {                                                                                                     \
    if (ctas == 2) B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<MODE_, 2>, ma, mb, p, ep));        \
    else B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<MODE_, 1>, ma, mb, p, ep));                  \
  } while (0)
  switch (ep.mode) {
    case B200REC_EPI_STORE: TC_LAUNCH(0); break;
    case B200REC_EPI_ACCUM: TC_LAUNCH(1); break;
    case B200REC_EPI_SILU_DUAL: TC_LAUNCH(2); break;
    case B200REC_EPI_BIAS_RESID: TC_LAUNCH(3); break;
    case B200REC_EPI_RESBLOCK: TC_LAUNCH(4); break;
    case B200REC_EPI_GT_BITS: TC_LAUNCH(5); break;
    case B200REC_EPI_FOLD_HEADS: TC_LAUNCH(6); break;
    case B200REC_EPI_NCE_EXP: TC_LAUNCH(7); break;
    case B200REC_EPI_FOLD_ITEMS: TC_LAUNCH(8); break;
    default: b200rec_set_error("gemm: bad epilogue %d", ep.mode); return 1;
  }
#undef TC_LAUNCH
  if (p.tail_split > 0) {
    const int rows = TC_BM * ctas;
    dim3 fgrid(ceil_div_i(rows * (p.BN / 4), 256), p.num_m * p.num_n - p.tail_first);
    switch (ep.mode) {
      case B200REC_EPI_STORE: tail_fixup_kernel<0><<<fgrid, 256, 0, st>>>(p, ep, rows); break;
      case B200REC_EPI_ACCUM: tail_fixup_kernel<1><<<fgrid, 256, 0, st>>>(p, ep, rows); break;
      case B200REC_EPI_SILU_DUAL: tail_fixup_kernel<2><<<fgrid, 256, 0, st>>>(p, ep, rows); break;
      case B200REC_EPI_BIAS_RESID: tail_fixup_kernel<3><<<fgrid, 256, 0, st>>>(p, ep, rows); break;
      default: tail_fixup_kernel<4><<<fgrid, 256, 0, st>>>(p, ep, rows); break;
    }
  }
  if (p.split_k > 1) {
    const int64_t n4 = (int64_t)a->M * a->ldc / 4;
    splitk_reduce_kernel<<<(int)std::min<int64_t>((n4 + 255) / 256, 148 * 8), 256, 0, st>>>(
        (const float*)a->splitk_ws, p.split_k, p.split_stride, n4, (float*)a->C, ep_in.alpha, ep_in.alpha_dev);
  }
  B200_LAUNCH_OK();
  return 0;
}

// Grouped launch: n (<= TC_MAX_GROUPS) problems with identical shape / layout / epilogue (STORE or ACCUM, no bias /
// residual), different A, B, C.  The caller guarantees that the outputs do not alias.
int gemm_tc_launch_grouped(const b200rec_gemm_args* a, int n, const EpiParams& ep_in, cudaStream_t st) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaFuncSetAttribute(gemm_tc_grouped_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    B200_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major; p.b_mn = a->b_major;
  const int ctas = (a->M > TC_BM && g_force_ctas != 1) ? 2 : 1;
  const int units = num_sms / ctas;
  p.BN = a->N > 128 ? 256 : 128;
  if ((g_force_bn == 128 || g_force_bn == 256) && ep_in.mode != B200REC_EPI_NCE_EXP) p.BN = g_force_bn;
  p.split_k = 1;
  p.kb_per_split = ceil_div_i(a->K, TC_BK);
  p.split_stride = 0;
  p.groups = n;
  p.tail_first = 0; p.tail_split = 0; p.tail_kb = 0; p.tail_ws = nullptr;
  p.n_fastest = 0;
  const uint32_t stage_bytes = TC_BM * 128u + (uint32_t)(p.BN / ctas) * 128u;
  p.stages = (TC_SMEM_LIMIT - 1024 - 256 - TC_STAGE_BYTES) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  p.num_m = ceil_div_i(a->M, TC_BM * ctas);
  p.num_n = ceil_div_i(a->N, p.BN);
  TcGroupArgs ga;
  memset(&ga, 0, sizeof(ga));
  for (int g = 0; g < n; ++g) {
    const b200rec_gemm_args* x = a + g;
    B200_CHECK_ARG(((uintptr_t)x->A & 15) == 0 && ((uintptr_t)x->B & 15) == 0, "gemm_grouped: A/B must be 16-byte aligned");
    B200_CHECK_ARG(x->lda % 8 == 0 && x->ldb % 8 == 0, "gemm_grouped: lda/ldb must be multiples of 8");
    if (!p.a_mn) {
      if (make_map(&ga.a[g], x->A, (uint64_t)x->K, (uint64_t)x->M, (uint64_t)x->lda, 64, TC_BM)) return 1;
    } else {
      if (make_map(&ga.a[g], x->A, (uint64_t)x->M, (uint64_t)x->K, (uint64_t)x->lda, 64, 64)) return 1;
    }
    if (!p.b_mn) {
      if (make_map(&ga.b[g], x->B, (uint64_t)x->K, (uint64_t)x->N, (uint64_t)x->ldb, 64, (uint32_t)(p.BN / ctas))) return 1;
    } else {
      if (make_map(&ga.b[g], x->B, (uint64_t)x->N, (uint64_t)x->K, (uint64_t)x->ldb, 64, 64)) return 1;
    }
    ga.C[g] = x->C;
    ga.row_scale[g] = x->row_scale;
    ga.nce_mref[g] = x->nce_mref; ga.nce_thr[g] = x->nce_thr; ga.nce_stats[g] = x->nce_stats;
  }
  const int tiles = p.num_m * p.num_n * n;
  const int grid = (tiles < units ? tiles : units) * ctas;
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + TC_STAGE_BYTES;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  EpiParams ep = ep_in;
  if (ep.mode == B200REC_EPI_STORE) {
    if (ctas == 2) B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<0, 2>, ga, p, ep));
    else B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<0, 1>, ga, p, ep));
  } else if (ep.mode == B200REC_EPI_ACCUM) {
    if (ctas == 2) B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<1, 2>, ga, p, ep));
    else B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<1, 1>, ga, p, ep));
  } else if (ep.mode == B200REC_EPI_GT_BITS) {
    if (ctas == 2) B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<5, 2>, ga, p, ep));
    else B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<5, 1>, ga, p, ep));
  } else if (ep.mode == B200REC_EPI_NCE_EXP) {
    if (ctas == 2) B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<7, 2>, ga, p, ep));
    else B200_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_grouped_kernel<7, 1>, ga, p, ep));
  } else {
    b200rec_set_error("gemm_grouped: epilogue %d not supported (STORE / ACCUM / GT_BITS / NCE_EXP)", ep.mode);
    return 1;
  }
  B200_LAUNCH_OK();
  return 0;
}
