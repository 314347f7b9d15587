// HSTU attention on the 5th-gen tensor cores (sm_100a): bf16 operands fed by TMA, QK^T / AV and the
// backward contractions issued as tcgen05.mma with fp32 accumulators in TMEM; the pointwise SiLU, the
// 1/n_pad scale and the jagged causal mask run on the TMEM->register path and the probabilities go
// back to shared memory (128B-swizzled K-major) as the A operand of the second MMA.  Scores never
// touch HBM.
//
// Tiling is over the GLOBAL jagged token axis: a query tile is 128 consecutive
// tokens (it may span several short sequences), its key tiles run from the tile holding the first
// token's sequence start up to the diagonal tile, and the mask is
//   keep(i, j) = seq_start(i) <= j <= i  and  key_valid[j]
// so sequences of any length (L = 50 ... 400+) use the same kernel with no padding rows.
// Backward = two deterministic passes (dQ per query tile; dK/dV per key tile, computed in the
// transposed orientation so that keys sit on the TMEM lanes), each recomputing S (App. D.1).
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

int b200_make_map_bf16(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box0,
                       uint32_t box1, int swizzle_bytes);

// sequence index of token t: largest b with seq_off[b] <= t
__device__ __forceinline__ int find_seq(const int32_t* __restrict__ seq_off, int B, int t) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (seq_off[mid] <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// 32 consecutive columns [col0, col0+32) of row `row` -> bf16 into a K-major SWIZZLE_128B tile made of
// 64-column blocks of [R rows x 128 bytes] (the canonical UMMA A-operand layout TMA would produce).
__device__ __forceinline__ void put_row32_sw128(uint32_t tile, int R, int row, int col0, const float (&v)[32]) {
  const uint32_t base = tile + (uint32_t)(col0 >> 6) * (uint32_t)(R * 128) + (uint32_t)row * 128u;
  const int cin = (col0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
      w[e] = *reinterpret_cast<uint32_t*>(&h2);
    }
    const uint32_t addr = base + (uint32_t)(((cin + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}

template <int DH>
struct AtCfg {
  static constexpr int SWZ = DH * 2;                       // bytes per tile row = TMA swizzle span
  static constexpr uint32_t LAY = SWZ == 128 ? 2u : 4u;    // UMMA layout code
  static constexpr uint32_t ATOM = 8 * SWZ;                // 8-row swizzle atom
  // K-major operand (rows = M/N index, DH along K), k-step ks of 16 elements
  static __device__ __forceinline__ uint64_t desc_k(uint32_t tile, int ks) {
    return umma_smem_desc_l(tile + ks * 32, 16, ATOM, LAY);
  }
  // MN-major B operand ([K rows][N = DH] as stored), k-step ks of 16 rows
  static __device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int ks) {
    return umma_smem_desc_l(tile + ks * 16 * SWZ, ATOM, ATOM, LAY);
  }
};
// K-major A operand written by put_row32_sw128 (128 rows, K = keys/queries), k-step ks of 16 columns
__device__ __forceinline__ uint64_t desc_p(uint32_t tile, int ks) {
  return umma_smem_desc_l(tile + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024, 2u);
}

// --------------------------------------------------------------------------------------- forward
// zero a 32-column chunk of a row (fully masked chunk: no TMEM read, no SiLU)
__device__ __forceinline__ void put_zero32_sw128(uint32_t tile, int R, int row, int col0) {
  const uint32_t base = tile + (uint32_t)(col0 >> 6) * (uint32_t)(R * 128) + (uint32_t)row * 128u;
  const int cin = (col0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = base + (uint32_t)(((cin + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
  }
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 128 queries x 64-key tiles: 48 KB of shared memory and 128 TMEM columns per CTA -> 4 CTAs per SM hide the
// TMA -> MMA -> SiLU -> MMA latency chain of one another (the kernel is latency-, not throughput-bound).
template <int DH>
__global__ void __launch_bounds__(128)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                   const int32_t* __restrict__ seq_off, int B, const uint8_t* __restrict__ key_valid, int T, int D,
                   float inv_n, float* __restrict__ out) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + TILE, sV = sK + TILE / 2, sP = sV + TILE / 2;  // sP: 128 x 128 B
  const uint32_t bars = sP + 16384;
  const uint32_t bar_q = bars, bar_kv = bars + 8, bar_s = bars + 16, bar_o = bars + 24, tslot = bars + 32;
  __shared__ uint8_t s_kvalid[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  const int ti = q0 + tid;
  if (tid == 0) {
    tma_prefetch_desc(&map128);
    tma_prefetch_desc(&map64);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tS = 0, tO = 64;  // TMEM columns

  const int my_start = ti < T ? seq_off[find_seq(seq_off, B, ti)] : INT_MAX;
  const int w_min_start = warp_min_i(my_start);           // earliest key any row of this warp may see
  const int w_max_t = min(q0 + warp * 32 + 31, T - 1);    // latest key (causal) any row of this warp may see
  const int kt_first = seq_off[find_seq(seq_off, B, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, TILE);
    tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
  }
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
  uint32_t ph = 0;
  for (int kt = kt_first; kt <= kt_last; ++kt, ph ^= 1u) {
    const int k0 = kt * 64;
    if (tid == 0) {
      if (kt > kt_first) mbar_wait(bar_o, ph ^ 1u);  // previous P*V retired: K, V and P buffers are free
      mbar_arrive_expect_tx(bar_kv, TILE);
      tma_load_2d(sK, &map64, bar_kv, 3 * D + h * DH, k0);
      tma_load_2d(sV, &map64, bar_kv, 1 * D + h * DH, k0);
    }
    if (tid < 64) s_kvalid[tid] = (k0 + tid < T) ? key_valid[k0 + tid] : 0;
    if (tid == 0) {
      if (kt == kt_first) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + tS, C::desc_k(sQ, ks), C::desc_k(sK, ks), idesc_s, ks > 0);
      umma_commit(bar_s);
    }
    __syncthreads();
    mbar_wait(bar_s, ph);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int c0 = k0 + c * 32;
      if (c0 > w_max_t || c0 + 31 < w_min_start) {   // warp-uniform: whole chunk masked for these 32 rows
        put_zero32_sw128(sP, 128, tid, c * 32);
        continue;
      }
      float v[32];
      tmem_ld_32x32(t_lane + tS + c * 32, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int tj = c0 + e;
        const bool keep = (tj <= ti) && (tj >= my_start) && s_kvalid[c * 32 + e];
        v[e] = keep ? silu_f(v[e]) * inv_n : 0.f;
      }
      put_row32_sw128(sP, 128, tid, c * 32, v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tO, desc_p(sP, ks), C::desc_mn(sV, ks), idesc_o, (kt > kt_first || ks > 0));
      umma_commit(bar_o);
    }
  }
  mbar_wait(bar_o, ph ^ 1u);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < DH / 32; ++c) {
    float v[32];
    tmem_ld_32x32(t_lane + tO + c * 32, v);
    if (ti < T) {
      float* dst = out + (int64_t)ti * D + h * DH + c * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float o4[4] = {v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]};
        store4<float>(dst + 4 * e, o4);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}


// Sequence offsets staged in shared memory: the per-CTA prologue (and the dK/dV pass's per-tile metadata) does a binary
// search per thread; against global memory that is ~6 dependent loads of L2 latency, a third of a short CTA's lifetime.
#define AT_SEQ_SMEM 1025
struct SeqTab {
  const int32_t* p;     // shared-memory copy when it fits, else the global array
  int B;
};
__device__ __forceinline__ SeqTab seq_tab_load(int32_t* s_seq, const int32_t* __restrict__ seq_off, int B) {
  SeqTab t;
  t.B = B;
  if (B + 1 <= AT_SEQ_SMEM) {
    for (int i = threadIdx.x; i <= B; i += blockDim.x) s_seq[i] = seq_off[i];
    t.p = s_seq;
  } else {
    t.p = seq_off;
  }
  return t;     // visible after the caller's __syncthreads()
}
__device__ __forceinline__ int find_seq(const SeqTab& t, int tok) { return find_seq(t.p, t.B, tok); }

// 32 packed bf16 pairs of this thread's chunk -> K-major SWIZZLE_128B tile [128 rows x 64 cols]
__device__ __forceinline__ void put_pk32_sw128(uint32_t tile, int row, int colhalf, const uint32_t (&pk)[16]) {
  const uint32_t rbase = tile + (uint32_t)row * 128u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = rbase + (uint32_t)(((colhalf * 4 + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                 "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  tmem_ld_wait();
}


// --------------------------------------------------------------------------------------- forward, pipelined
// Same tiling (128 queries x 64-key tiles), but the per-key-tile chain  TMA -> S = Q K^T -> SiLU -> P -> O += P V  is
// software-pipelined inside the CTA and WARP-SPECIALISED (ncu of the first pipelined version, where thread 0 issued
// between block barriers: 55 % of the stall samples sat in __syncthreads waiting for that one thread's serial
// instruction stream):
//   * warps 0..7 = 256 pointwise threads, two per query row (32 of the 64 keys each); warp 8 = issuer: one elected
//     lane issues every TMA and every tcgen05.mma.  There is NO block barrier in the loop: the pointwise threads hand
//     P over through an mbarrier (256 arrivals), the issuer hands S / O back through tcgen05.commit;
//   * TWO S accumulators in TMEM: S(it+2) is issued right behind PV(it), so it retires one full iteration before the
//     pointwise threads need it; K in a 3-deep, V in a 2-deep shared-memory ring, loaded >= 1 iteration ahead;
//   * chunks with no masked element (interior of a sequence, keys all valid) skip every compare / select; key validity
//     travels as one ballot word per 32-key chunk;
//   * 1/n_pad is applied once to the output row instead of to every probability.
// The pointwise SiLU (one MUFU per score against 256 tensor flops at dh = 64) bounds the kernel at 1/2 of the tensor
// peak.  72 KB of shared memory and 256 TMEM columns (S0 | S1 | O) per CTA: two CTAs per SM.
#define AT_THREADS 288
__device__ __forceinline__ uint32_t kvalid_bits(const uint8_t* __restrict__ key_valid, int c0, int lane, int T) {
  const int kk = c0 + lane;
  return __ballot_sync(0xffffffffu, kk >= 0 && kk < T && key_valid[kk] != 0);
}

template <int DH>
__global__ void __launch_bounds__(AT_THREADS)
attn_tc_fwd_pipe_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                        const int32_t* __restrict__ seq_off, int B, const uint8_t* __restrict__ key_valid, int T, int D,
                        float inv_n, float* __restrict__ out) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ, HT = TILE / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int NS = 3;                       // K / V ring depth (NS = 4 measured no faster: TMA latency is hidden)
  const uint32_t sQ = base, sK0 = sQ + TILE, sV0 = sK0 + NS * HT, sP0 = sV0 + NS * HT;   // sP0: two 16 KB P tiles
  const uint32_t bars = sP0 + 2 * 16384;
  const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bar_k0 + 8 * NS, bar_s0 = bar_v0 + 8 * NS,
                 bar_o0 = bar_s0 + 16, bar_p0 = bar_o0 + 16, tslot = bar_p0 + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(bar_k0 + 8 * i, 1); mbar_init(bar_v0 + 8 * i, 1); }
    // two hand-over barriers, alternating by tile parity: with P double-buffered a fast warp may arrive for tile it+1
    // before a slow one has arrived for tile it, and one barrier must never see two arrivals of a thread in one phase
    for (int i = 0; i < 2; ++i) { mbar_init(bar_s0 + 8 * i, 1); mbar_init(bar_o0 + 8 * i, 1); mbar_init(bar_p0 + 8 * i, 256); }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tslot, 256);
  __shared__ int32_t s_seq[AT_SEQ_SMEM];
  const SeqTab seqs = seq_tab_load(s_seq, seq_off, B);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int kt_first = seqs.p[find_seq(seqs, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  const int n_it = kt_last - kt_first + 1;
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t tO = 128;  // TMEM columns: S0 = 0, S1 = 64, O = 128

  if (warp == 8) {
    // ===================== issuer =====================
    if (lane == 0) {
      tma_prefetch_desc(&map128);
      tma_prefetch_desc(&map64);
      const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
      auto load_k = [&](int it) {
        const int j = it % NS;
        mbar_arrive_expect_tx(bar_k0 + 8 * j, HT);
        tma_load_2d(sK0 + j * HT, &map64, bar_k0 + 8 * j, 3 * D + h * DH, (kt_first + it) * 64);
      };
      auto load_v = [&](int it) {
        const int j = it % NS;
        mbar_arrive_expect_tx(bar_v0 + 8 * j, HT);
        tma_load_2d(sV0 + j * HT, &map64, bar_v0 + 8 * j, 1 * D + h * DH, (kt_first + it) * 64);
      };
      auto issue_s = [&](int it) {
        const int j = it % NS;
        mbar_wait(bar_k0 + 8 * j, (uint32_t)(it / NS) & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16(tmem + 64 * (it & 1), C::desc_k(sQ, ks), C::desc_k(sK0 + j * HT, ks), idesc_s, ks > 0);
        umma_commit(bar_s0 + 8 * (it & 1));
      };
      mbar_arrive_expect_tx(bar_q, TILE);
      tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
      for (int i = 0; i < NS && i < n_it; ++i) { load_k(i); load_v(i); }
      mbar_wait(bar_q, 0);
      issue_s(0);
      if (n_it > 1) issue_s(1);
      for (int it = 0; it < n_it; ++it) {
        const int b = it & 1;
        mbar_wait(bar_p0 + 8 * b, (uint32_t)(it >> 1) & 1u);   // P(it) written; S(it) consumed (so S[b] and K(it)'s buffer are free)
        tc_fence_after();
        if (it + NS < n_it) load_k(it + NS);
        if (it >= 2 && it - 2 + NS < n_it) {
          // V(it-2)'s slot: the pointwise threads only arrive on bar_p(it) after seeing PV(it-2) retire
          load_v(it - 2 + NS);
        }
        mbar_wait(bar_v0 + 8 * (it % NS), (uint32_t)(it / NS) & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + tO, desc_p(sP0 + b * 16384, ks), C::desc_mn(sV0 + (it % NS) * HT, ks), idesc_o,
                    (it > 0 || ks > 0));
        umma_commit(bar_o0 + 8 * b);
        if (it + 2 < n_it) issue_s(it + 2);
      }
    }
  } else {
    // ===================== pointwise warps =====================
    const int quarter = warp & 3, colhalf = warp >> 2;
    const int row = quarter * 32 + lane;
    const int ti = q0 + row;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    const int my_start = ti < T ? seqs.p[find_seq(seqs, ti)] : INT_MAX;
    const int w_min_start = warp_min_i(my_start);                       // earliest key any row of this warp may see
    const int w_max_start = warp_max_i(ti < T ? my_start : INT_MIN);    // latest sequence start among the warp's rows
    const int w_min_t = q0 + quarter * 32;                              // first row of the warp
    const int w_max_t = min(q0 + quarter * 32 + 31, T - 1);             // latest key (causal) any row may see
    const bool w_all_rows = q0 + quarter * 32 + 31 < T;
    // key validity of this thread's key column, fetched one iteration ahead of its use
    auto kv_fetch = [&](int it) -> uint8_t {
      const int kk = (kt_first + it) * 64 + colhalf * 32 + lane;
      return (it < n_it && kk < T) ? key_valid[kk] : (uint8_t)0;
    };
    uint8_t kv_next = kv_fetch(0);
    for (int it = 0; it < n_it; ++it) {
      const int b = it & 1;
      const int c0 = (kt_first + it) * 64 + colhalf * 32;
      const bool skip = c0 > w_max_t || c0 + 31 < w_min_start;   // warp-uniform: whole chunk masked for these 32 rows
      const uint32_t kvb = __ballot_sync(0xffffffffu, kv_next != 0);
      kv_next = kv_fetch(it + 1);
      mbar_wait(bar_s0 + 8 * b, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      uint32_t pk[16];
      if (skip) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      } else {
        float v[32];
        tmem_ld_32x32(t_lane + 64 * b + colhalf * 32, v);
        if (w_all_rows && c0 + 31 <= w_min_t && c0 >= w_max_start && kvb == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(silu_fast_f(v[2 * e]), silu_fast_f(v[2 * e + 1]));
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float a[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int tj = c0 + 2 * e + u;
              const bool keep = (tj <= ti) && (tj >= my_start) && ((kvb >> (2 * e + u)) & 1u);
              a[u] = keep ? silu_fast_f(v[2 * e + u]) : 0.f;
            }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a[0], a[1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
      }
      if (it > 1) mbar_wait(bar_o0 + 8 * b, (uint32_t)((it >> 1) - 1) & 1u);   // PV(it-2) retired: this P tile is free
      put_pk32_sw128(sP0 + b * 16384, row, colhalf, pk);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p0 + 8 * b);
    }
    if (n_it > 1) mbar_wait(bar_o0 + 8 * ((n_it - 2) & 1), (uint32_t)((n_it - 2) >> 1) & 1u);
    mbar_wait(bar_o0 + 8 * ((n_it - 1) & 1), (uint32_t)((n_it - 1) >> 1) & 1u);
    tc_fence_after();
    constexpr int HALF = DH / 2;     // warp (quarter, colhalf) stores columns [colhalf * DH/2, +DH/2) of its 32 rows
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + tO + colhalf * HALF + c * 16, r);
      if (ti < T) {
        float* dst = out + (int64_t)ti * D + h * DH + colhalf * HALF + c * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float o4[4] = {__uint_as_float(r[4 * e]) * inv_n, __uint_as_float(r[4 * e + 1]) * inv_n,
                         __uint_as_float(r[4 * e + 2]) * inv_n, __uint_as_float(r[4 * e + 3]) * inv_n};
          store4<float>(dst + 4 * e, o4);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// --------------------------------------------------------------------------------------- backward: dQ
template <int DH>
__global__ void __launch_bounds__(128)
attn_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                      const __grid_constant__ CUtensorMap mapdo128, const int32_t* __restrict__ seq_off, int B,
                      const uint8_t* __restrict__ key_valid, int T, int D, float inv_n,
                      const bf16* __restrict__ pre_q, int64_t ld, bf16* __restrict__ d_pre_q) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + TILE, sK = sdO + TILE, sV = sK + TILE / 2, sdS = sV + TILE / 2;  // sdS 16 KB
  const uint32_t bars = sdS + 16384;
  const uint32_t bar_q = bars, bar_kv = bars + 8, bar_s = bars + 16, bar_o = bars + 24, tslot = bars + 32;
  __shared__ uint8_t s_kvalid[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  const int ti = q0 + tid;
  if (tid == 0) {
    tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo128);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tS = 0, tdA = 64, tdQ = 128;

  const int my_start = ti < T ? seq_off[find_seq(seq_off, B, ti)] : INT_MAX;
  const int w_min_start = warp_min_i(my_start);
  const int w_max_t = min(q0 + warp * 32 + 31, T - 1);
  const int kt_first = seq_off[find_seq(seq_off, B, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * TILE);
    tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
    tma_load_2d(sdO, &mapdo128, bar_q, h * DH, q0);
  }
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);
  uint32_t ph = 0;
  for (int kt = kt_first; kt <= kt_last; ++kt, ph ^= 1u) {
    const int k0 = kt * 64;
    if (tid == 0) {
      if (kt > kt_first) mbar_wait(bar_o, ph ^ 1u);
      mbar_arrive_expect_tx(bar_kv, TILE);
      tma_load_2d(sK, &map64, bar_kv, 3 * D + h * DH, k0);
      tma_load_2d(sV, &map64, bar_kv, 1 * D + h * DH, k0);
    }
    if (tid < 64) s_kvalid[tid] = (k0 + tid < T) ? key_valid[k0 + tid] : 0;
    if (tid == 0) {
      if (kt == kt_first) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + tS, C::desc_k(sQ, ks), C::desc_k(sK, ks), idesc_s, ks > 0);      // S  = Q K^T
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + tdA, C::desc_k(sdO, ks), C::desc_k(sV, ks), idesc_s, ks > 0);    // dA = dO V^T
      umma_commit(bar_s);
    }
    __syncthreads();
    mbar_wait(bar_s, ph);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      if (k0 + c * 32 > w_max_t || k0 + c * 32 + 31 < w_min_start) {  // warp-uniform: chunk fully masked
        put_zero32_sw128(sdS, 128, tid, c * 32);
        continue;
      }
      float s[32], da[32];
      tmem_ld_32x32(t_lane + tS + c * 32, s);
      tmem_ld_32x32(t_lane + tdA + c * 32, da);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int tj = k0 + c * 32 + e;
        const bool keep = (tj <= ti) && (tj >= my_start) && s_kvalid[c * 32 + e];
        s[e] = keep ? da[e] * inv_n * silu_grad_fast_f(s[e]) : 0.f;
      }
      put_row32_sw128(sdS, 128, tid, c * 32, s);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdQ, desc_p(sdS, ks), C::desc_mn(sK, ks), idesc_q, (kt > kt_first || ks > 0));  // dQ += dS K
      umma_commit(bar_o);
    }
  }
  mbar_wait(bar_o, ph ^ 1u);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < DH / 32; ++c) {
    float v[32];
    tmem_ld_32x32(t_lane + tdQ + c * 32, v);
    if (ti < T) {
      const int64_t off = (int64_t)ti * ld + h * DH + c * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float p4[4], o4[4];
        load4<bf16>(pre_q + off + 4 * e, p4);
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = v[4 * e + k] * silu_grad_fast_f(p4[k]);
        store4<bf16>(d_pre_q + off + 4 * e, o4);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// --------------------------------------------------------------------------------------- backward: dK, dV
template <int DH>
__global__ void __launch_bounds__(128)
attn_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                       const __grid_constant__ CUtensorMap mapdo64, const int32_t* __restrict__ seq_off, int B,
                       const uint8_t* __restrict__ key_valid, int T, int D, float inv_n,
                       const bf16* __restrict__ pre_k, const bf16* __restrict__ pre_v, int64_t ld,
                       bf16* __restrict__ d_pre_k, bf16* __restrict__ d_pre_v) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + TILE, sQ = sV + TILE, sdO = sQ + TILE / 2, sPT = sdO + TILE / 2,
                 sdST = sPT + 16384;
  const uint32_t bars = sdST + 16384;
  const uint32_t bar_kv = bars, bar_q = bars + 8, bar_s = bars + 16, bar_o = bars + 24, tslot = bars + 32;
  __shared__ int s_qstart[64];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int k0 = blockIdx.x * 128, h = blockIdx.y;
  const int tj = k0 + tid;
  if (tid == 0) {
    tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo64);
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tS = 0, tdA = 64, tdK = 128, tdV = 192;

  const bool kv_ok = tj < T && key_valid[tj] != 0;
  const int k_last = min(k0 + 127, T - 1);
  const int q_hi = seq_off[find_seq(seq_off, B, k_last) + 1] - 1;   // last token of the last key's sequence
  const int qt_first = k0 >> 6, qt_last = q_hi >> 6;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * TILE);
    tma_load_2d(sK, &map128, bar_kv, 3 * D + h * DH, k0);
    tma_load_2d(sV, &map128, bar_kv, 1 * D + h * DH, k0);
  }
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);
  uint32_t ph = 0;
  for (int qt = qt_first; qt <= qt_last; ++qt, ph ^= 1u) {
    const int i0 = qt * 64;
    if (tid == 0) {
      if (qt > qt_first) mbar_wait(bar_o, ph ^ 1u);
      mbar_arrive_expect_tx(bar_q, TILE);
      tma_load_2d(sQ, &map64, bar_q, 2 * D + h * DH, i0);
      tma_load_2d(sdO, &mapdo64, bar_q, h * DH, i0);
    }
    if (tid < 64) s_qstart[tid] = (i0 + tid < T) ? seq_off[find_seq(seq_off, B, i0 + tid)] : INT_MAX;
    if (tid == 0) {
      if (qt == qt_first) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + tS, C::desc_k(sK, ks), C::desc_k(sQ, ks), idesc_s, ks > 0);      // S^T  = K Q^T
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + tdA, C::desc_k(sV, ks), C::desc_k(sdO, ks), idesc_s, ks > 0);    // dA^T = V dO^T
      umma_commit(bar_s);
    }
    __syncthreads();
    mbar_wait(bar_s, ph);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      if (i0 + c * 32 + 31 < k0 + warp * 32) {   // warp-uniform: every query of the chunk precedes every key
        put_zero32_sw128(sPT, 128, tid, c * 32);
        put_zero32_sw128(sdST, 128, tid, c * 32);
        continue;
      }
      float s[32], da[32];
      tmem_ld_32x32(t_lane + tS + c * 32, s);
      tmem_ld_32x32(t_lane + tdA + c * 32, da);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int tq = i0 + c * 32 + e;
        const bool keep = kv_ok && (tj <= tq) && (tj >= s_qstart[c * 32 + e]);
        const float sv = s[e];
        const float sg = sigmoid_fast_f(sv);                          // ONE MUFU for both P^T and dS^T
        s[e] = keep ? sv * sg * inv_n : 0.f;                          // P^T
        da[e] = keep ? da[e] * inv_n * sg * fmaf(sv, 1.f - sg, 1.f) : 0.f;   // dS^T
      }
      put_row32_sw128(sPT, 128, tid, c * 32, s);
      put_row32_sw128(sdST, 128, tid, c * 32, da);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdV, desc_p(sPT, ks), C::desc_mn(sdO, ks), idesc_g, (qt > qt_first || ks > 0));  // dV += P^T dO
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdK, desc_p(sdST, ks), C::desc_mn(sQ, ks), idesc_g, (qt > qt_first || ks > 0));  // dK += dS^T Q
      umma_commit(bar_o);
    }
  }
  mbar_wait(bar_o, ph ^ 1u);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < DH / 32; ++c) {
    float gk[32], gv[32];
    tmem_ld_32x32(t_lane + tdK + c * 32, gk);
    tmem_ld_32x32(t_lane + tdV + c * 32, gv);
    if (tj < T) {
      const int64_t off = (int64_t)tj * ld + h * DH + c * 32;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float p4[4], o4[4];
        load4<bf16>(pre_k + off + 4 * e, p4);
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = gk[4 * e + k] * silu_grad_fast_f(p4[k]);
        store4<bf16>(d_pre_k + off + 4 * e, o4);
        load4<bf16>(pre_v + off + 4 * e, p4);
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = gv[4 * e + k] * silu_grad_fast_f(p4[k]);
        store4<bf16>(d_pre_v + off + 4 * e, o4);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// --------------------------------------------------------------------------------------- backward, pipelined
// Both backward passes use the same warp-specialised structure as the forward (warps 0..7 pointwise, warp 8 issuer, no
// block barrier in the loop).  S and dA are single-buffered (TMEM: S | dA | gradient accumulators = 256 columns, two
// CTAs per SM), so inside one CTA the pointwise phase and the S / dA MMAs of the next tile alternate; the second CTA
// of the SM fills the gaps.  The next tile's S / dA MMAs are issued back to back with this tile's gradient MMAs, ONE
// commit covers them all: a single wait at the top of the next iteration also proves that the dS / P buffers and the
// oldest operand buffer are free.  Operand tiles (K, V for the dQ pass; Q, dO for the dK/dV pass) are prefetched into
// shared-memory rings >= 1 iteration ahead; unmasked chunks skip the compares; 1/n_pad is applied once per output row.
template <int DH>
__global__ void __launch_bounds__(AT_THREADS)
attn_tc_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                           const __grid_constant__ CUtensorMap mapdo128, const int32_t* __restrict__ seq_off, int B,
                           const uint8_t* __restrict__ key_valid, int T, int D, float inv_n,
                           const bf16* __restrict__ pre_q, int64_t ld, bf16* __restrict__ d_pre_q) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ, HT = TILE / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + TILE, sK0 = sdO + TILE, sV0 = sK0 + 3 * HT, sdS = sV0 + 3 * HT;
  const uint32_t bars = sdS + 16384;
  const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bars + 32, bar_s = bars + 56, bar_g = bars + 64,
                 bar_p = bars + 72, tslot = bars + 80;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < 3; ++i) mbar_init(bar_k0 + 8 * i, 1);
    for (int i = 0; i < 3; ++i) mbar_init(bar_v0 + 8 * i, 1);
    mbar_init(bar_s, 1); mbar_init(bar_g, 1);
    mbar_init(bar_p, 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tslot, 256);
  __shared__ int32_t s_seq[AT_SEQ_SMEM];
  const SeqTab seqs = seq_tab_load(s_seq, seq_off, B);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int kt_first = seqs.p[find_seq(seqs, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  const int n_it = kt_last - kt_first + 1;
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t tS = 0, tdA = 64, tdQ = 128;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo128);
      const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);
      auto load_k = [&](int it) {
        const int j = it % 3;
        mbar_arrive_expect_tx(bar_k0 + 8 * j, HT);
        tma_load_2d(sK0 + j * HT, &map64, bar_k0 + 8 * j, 3 * D + h * DH, (kt_first + it) * 64);
      };
      auto load_v = [&](int it) {
        const int j = it % 3;
        mbar_arrive_expect_tx(bar_v0 + 8 * j, HT);
        tma_load_2d(sV0 + j * HT, &map64, bar_v0 + 8 * j, 1 * D + h * DH, (kt_first + it) * 64);
      };
      auto issue_s_da = [&](int it) {   // S(it) = Q K^T, dA(it) = dO V^T
        mbar_wait(bar_k0 + 8 * (it % 3), (uint32_t)(it / 3) & 1u);
        mbar_wait(bar_v0 + 8 * (it % 3), (uint32_t)(it / 3) & 1u);
        tc_fence_after();
        const uint32_t k = sK0 + (uint32_t)(it % 3) * HT, v = sV0 + (uint32_t)(it % 3) * HT;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tS, C::desc_k(sQ, ks), C::desc_k(k, ks), idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tdA, C::desc_k(sdO, ks), C::desc_k(v, ks), idesc_s, ks > 0);
      };
      mbar_arrive_expect_tx(bar_q, 2 * TILE);
      tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
      tma_load_2d(sdO, &mapdo128, bar_q, h * DH, q0);
      for (int i = 0; i < 3 && i < n_it; ++i) { load_k(i); load_v(i); }
      mbar_wait(bar_q, 0);
      issue_s_da(0);
      umma_commit(bar_s);
      for (int it = 0; it < n_it; ++it) {
        // dS(it) written; S / dA(it) consumed.  The pointwise threads waited for bar_g(it-1) (dQ(it-1) retired) before
        // they wrote dS(it): K(it-1)'s ring slot is free; V(it)'s slot is free since dA(it) retired (bar_s(it))
        mbar_wait(bar_p, (uint32_t)it & 1u);
        tc_fence_after();
        if (it >= 1 && it + 2 < n_it) load_k(it + 2);
        if (it + 3 < n_it) load_v(it + 3);
        // next tile's S / dA FIRST (the pointwise threads are waiting for them), this tile's dQ behind them: it retires
        // under the next pointwise phase
        if (it + 1 < n_it) {
          issue_s_da(it + 1);
          umma_commit(bar_s);
        }
        const uint32_t k = sK0 + (uint32_t)(it % 3) * HT;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + tdQ, desc_p(sdS, ks), C::desc_mn(k, ks), idesc_q, (it > 0 || ks > 0));      // dQ += dS K
        umma_commit(bar_g);
      }
    }
  } else {
    const int quarter = warp & 3, colhalf = warp >> 2;
    const int row = quarter * 32 + lane;
    const int ti = q0 + row;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    const int my_start = ti < T ? seqs.p[find_seq(seqs, ti)] : INT_MAX;
    const int w_min_start = warp_min_i(my_start);
    const int w_max_start = warp_max_i(ti < T ? my_start : INT_MIN);
    const int w_min_t = q0 + quarter * 32;
    const int w_max_t = min(q0 + quarter * 32 + 31, T - 1);
    const bool w_all_rows = q0 + quarter * 32 + 31 < T;
    for (int it = 0; it < n_it; ++it) {
      const int c0 = (kt_first + it) * 64 + colhalf * 32;
      const bool skip = c0 > w_max_t || c0 + 31 < w_min_start;   // warp-uniform: chunk fully masked
      const uint32_t kvb = skip ? 0u : kvalid_bits(key_valid, c0, lane, T);
      mbar_wait(bar_s, (uint32_t)it & 1u);        // S(it), dA(it) retired
      tc_fence_after();
      uint32_t pk[16];
      if (skip) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      } else {
        float sv[32], da[32];
        tmem_ld_32x32(t_lane + tS + colhalf * 32, sv);
        tmem_ld_32x32(t_lane + tdA + colhalf * 32, da);
        if (w_all_rows && c0 + 31 <= w_min_t && c0 >= w_max_start && kvb == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(da[2 * e] * silu_grad_fast_f(sv[2 * e]),
                                                      da[2 * e + 1] * silu_grad_fast_f(sv[2 * e + 1]));
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float a[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int tj = c0 + 2 * e + u;
              const bool keep = (tj <= ti) && (tj >= my_start) && ((kvb >> (2 * e + u)) & 1u);
              a[u] = keep ? da[2 * e + u] * silu_grad_fast_f(sv[2 * e + u]) : 0.f;
            }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a[0], a[1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
      }
      if (it > 0) mbar_wait(bar_g, (uint32_t)(it - 1) & 1u);    // dQ(it-1) retired: sdS is free
      put_pk32_sw128(sdS, row, colhalf, pk);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_g, (uint32_t)(n_it - 1) & 1u);
    tc_fence_after();
    constexpr int HALF = DH / 2;
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_lane + tdQ + colhalf * HALF + c * 16, r);
      if (ti < T) {
        const int64_t off = (int64_t)ti * ld + h * DH + colhalf * HALF + c * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float p4[4], o4[4];
          load4<bf16>(pre_q + off + 4 * e, p4);
#pragma unroll
          for (int k = 0; k < 4; ++k) o4[k] = __uint_as_float(r[4 * e + k]) * inv_n * silu_grad_fast_f(p4[k]);
          store4<bf16>(d_pre_q + off + 4 * e, o4);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <int DH>
__global__ void __launch_bounds__(AT_THREADS)
attn_tc_bwd_dkv_pipe_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                            const __grid_constant__ CUtensorMap mapdo64, const int32_t* __restrict__ seq_off, int B,
                            const uint8_t* __restrict__ key_valid, int T, int D, float inv_n,
                            const bf16* __restrict__ pre_k, const bf16* __restrict__ pre_v, int64_t ld,
                            bf16* __restrict__ d_pre_k, bf16* __restrict__ d_pre_v) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ, HT = TILE / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + TILE, sQ0 = sV + TILE, sdO0 = sQ0 + 2 * HT, sPT = sdO0 + 2 * HT, sdST = sPT + 16384;
  const uint32_t bars = sdST + 16384;
  const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_s = bars + 24, bar_g = bars + 32, bar_p = bars + 40,
                 bar_m0 = bars + 48, tslot = bars + 64;
  __shared__ int s_qstart[2][64];    // sequence start of every query of the tile (INT_MAX beyond T), by the issuer warp
  __shared__ int s_qsmax[2][2];      // per 32-query chunk: the latest of them
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y;
  if (tid == 0) {
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_q0 + 8 * i, 1); mbar_init(bar_m0 + 8 * i, 32); }
    mbar_init(bar_s, 1); mbar_init(bar_g, 1);
    mbar_init(bar_p, 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tslot, 256);
  __shared__ int32_t s_seq[AT_SEQ_SMEM];
  const SeqTab seqs = seq_tab_load(s_seq, seq_off, B);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int k_last = min(k0 + 127, T - 1);
  const int q_hi = seqs.p[find_seq(seqs, k_last) + 1] - 1;   // last token of the last key's sequence
  const int qt_first = k0 >> 6, qt_last = q_hi >> 6;
  const int n_it = qt_last - qt_first + 1;
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t tS = 0, tdA = 64, tdK = 128, tdV = 192;

  if (warp == 8) {
    // every lane: query metadata of a tile (two queries per lane) -> s_qstart / s_qsmax, published through bar_m
    auto meta = [&](int it) {
      const int buf = it & 1, i0 = (qt_first + it) * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int q = i0 + half * 32 + lane;
        const int qs = q < T ? seqs.p[find_seq(seqs, q)] : INT_MAX;
        s_qstart[buf][half * 32 + lane] = qs;
        const int mx = warp_max_i(qs);
        if (lane == 0) s_qsmax[buf][half] = mx;
      }
      mbar_arrive(bar_m0 + 8 * buf);
    };
    const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);
    auto issue_s_da = [&](int it) {   // S^T(it) = K Q^T, dA^T(it) = V dO^T   (lane 0)
      mbar_wait(bar_q0 + 8 * (it & 1), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t q = sQ0 + (uint32_t)(it & 1) * HT, g = sdO0 + (uint32_t)(it & 1) * HT;
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tS, C::desc_k(sK, ks), C::desc_k(q, ks), idesc_s, ks > 0);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tdA, C::desc_k(sV, ks), C::desc_k(g, ks), idesc_s, ks > 0);
    };
    auto load_q_do = [&](int it) {
      const int j = it & 1, i0 = (qt_first + it) * 64;
      mbar_arrive_expect_tx(bar_q0 + 8 * j, TILE);
      tma_load_2d(sQ0 + j * HT, &map64, bar_q0 + 8 * j, 2 * D + h * DH, i0);
      tma_load_2d(sdO0 + j * HT, &mapdo64, bar_q0 + 8 * j, h * DH, i0);
    };
    if (lane == 0) {
      tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo64);
      mbar_arrive_expect_tx(bar_kv, 2 * TILE);
      tma_load_2d(sK, &map128, bar_kv, 3 * D + h * DH, k0);
      tma_load_2d(sV, &map128, bar_kv, 1 * D + h * DH, k0);
      load_q_do(0);
      if (n_it > 1) load_q_do(1);
    }
    __syncwarp();
    meta(0);
    if (n_it > 1) meta(1);
    if (lane == 0) {
      mbar_wait(bar_kv, 0);
      issue_s_da(0);
      umma_commit(bar_s);
    }
    __syncwarp();
    for (int it = 0; it < n_it; ++it) {
      // P^T / dS^T(it) written; S^T / dA^T(it) consumed.  The pointwise threads waited for bar_g(it-1) (dV / dK(it-1)
      // retired) before writing: the Q / dO slot of tile it-1 is free; the metadata slot of tile it is free
      mbar_wait(bar_p, (uint32_t)it & 1u);
      __syncwarp();
      if (lane == 0) {
        tc_fence_after();
        if (it >= 1 && it + 1 < n_it) load_q_do(it + 1);
        if (it + 1 < n_it) {                       // next tile's S^T / dA^T first, this tile's dV / dK behind them
          issue_s_da(it + 1);
          umma_commit(bar_s);
        }
        const uint32_t q = sQ0 + (uint32_t)(it & 1) * HT, g = sdO0 + (uint32_t)(it & 1) * HT;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + tdV, desc_p(sPT, ks), C::desc_mn(g, ks), idesc_g, (it > 0 || ks > 0));   // dV += P^T dO
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem + tdK, desc_p(sdST, ks), C::desc_mn(q, ks), idesc_g, (it > 0 || ks > 0));  // dK += dS^T Q
        umma_commit(bar_g);
      }
      __syncwarp();
      if (it + 2 < n_it) meta(it + 2);              // slot of tile it: all its readers have arrived on bar_p(it)
    }
  } else {
    const int quarter = warp & 3, colhalf = warp >> 2;
    const int row = quarter * 32 + lane;
    const int tj = k0 + row;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    const bool kv_ok = tj < T && key_valid[tj] != 0;
    const bool w_kv_all = __all_sync(0xffffffffu, kv_ok);
    const int w_key_lo = k0 + quarter * 32, w_key_hi = k0 + quarter * 32 + 31;
    for (int it = 0; it < n_it; ++it) {
      const int b = it & 1;
      const int c0 = (qt_first + it) * 64 + colhalf * 32;            // first query of this thread's chunk
      mbar_wait(bar_m0 + 8 * b, (uint32_t)(it >> 1) & 1u);
      mbar_wait(bar_s, (uint32_t)it & 1u);        // S^T(it), dA^T(it) retired
      tc_fence_after();
      uint32_t pp[16], pd[16];
      if (c0 + 31 < w_key_lo) {                    // warp-uniform: every query of the chunk precedes every key
#pragma unroll
        for (int e = 0; e < 16; ++e) { pp[e] = 0u; pd[e] = 0u; }
      } else {
        float sv[32], da[32];
        tmem_ld_32x32(t_lane + tS + colhalf * 32, sv);
        tmem_ld_32x32(t_lane + tdA + colhalf * 32, da);
        const bool fast = w_kv_all && c0 >= w_key_hi && s_qsmax[b][colhalf] <= w_key_lo;
        if (fast) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float p2[2], d2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float x = sv[2 * e + u];
              const float sg = sigmoid_fast_f(x);                      // ONE MUFU for both P^T and dS^T
              p2[u] = x * sg;
              d2[u] = da[2 * e + u] * sg * fmaf(x, 1.f - sg, 1.f);
            }
            __nv_bfloat162 hp = __floats2bfloat162_rn(p2[0], p2[1]);
            __nv_bfloat162 hd = __floats2bfloat162_rn(d2[0], d2[1]);
            pp[e] = *reinterpret_cast<uint32_t*>(&hp);
            pd[e] = *reinterpret_cast<uint32_t*>(&hd);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float p2[2], d2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int tq = c0 + 2 * e + u;
              const bool keep = kv_ok && (tj <= tq) && (tj >= s_qstart[b][colhalf * 32 + 2 * e + u]);
              const float x = sv[2 * e + u];
              const float sg = sigmoid_fast_f(x);
              p2[u] = keep ? x * sg : 0.f;
              d2[u] = keep ? da[2 * e + u] * sg * fmaf(x, 1.f - sg, 1.f) : 0.f;
            }
            __nv_bfloat162 hp = __floats2bfloat162_rn(p2[0], p2[1]);
            __nv_bfloat162 hd = __floats2bfloat162_rn(d2[0], d2[1]);
            pp[e] = *reinterpret_cast<uint32_t*>(&hp);
            pd[e] = *reinterpret_cast<uint32_t*>(&hd);
          }
        }
      }
      if (it > 0) mbar_wait(bar_g, (uint32_t)(it - 1) & 1u);    // dV / dK(it-1) retired: sPT / sdST are free
      put_pk32_sw128(sPT, row, colhalf, pp);
      put_pk32_sw128(sdST, row, colhalf, pd);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_g, (uint32_t)(n_it - 1) & 1u);
    tc_fence_after();
    constexpr int HALF = DH / 2;
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t gk[16], gv[16];
      tmem_ld_32x32b_x16(t_lane + tdK + colhalf * HALF + c * 16, gk);
      tmem_ld_32x32b_x16(t_lane + tdV + colhalf * HALF + c * 16, gv);
      if (tj < T) {
        const int64_t off = (int64_t)tj * ld + h * DH + colhalf * HALF + c * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float p4[4], o4[4];
          load4<bf16>(pre_k + off + 4 * e, p4);
#pragma unroll
          for (int k = 0; k < 4; ++k) o4[k] = __uint_as_float(gk[4 * e + k]) * inv_n * silu_grad_fast_f(p4[k]);
          store4<bf16>(d_pre_k + off + 4 * e, o4);
          load4<bf16>(pre_v + off + 4 * e, p4);
#pragma unroll
          for (int k = 0; k < 4; ++k) o4[k] = __uint_as_float(gv[4 * e + k]) * inv_n * silu_grad_fast_f(p4[k]);
          store4<bf16>(d_pre_v + off + 4 * e, o4);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// --------------------------------------------------------------------------------------- host
template <int DH>
static int attn_tc_fwd_launch(const bf16* act_base, int64_t ld, const int32_t* seq_off, int B, const uint8_t* key_valid,
                              int T, int n_heads, float inv_n, float* out, cudaStream_t st) {
  const int D = n_heads * DH;
  CUtensorMap m128, m64;
  if (b200_make_map_bf16(&m128, act_base, (uint64_t)ld, (uint64_t)T, (uint64_t)ld, DH, 128, DH * 2)) return 1;
  if (b200_make_map_bf16(&m64, act_base, (uint64_t)ld, (uint64_t)T, (uint64_t)ld, DH, 64, DH * 2)) return 1;
  dim3 grid(ceil_div_i(T, 128), n_heads);
  static int use_pipe = -1;
  if (use_pipe < 0) {
    const char* e = getenv("B200REC_ATTN_PIPE");
    use_pipe = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (use_pipe) {
    size_t smem_p = 128 * DH * 2 + 6 * 64 * DH * 2 + 2 * 16384 + 256 + 1024;
    { static bool once_p = false; if (!once_p) { B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); once_p = true; } }
    attn_tc_fwd_pipe_kernel<DH><<<grid, AT_THREADS, smem_p, st>>>(m128, m64, seq_off, B, key_valid, T, D, inv_n, out);
    B200_LAUNCH_OK();
    return 0;
  }
  size_t smem = 2 * 128 * DH * 2 + 16384 + 256 + 1024;
  { static bool once_1 = false; if (!once_1) { B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once_1 = true; } }
  attn_tc_fwd_kernel<DH><<<grid, 128, smem, st>>>(m128, m64, seq_off, B, key_valid, T, D, inv_n, out);
  B200_LAUNCH_OK();
  return 0;
}

template <int DH>
static int attn_tc_bwd_launch(const bf16* act_base, const bf16* pre_base, int64_t ld, const int32_t* seq_off, int B,
                              const uint8_t* key_valid, int T, int n_heads, float inv_n, const bf16* d_out,
                              bf16* d_pre_base, cudaStream_t st) {
  const int D = n_heads * DH;
  CUtensorMap m128, m64, do128, do64;
  if (b200_make_map_bf16(&m128, act_base, (uint64_t)ld, (uint64_t)T, (uint64_t)ld, DH, 128, DH * 2)) return 1;
  if (b200_make_map_bf16(&m64, act_base, (uint64_t)ld, (uint64_t)T, (uint64_t)ld, DH, 64, DH * 2)) return 1;
  if (b200_make_map_bf16(&do128, d_out, (uint64_t)D, (uint64_t)T, (uint64_t)D, DH, 128, DH * 2)) return 1;
  if (b200_make_map_bf16(&do64, d_out, (uint64_t)D, (uint64_t)T, (uint64_t)D, DH, 64, DH * 2)) return 1;
  size_t smem_q = 3 * 128 * DH * 2 + 16384 + 256 + 1024;
  size_t smem_kv = 3 * 128 * DH * 2 + 32768 + 256 + 1024;
  { static bool once_2 = false; if (!once_2) { B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_q)); once_2 = true; } }
  { static bool once_3 = false; if (!once_3) { B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_kv)); once_3 = true; } }
  dim3 grid(ceil_div_i(T, 128), n_heads);
  static int use_pipe = -1;
  if (use_pipe < 0) {
    const char* e = getenv("B200REC_ATTN_PIPE");
    use_pipe = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (use_pipe) {
    size_t smem_pq = 2 * 128 * DH * 2 + 6 * 64 * DH * 2 + 16384 + 256 + 1024;
    size_t smem_pkv = 2 * 128 * DH * 2 + 4 * 64 * DH * 2 + 32768 + 256 + 1024;
    { static bool once_p = false; if (!once_p) {
        B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dq_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pq));
        B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dkv_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pkv));
        once_p = true; } }
    attn_tc_bwd_dq_pipe_kernel<DH><<<grid, AT_THREADS, smem_pq, st>>>(m128, m64, do128, seq_off, B, key_valid, T, D, inv_n,
                                                               pre_base + 2 * D, ld, d_pre_base + 2 * D);
    attn_tc_bwd_dkv_pipe_kernel<DH><<<grid, AT_THREADS, smem_pkv, st>>>(m128, m64, do64, seq_off, B, key_valid, T, D, inv_n,
                                                                 pre_base + 3 * D, pre_base + 1 * D, ld,
                                                                 d_pre_base + 3 * D, d_pre_base + 1 * D);
    B200_LAUNCH_OK();
    return 0;
  }
  // column slices of the [T, 4D] buffers: u | v | q | k
  attn_tc_bwd_dq_kernel<DH><<<grid, 128, smem_q, st>>>(m128, m64, do128, seq_off, B, key_valid, T, D, inv_n,
                                                       pre_base + 2 * D, ld, d_pre_base + 2 * D);
  attn_tc_bwd_dkv_kernel<DH><<<grid, 128, smem_kv, st>>>(m128, m64, do64, seq_off, B, key_valid, T, D, inv_n,
                                                         pre_base + 3 * D, pre_base + 1 * D, ld, d_pre_base + 3 * D,
                                                         d_pre_base + 1 * D);
  B200_LAUNCH_OK();
  return 0;
}

// act / pre / d_pre point at column 0 of the [T, 4D] uvqk buffers (bf16, leading dimension ld = 4D).
extern "C" int b200rec_hstu_attn_tc_fwd(const void* act, int ld, const int32_t* seq_off, const uint8_t* key_valid,
                                        int B, int T, int n_heads, int dh, float inv_n, float* out, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld == 4 * n_heads * dh, "attn_tc: ld must be 4*D");
  B200_CHECK_ARG(((uintptr_t)act & 15) == 0, "attn_tc: act must be 16-byte aligned");
  if (dh == 64)
    return attn_tc_fwd_launch<64>((const bf16*)act, ld, seq_off, B, key_valid, T, n_heads, inv_n, out,
                                  (cudaStream_t)stream);
  if (dh == 32)
    return attn_tc_fwd_launch<32>((const bf16*)act, ld, seq_off, B, key_valid, T, n_heads, inv_n, out,
                                  (cudaStream_t)stream);
  b200rec_set_error("attn_tc: head dim %d not supported (32 or 64)", dh);
  return 1;
}

extern "C" int b200rec_hstu_attn_tc_bwd(const void* act, const void* pre, int ld, const int32_t* seq_off,
                                        const uint8_t* key_valid, int B, int T, int n_heads, int dh, float inv_n,
                                        const void* d_out, void* d_pre, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld == 4 * n_heads * dh, "attn_tc: ld must be 4*D");
  B200_CHECK_ARG(((uintptr_t)act & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "attn_tc: 16-byte alignment");
  if (dh == 64)
    return attn_tc_bwd_launch<64>((const bf16*)act, (const bf16*)pre, ld, seq_off, B, key_valid, T, n_heads, inv_n,
                                  (const bf16*)d_out, (bf16*)d_pre, (cudaStream_t)stream);
  if (dh == 32)
    return attn_tc_bwd_launch<32>((const bf16*)act, (const bf16*)pre, ld, seq_off, B, key_valid, T, n_heads, inv_n,
                                  (const bf16*)d_out, (bf16*)d_pre, (cudaStream_t)stream);
  b200rec_set_error("attn_tc: head dim %d not supported (32 or 64)", dh);
  return 1;
}
