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


// --------------------------------------------------------------------------------------- forward, pipelined
// Same tiling (128 queries x 64-key tiles), but the per-key-tile chain  TMA -> S = Q K^T -> SiLU -> P -> O += P V  is
// software-pipelined inside the CTA instead of relying on co-resident CTAs alone:
//   * K / V double-buffered in shared memory: the TMA of tile it+2 (K) / it+1 (V) is issued as soon as the MMA that read
//     the buffer has retired, a full iteration before it is needed;
//   * TWO S accumulators in TMEM: S(it+1) is issued at the START of iteration it, so the tensor core computes it while
//     the 256 threads run the SiLU of S(it); PV(it) is issued at the end of the iteration and retires under SiLU(it+1);
//   * 256 threads, two per query row (each takes 32 of the 64 keys): half the dependent chain per thread;
//   * chunks with no masked element (interior of a sequence, keys all valid) skip every compare / select;
//   * 1/n_pad is applied once to the output row instead of to every probability.
// The critical path of an iteration is tcgen05.ld -> SiLU -> st.shared -> barrier; the pointwise SiLU (one MUFU per
// score against 256 tensor flops, dh = 64) bounds the kernel at 1/2 of the tensor peak.
// 64 KB of shared memory and 256 TMEM columns (S0 | S1 | O) per CTA: two CTAs per SM.
template <int DH>
__global__ void __launch_bounds__(256)
attn_tc_fwd_pipe_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                        const int32_t* __restrict__ seq_off, int B, const uint8_t* __restrict__ key_valid, int T, int D,
                        float inv_n, float* __restrict__ out) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ;      // Q tile bytes; a 64-row K or V tile is TILE / 2
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK0 = sQ + TILE, sV0 = sK0 + TILE, sP = sV0 + TILE;   // sK0 / sV0: two TILE/2 buffers each
  const uint32_t bars = sP + 16384;
  const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bars + 24, bar_s0 = bars + 40, bar_o = bars + 56,
                 tslot = bars + 64;
  __shared__ uint8_t s_kvalid[2][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, colhalf = warp >> 2;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  const int row = quarter * 32 + lane;
  const int ti = q0 + row;
  if (tid == 0) {
    tma_prefetch_desc(&map128);
    tma_prefetch_desc(&map64);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_k0 + 8 * i, 1); mbar_init(bar_v0 + 8 * i, 1); mbar_init(bar_s0 + 8 * i, 1); }
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  const int kt_first = seq_off[find_seq(seq_off, B, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  const int n_it = kt_last - kt_first + 1;
  if (tid < 64) {
    const int kk = kt_first * 64 + tid;
    s_kvalid[0][tid] = kk < T ? key_valid[kk] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t tO = 128;  // TMEM columns: S0 = 0, S1 = 64, O = 128

  const int my_start = ti < T ? seq_off[find_seq(seq_off, B, ti)] : INT_MAX;
  const int w_min_start = warp_min_i(my_start);                       // earliest key any row of this warp may see
  const int w_max_start = warp_max_i(ti < T ? my_start : INT_MIN);    // latest sequence start among the warp's rows
  const int w_min_t = q0 + quarter * 32;                              // first row of the warp
  const int w_max_t = min(q0 + quarter * 32 + 31, T - 1);             // latest key (causal) any row may see
  const bool w_all_rows = q0 + quarter * 32 + 31 < T;
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, TILE);
    tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
    for (int i = 0; i < 2 && i < n_it; ++i) {
      mbar_arrive_expect_tx(bar_k0 + 8 * i, TILE / 2);
      tma_load_2d(sK0 + i * (TILE / 2), &map64, bar_k0 + 8 * i, 3 * D + h * DH, (kt_first + i) * 64);
      mbar_arrive_expect_tx(bar_v0 + 8 * i, TILE / 2);
      tma_load_2d(sV0 + i * (TILE / 2), &map64, bar_v0 + 8 * i, 1 * D + h * DH, (kt_first + i) * 64);
    }
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k0, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + 0, C::desc_k(sQ, ks), C::desc_k(sK0, ks), idesc_s, ks > 0);
    umma_commit(bar_s0);
  }
  for (int it = 0; it < n_it; ++it) {
    const int b = it & 1;
    const uint32_t use_ph = (uint32_t)(it >> 1) & 1u;          // parity of this use of the K / V / S buffers b
    const int k0 = (kt_first + it) * 64;
    if (tid == 0 && it + 1 < n_it) {
      // S(it+1) into the other accumulator: it was last read by SiLU(it-1), which every thread finished before the
      // barrier that ended iteration it-1
      mbar_wait(bar_k0 + 8 * (b ^ 1), (uint32_t)((it + 1) >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16(tmem + 64 * (b ^ 1), C::desc_k(sQ, ks), C::desc_k(sK0 + (b ^ 1) * (TILE / 2), ks), idesc_s, ks > 0);
      umma_commit(bar_s0 + 8 * (b ^ 1));
    }
    mbar_wait(bar_s0 + 8 * b, use_ph);
    tc_fence_after();
    if (tid == 0 && it + 2 < n_it) {                             // S(it) has retired: its K buffer is free
      mbar_arrive_expect_tx(bar_k0 + 8 * b, TILE / 2);
      tma_load_2d(sK0 + b * (TILE / 2), &map64, bar_k0 + 8 * b, 3 * D + h * DH, k0 + 128);
    }
    const int c0 = k0 + colhalf * 32;
    uint32_t pk[16];
    __syncwarp();                                                // warp 0: lane 0 rejoins before the aligned TMEM load
    const bool skip = c0 > w_max_t || c0 + 31 < w_min_start;     // warp-uniform: whole chunk masked for these 32 rows
    if (skip) {
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = 0u;
    } else {
      float v[32];
      tmem_ld_32x32(t_lane + 64 * b + colhalf * 32, v);
      const bool kv_all = __all_sync(0xffffffffu, s_kvalid[b][colhalf * 32 + lane] != 0);
      if (w_all_rows && c0 + 31 <= w_min_t && c0 >= w_max_start && kv_all) {
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
            const bool keep = (tj <= ti) && (tj >= my_start) && s_kvalid[b][colhalf * 32 + 2 * e + u];
            a[u] = keep ? silu_fast_f(v[2 * e + u]) : 0.f;
          }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(a[0], a[1]);
          pk[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
      }
    }
    if (it > 0) {
      mbar_wait(bar_o, (uint32_t)(it - 1) & 1u);                 // PV(it-1) retired: sP and its V buffer are free
      if (tid == 0 && it + 1 < n_it) {
        mbar_arrive_expect_tx(bar_v0 + 8 * (b ^ 1), TILE / 2);
        tma_load_2d(sV0 + (b ^ 1) * (TILE / 2), &map64, bar_v0 + 8 * (b ^ 1), 1 * D + h * DH, k0 + 64);
      }
    }
    {
      // 32 keys of this row -> bf16 into the K-major SWIZZLE_128B P tile (128 rows x 64 keys = one 128-byte row each)
      const uint32_t rbase = sP + (uint32_t)row * 128u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t addr = rbase + (uint32_t)(((colhalf * 4 + q) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                     : "memory");
      }
    }
    if (tid < 64 && it + 1 < n_it) {
      const int kk = k0 + 64 + tid;
      s_kvalid[b ^ 1][tid] = kk < T ? key_valid[kk] : 0;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v0 + 8 * b, use_ph);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tO, desc_p(sP, ks), C::desc_mn(sV0 + b * (TILE / 2), ks), idesc_o, (it > 0 || ks > 0));
      umma_commit(bar_o);
    }
  }
  mbar_wait(bar_o, (uint32_t)(n_it - 1) & 1u);
  tc_fence_after();
  {
    // warp (quarter, colhalf) stores columns [colhalf * DH/2, +DH/2) of its 32 rows
    constexpr int HALF = DH / 2;
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(t_lane + tO + colhalf * HALF + c * 16));
      tmem_ld_wait();
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
  if (warp == 0) {
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
// Both backward passes keep the two-CTAs-per-SM ping-pong (one CTA's pointwise phase runs under the other's MMAs) and
// take the operand loads off the critical path: K / V (dQ pass) and Q / dO (dK,dV pass) tiles are prefetched into
// rotating shared-memory buffers a full iteration ahead, the next tile's S / dA MMAs are issued back to back with this
// tile's gradient MMA by the same elected thread (ONE commit covers all three, so one wait at the top of the next
// iteration also proves that the dS / P buffers and the oldest operand buffer are free), 256 threads share the rows two
// per row, unmasked chunks skip the compares and 1/n_pad is applied once per output row.
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

template <int DH>
__global__ void __launch_bounds__(256)
attn_tc_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                           const __grid_constant__ CUtensorMap mapdo128, const int32_t* __restrict__ seq_off, int B,
                           const uint8_t* __restrict__ key_valid, int T, int D, float inv_n,
                           const bf16* __restrict__ pre_q, int64_t ld, bf16* __restrict__ d_pre_q) {
  using C = AtCfg<DH>;
  constexpr uint32_t TILE = 128 * C::SWZ, HT = TILE / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + TILE, sK0 = sdO + TILE, sV0 = sK0 + 3 * HT, sdS = sV0 + 2 * HT;
  const uint32_t bars = sdS + 16384;
  const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bars + 32, bar_s = bars + 48, bar_fin = bars + 56,
                 tslot = bars + 64;
  __shared__ uint8_t s_kvalid[2][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, colhalf = warp >> 2;
  const int q0 = blockIdx.x * 128, h = blockIdx.y;
  const int row = quarter * 32 + lane;
  const int ti = q0 + row;
  if (tid == 0) {
    tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo128);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 3; ++i) mbar_init(bar_k0 + 8 * i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_v0 + 8 * i, 1);
    mbar_init(bar_s, 1); mbar_init(bar_fin, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  const int kt_first = seq_off[find_seq(seq_off, B, q0)] >> 6;
  const int kt_last = min(q0 + 127, T - 1) >> 6;
  const int n_it = kt_last - kt_first + 1;
  if (tid < 64) {
    const int kk = kt_first * 64 + tid;
    s_kvalid[0][tid] = kk < T ? key_valid[kk] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t tS = 0, tdA = 64, tdQ = 128;

  const int my_start = ti < T ? seq_off[find_seq(seq_off, B, ti)] : INT_MAX;
  const int w_min_start = warp_min_i(my_start);
  const int w_max_start = warp_max_i(ti < T ? my_start : INT_MIN);
  const int w_min_t = q0 + quarter * 32;
  const int w_max_t = min(q0 + quarter * 32 + 31, T - 1);
  const bool w_all_rows = q0 + quarter * 32 + 31 < T;
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);
  auto issue_s_da = [&](int it) {   // S(it) = Q K^T, dA(it) = dO V^T  (elected thread)
    const uint32_t k = sK0 + (uint32_t)(it % 3) * HT, v = sV0 + (uint32_t)(it & 1) * HT;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tS, C::desc_k(sQ, ks), C::desc_k(k, ks), idesc_s, ks > 0);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16(tmem + tdA, C::desc_k(sdO, ks), C::desc_k(v, ks), idesc_s, ks > 0);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * TILE);
    tma_load_2d(sQ, &map128, bar_q, 2 * D + h * DH, q0);
    tma_load_2d(sdO, &mapdo128, bar_q, h * DH, q0);
    for (int i = 0; i < 2 && i < n_it; ++i) {
      mbar_arrive_expect_tx(bar_k0 + 8 * i, HT);
      tma_load_2d(sK0 + i * HT, &map64, bar_k0 + 8 * i, 3 * D + h * DH, (kt_first + i) * 64);
      mbar_arrive_expect_tx(bar_v0 + 8 * i, HT);
      tma_load_2d(sV0 + i * HT, &map64, bar_v0 + 8 * i, 1 * D + h * DH, (kt_first + i) * 64);
    }
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k0, 0);
    mbar_wait(bar_v0, 0);
    tc_fence_after();
    issue_s_da(0);
    umma_commit(bar_s);
  }
  for (int it = 0; it < n_it; ++it) {
    const int b2 = it & 1;
    const int k0 = (kt_first + it) * 64;
    mbar_wait(bar_s, (uint32_t)it & 1u);        // S(it), dA(it) and dQ(it-1) retired
    tc_fence_after();
    if (tid == 0 && it + 2 < n_it) {
      const int j = (it + 2) % 3;                // held K(it-1): free
      mbar_arrive_expect_tx(bar_k0 + 8 * j, HT);
      tma_load_2d(sK0 + j * HT, &map64, bar_k0 + 8 * j, 3 * D + h * DH, k0 + 128);
      mbar_arrive_expect_tx(bar_v0 + 8 * b2, HT);   // held V(it): dA(it) retired
      tma_load_2d(sV0 + b2 * HT, &map64, bar_v0 + 8 * b2, 1 * D + h * DH, k0 + 128);
    }
    const int c0 = k0 + colhalf * 32;
    uint32_t pk[16];
    __syncwarp();
    if (c0 > w_max_t || c0 + 31 < w_min_start) {   // warp-uniform: chunk fully masked
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = 0u;
    } else {
      float sv[32], da[32];
      tmem_ld_32x32(t_lane + tS + colhalf * 32, sv);
      tmem_ld_32x32(t_lane + tdA + colhalf * 32, da);
      const bool kv_all = __all_sync(0xffffffffu, s_kvalid[b2][colhalf * 32 + lane] != 0);
      if (w_all_rows && c0 + 31 <= w_min_t && c0 >= w_max_start && kv_all) {
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
            const bool keep = (tj <= ti) && (tj >= my_start) && s_kvalid[b2][colhalf * 32 + 2 * e + u];
            a[u] = keep ? da[2 * e + u] * silu_grad_fast_f(sv[2 * e + u]) : 0.f;
          }
          __nv_bfloat162 h2 = __floats2bfloat162_rn(a[0], a[1]);
          pk[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
      }
    }
    put_pk32_sw128(sdS, row, colhalf, pk);
    if (tid < 64 && it + 1 < n_it) {
      const int kk = k0 + 64 + tid;
      s_kvalid[b2 ^ 1][tid] = kk < T ? key_valid[kk] : 0;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t k = sK0 + (uint32_t)(it % 3) * HT;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdQ, desc_p(sdS, ks), C::desc_mn(k, ks), idesc_q, (it > 0 || ks > 0));      // dQ += dS K
      if (it + 1 < n_it) {
        mbar_wait(bar_k0 + 8 * ((it + 1) % 3), (uint32_t)((it + 1) / 3) & 1u);
        mbar_wait(bar_v0 + 8 * (b2 ^ 1), (uint32_t)((it + 1) >> 1) & 1u);
        tc_fence_after();
        issue_s_da(it + 1);
        umma_commit(bar_s);
      } else {
        umma_commit(bar_fin);
      }
    }
  }
  mbar_wait(bar_fin, 0);
  tc_fence_after();
  {
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
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <int DH>
__global__ void __launch_bounds__(256)
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
  const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_s = bars + 24, bar_fin = bars + 32, tslot = bars + 40;
  __shared__ int s_qstart[2][64];
  __shared__ int s_qsmax[2][2];      // per 32-query chunk: latest sequence start (INT_MAX when a query is beyond T)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, colhalf = warp >> 2;
  const int k0 = blockIdx.x * 128, h = blockIdx.y;
  const int row = quarter * 32 + lane;
  const int tj = k0 + row;
  if (tid == 0) {
    tma_prefetch_desc(&map128); tma_prefetch_desc(&map64); tma_prefetch_desc(&mapdo64);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_q0 + 8 * i, 1);
    mbar_init(bar_s, 1); mbar_init(bar_fin, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  const bool kv_ok = tj < T && key_valid[tj] != 0;
  const int k_last = min(k0 + 127, T - 1);
  const int q_hi = seq_off[find_seq(seq_off, B, k_last) + 1] - 1;   // last token of the last key's sequence
  const int qt_first = k0 >> 6, qt_last = q_hi >> 6;
  const int n_it = qt_last - qt_first + 1;
  auto load_qstart = [&](int buf, int i0) {      // threads 0..63: sequence start of every query of the tile
    const int qs = (i0 + tid < T) ? seq_off[find_seq(seq_off, B, i0 + tid)] : INT_MAX;
    s_qstart[buf][tid] = qs;
    const int mx = warp_max_i(qs);
    if (lane == 0) s_qsmax[buf][warp] = mx;
  };
  if (tid < 64) load_qstart(0, qt_first * 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
  const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t tS = 0, tdA = 64, tdK = 128, tdV = 192;
  const bool w_kv_all = __all_sync(0xffffffffu, kv_ok);
  const int w_key_lo = k0 + quarter * 32, w_key_hi = k0 + quarter * 32 + 31;
  const uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);
  auto issue_s_da = [&](int it) {   // S^T(it) = K Q^T, dA^T(it) = V dO^T
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
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * TILE);
    tma_load_2d(sK, &map128, bar_kv, 3 * D + h * DH, k0);
    tma_load_2d(sV, &map128, bar_kv, 1 * D + h * DH, k0);
    load_q_do(0);
    if (n_it > 1) load_q_do(1);
    mbar_wait(bar_kv, 0);
    mbar_wait(bar_q0, 0);
    tc_fence_after();
    issue_s_da(0);
    umma_commit(bar_s);
  }
  for (int it = 0; it < n_it; ++it) {
    const int b = it & 1;
    const int i0 = (qt_first + it) * 64;
    mbar_wait(bar_s, (uint32_t)it & 1u);        // S^T(it), dA^T(it), dV(it-1), dK(it-1) retired
    tc_fence_after();
    if (tid == 0 && it >= 1 && it + 1 < n_it) load_q_do(it + 1);   // buffer of tile it-1 is free
    const int c0 = i0 + colhalf * 32;            // first query of this thread's chunk
    uint32_t pp[16], pd[16];
    __syncwarp();
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
    put_pk32_sw128(sPT, row, colhalf, pp);
    put_pk32_sw128(sdST, row, colhalf, pd);
    if (tid < 64 && it + 1 < n_it) load_qstart(b ^ 1, i0 + 64);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t q = sQ0 + (uint32_t)b * HT, g = sdO0 + (uint32_t)b * HT;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdV, desc_p(sPT, ks), C::desc_mn(g, ks), idesc_g, (it > 0 || ks > 0));   // dV += P^T dO
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem + tdK, desc_p(sdST, ks), C::desc_mn(q, ks), idesc_g, (it > 0 || ks > 0));  // dK += dS^T Q
      if (it + 1 < n_it) {
        mbar_wait(bar_q0 + 8 * (b ^ 1), (uint32_t)((it + 1) >> 1) & 1u);
        tc_fence_after();
        issue_s_da(it + 1);
        umma_commit(bar_s);
      } else {
        umma_commit(bar_fin);
      }
    }
  }
  mbar_wait(bar_fin, 0);
  tc_fence_after();
  {
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
  if (warp == 0) {
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
    size_t smem_p = 3 * 128 * DH * 2 + 16384 + 256 + 1024;
    { static bool once_p = false; if (!once_p) { B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); once_p = true; } }
    attn_tc_fwd_pipe_kernel<DH><<<grid, 256, smem_p, st>>>(m128, m64, seq_off, B, key_valid, T, D, inv_n, out);
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
    size_t smem_pq = 2 * 128 * DH * 2 + 5 * 64 * DH * 2 + 16384 + 256 + 1024;
    size_t smem_pkv = 2 * 128 * DH * 2 + 4 * 64 * DH * 2 + 32768 + 256 + 1024;
    { static bool once_p = false; if (!once_p) {
        B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dq_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pq));
        B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_dkv_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pkv));
        once_p = true; } }
    attn_tc_bwd_dq_pipe_kernel<DH><<<grid, 256, smem_pq, st>>>(m128, m64, do128, seq_off, B, key_valid, T, D, inv_n,
                                                               pre_base + 2 * D, ld, d_pre_base + 2 * D);
    attn_tc_bwd_dkv_pipe_kernel<DH><<<grid, 256, smem_pkv, st>>>(m128, m64, do64, seq_off, B, key_valid, T, D, inv_n,
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
