// Epilogues shared by the tcgen05 GEMM and the fp32 SIMT verification GEMM.
#pragma once
#include "common.cuh"

struct EpiParams {
  void* C; int64_t ldc; int c_dtype;
  void* C2; int64_t ldc2; int c2_dtype;
  int mode; float alpha;
  const float* alpha_dev;  // optional device scalar multiplied into alpha (upstream grad scale)
  const float* bias;
  const float* resid; int64_t ldr;
  int n_split; int64_t c_split_stride, c2_split_stride;
  int M, N;
  int vec_ok;  // host-checked: every pointer / leading dimension allows 4-element vector access
  // B200REC_EPI_FOLD_HEADS
  int fold_hp;
  const uint8_t* fold_head_on; const int32_t* fold_head_cat; const uint32_t* fold_item_tags;
  int64_t fold_id_offset, fold_id_stride;
  const float* fold_thr; uint32_t* fold_cnt; unsigned long long* fold_keys; int fold_cap; int fold_groups;
  const uint32_t* fold_on_bits;   // FOLD_ITEMS: per-user bit mask of the heads that are on
  const float* gt_row; const float* gt_col;   // GT_BITS: bit = acc + gt_row[m] * gt_col[n] > alpha
  // STORE / ACCUM: optional per-row factor (fp32[M])
  const float* row_scale;
  // B200REC_EPI_NCE_EXP
  const float* nce_mref; const float* nce_thr; float* nce_stats; const float* nce_logit_scale; int nce_parts;
};

__device__ __forceinline__ void epi_store_scalar(void* base, int dtype, int64_t off, float v) {
  if (dtype == B200REC_F32) ((float*)base)[off] = v;
  else ((bf16*)base)[off] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ int64_t epi_offset(const EpiParams& p, int m, int n, int64_t ld, int64_t split_stride) {
  if (p.n_split > 0 && split_stride > 0) {
    int h = n / p.n_split;
    return (int64_t)h * split_stride + (int64_t)m * ld + (n - h * p.n_split);
  }
  return (int64_t)m * ld + n;
}

// one element (SIMT path).  GT_BITS uses atomicOr: the caller zero-fills C first.
__device__ __forceinline__ void epi_apply_scalar(const EpiParams& p, int m, int n, float acc) {
  if (m >= p.M || n >= p.N) return;
  switch (p.mode) {
    case B200REC_EPI_NCE_EXP:     // the exponentials were taken in the TMEM phase: plain store
    case B200REC_EPI_STORE:
      epi_store_scalar(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride),
                       p.alpha * (p.row_scale ? p.row_scale[m] : 1.f) * acc);
      break;
    case B200REC_EPI_ACCUM: {
      float* c = (float*)p.C + epi_offset(p, m, n, p.ldc, p.c_split_stride);
      *c += p.alpha * (p.row_scale ? p.row_scale[m] : 1.f) * acc;
      break;
    }
    case B200REC_EPI_SILU_DUAL:
      epi_store_scalar(p.C2, p.c2_dtype, epi_offset(p, m, n, p.ldc2, p.c2_split_stride), acc);
      epi_store_scalar(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), silu_f(acc));
      break;
    case B200REC_EPI_BIAS_RESID: {
      float v = acc + (p.bias ? p.bias[n] : 0.f) + (p.resid ? p.resid[(int64_t)m * p.ldr + n] : 0.f);
      epi_store_scalar(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
      break;
    }
    case B200REC_EPI_RESBLOCK: {
      float z = acc + (p.bias ? p.bias[n] : 0.f);
      int nr = p.n_split > 0 ? n % p.n_split : n;
      if (p.C2) epi_store_scalar(p.C2, p.c2_dtype, epi_offset(p, m, n, p.ldc2, p.c2_split_stride), z);
      epi_store_scalar(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride),
                       p.resid[(int64_t)m * p.ldr + nr] + silu_f(z));
      break;
    }
    case B200REC_EPI_GT_BITS:
      if (acc > p.alpha) {
        atomicOr((unsigned int*)p.C + (int64_t)m * p.ldc + (n >> 5), 1u << (n & 31));
        if (p.C2) ((uint8_t*)p.C2)[m] = 1;  // "row has a set bit" flag (caller zero-fills)
      }
      break;
  }
}

template <typename T>
__device__ __forceinline__ void store_chunk(T* dst, const float* v, int n_valid);
template <>
__device__ __forceinline__ void store_chunk<float>(float* dst, const float* v, int n_valid) {
  if (n_valid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f32x4 q = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
      reinterpret_cast<f32x4*>(dst)[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_valid) dst[i] = v[i];
  }
}
template <>
__device__ __forceinline__ void store_chunk<bf16>(bf16* dst, const float* v, int n_valid) {
  if (n_valid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
      __nv_bfloat162 d = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      q.x = *reinterpret_cast<uint32_t*>(&a);
      q.y = *reinterpret_cast<uint32_t*>(&b);
      q.z = *reinterpret_cast<uint32_t*>(&c);
      q.w = *reinterpret_cast<uint32_t*>(&d);
      reinterpret_cast<uint4*>(dst)[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_valid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

__device__ __forceinline__ void store_chunk_dt(void* base, int dtype, int64_t off, const float* v, int n_valid) {
  if (dtype == B200REC_F32) store_chunk<float>((float*)base + off, v, n_valid);
  else store_chunk<bf16>((bf16*)base + off, v, n_valid);
}

// 32 consecutive columns [n0, n0+32) of row m held by one thread (tcgen05.ld 32x32b layout).
// n0 is a multiple of 32 and (when n_split > 0) n_split is a multiple of 32, so a chunk never
// straddles a split boundary.
__device__ __forceinline__ void epi_apply_chunk32(const EpiParams& p, int m, int n0, float (&acc)[32]) {
  if (m >= p.M || n0 >= p.N) return;
  int n_valid = min(32, p.N - n0);
  switch (p.mode) {
    case B200REC_EPI_STORE: {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] *= p.alpha;
      store_chunk_dt(p.C, p.c_dtype, epi_offset(p, m, n0, p.ldc, p.c_split_stride), acc, n_valid);
      break;
    }
    case B200REC_EPI_ACCUM: {
      float* c = (float*)p.C + epi_offset(p, m, n0, p.ldc, p.c_split_stride);
      if (n_valid == 32 && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f32x4 q = reinterpret_cast<f32x4*>(c)[i];
          q.x += p.alpha * acc[4 * i]; q.y += p.alpha * acc[4 * i + 1];
          q.z += p.alpha * acc[4 * i + 2]; q.w += p.alpha * acc[4 * i + 3];
          reinterpret_cast<f32x4*>(c)[i] = q;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < n_valid) c[i] += p.alpha * acc[i];
      }
      break;
    }
    case B200REC_EPI_SILU_DUAL: {
      store_chunk_dt(p.C2, p.c2_dtype, epi_offset(p, m, n0, p.ldc2, p.c2_split_stride), acc, n_valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = silu_f(acc[i]);
      store_chunk_dt(p.C, p.c_dtype, epi_offset(p, m, n0, p.ldc, p.c_split_stride), acc, n_valid);
      break;
    }
    case B200REC_EPI_BIAS_RESID: {
      const float* r = p.resid ? p.resid + (int64_t)m * p.ldr + n0 : nullptr;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < n_valid) acc[i] += (p.bias ? __ldg(p.bias + n0 + i) : 0.f) + (r ? r[i] : 0.f);
      store_chunk_dt(p.C, p.c_dtype, epi_offset(p, m, n0, p.ldc, p.c_split_stride), acc, n_valid);
      break;
    }
    case B200REC_EPI_RESBLOCK: {
      int nr0 = p.n_split > 0 ? n0 % p.n_split : n0;
      const float* r = p.resid + (int64_t)m * p.ldr + nr0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < n_valid) acc[i] += (p.bias ? __ldg(p.bias + n0 + i) : 0.f);
      if (p.C2) store_chunk_dt(p.C2, p.c2_dtype, epi_offset(p, m, n0, p.ldc2, p.c2_split_stride), acc, n_valid);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < n_valid) acc[i] = r[i] + silu_f(acc[i]);
      store_chunk_dt(p.C, p.c_dtype, epi_offset(p, m, n0, p.ldc, p.c_split_stride), acc, n_valid);
      break;
    }
    case B200REC_EPI_GT_BITS: {
      uint32_t w = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) w |= (i < n_valid && acc[i] > p.alpha) ? (1u << i) : 0u;
      ((uint32_t*)p.C)[(int64_t)m * p.ldc + (n0 >> 5)] = w;
      if (w != 0u && p.C2) ((uint8_t*)p.C2)[m] = 1;  // "row has a set bit" flag (caller zero-fills)
      break;
    }
  }
}

// 4 consecutive columns [n, n+4) of row m held by one lane (coalesced phase of the tcgen05 epilogue:
// 8 lanes cover 128 contiguous bytes of a row, so every global access is a full-line transaction).
template <typename T>
__device__ __forceinline__ void st4_dt(T* p, const float (&v)[4]) { store4<T>(p, v); }

__device__ __forceinline__ void store4_dt(void* base, int dtype, int64_t off, const float (&v)[4]) {
  if (dtype == B200REC_F32) store4<float>((float*)base + off, v);
  else store4<bf16>((bf16*)base + off, v);
}

// out-of-line tail / unaligned path keeps the hot epilogue small (instruction-cache resident)
static __device__ __noinline__ void epi_apply_tail4(const EpiParams& p, int m, int n, float v0, float v1, float v2, float v3) {
  epi_apply_scalar(p, m, n, v0);
  epi_apply_scalar(p, m, n + 1, v1);
  epi_apply_scalar(p, m, n + 2, v2);
  epi_apply_scalar(p, m, n + 3, v3);
}

template <int MODE>
__device__ __forceinline__ void epi_apply_vec4(const EpiParams& p, int m, int n, float (&v)[4]) {
  if (m >= p.M || n >= p.N) return;
  if (!p.vec_ok || n + 4 > p.N) {
    epi_apply_tail4(p, m, n, v[0], v[1], v[2], v[3]);
    return;
  }
  if (MODE == B200REC_EPI_NCE_EXP) {
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  } else if (MODE == B200REC_EPI_STORE) {
    const float a = p.row_scale ? p.alpha * __ldg(p.row_scale + m) : p.alpha;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= a;
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  } else if (MODE == B200REC_EPI_ACCUM) {
    float* c = (float*)p.C + epi_offset(p, m, n, p.ldc, p.c_split_stride);
    const float a = p.row_scale ? p.alpha * __ldg(p.row_scale + m) : p.alpha;
    float o[4];
    load4<float>(c, o);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] += a * v[i];
    store4<float>(c, o);
  } else if (MODE == B200REC_EPI_SILU_DUAL) {
    store4_dt(p.C2, p.c2_dtype, epi_offset(p, m, n, p.ldc2, p.c2_split_stride), v);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = silu_fast_f(v[i]);
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  } else if (MODE == B200REC_EPI_BIAS_RESID) {
    if (p.bias) {
      float b[4];
      load4<float>(p.bias + n, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += b[i];
    }
    if (p.resid) {
      float r[4];
      load4<float>(p.resid + (int64_t)m * p.ldr + n, r);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += r[i];
    }
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  } else if (MODE == B200REC_EPI_RESBLOCK) {
    if (p.bias) {
      float b[4];
      load4<float>(p.bias + n, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += b[i];
    }
    if (p.C2) store4_dt(p.C2, p.c2_dtype, epi_offset(p, m, n, p.ldc2, p.c2_split_stride), v);
    int nr = p.n_split > 0 ? n % p.n_split : n;
    float r[4];
    load4<float>(p.resid + (int64_t)m * p.ldr + nr, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = r[i] + silu_fast_f(v[i]);
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  }
}

// Variant with the row-dependent addend (residual row, or C itself for ACCUM) already in registers:
// the caller issues all such loads of a chunk first so they overlap instead of serialising.
template <int MODE>
__device__ __forceinline__ bool epi_needs_prefetch() {
  return MODE == B200REC_EPI_ACCUM || MODE == B200REC_EPI_BIAS_RESID || MODE == B200REC_EPI_RESBLOCK;
}

template <int MODE>
__device__ __forceinline__ bool epi_prefetch_vec4(const EpiParams& p, int m, int n, float (&r)[4]) {
  if (m >= p.M || !p.vec_ok || n + 4 > p.N) return false;
  if (MODE == B200REC_EPI_ACCUM) {
    load4<float>((const float*)p.C + epi_offset(p, m, n, p.ldc, p.c_split_stride), r);
  } else if (MODE == B200REC_EPI_BIAS_RESID) {
    if (p.resid) load4<float>(p.resid + (int64_t)m * p.ldr + n, r);
    else r[0] = r[1] = r[2] = r[3] = 0.f;
  } else if (MODE == B200REC_EPI_RESBLOCK) {
    int nr = p.n_split > 0 ? n % p.n_split : n;
    load4<float>(p.resid + (int64_t)m * p.ldr + nr, r);
  }
  return true;
}

template <int MODE>
__device__ __forceinline__ void epi_finish_vec4(const EpiParams& p, int m, int n, float (&v)[4], const float (&r)[4],
                                                const float (&b)[4]) {
  if (MODE == B200REC_EPI_ACCUM) {
    const float a = p.row_scale ? p.alpha * __ldg(p.row_scale + m) : p.alpha;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = r[i] + a * v[i];
    store4<float>((float*)p.C + epi_offset(p, m, n, p.ldc, p.c_split_stride), o);
  } else if (MODE == B200REC_EPI_BIAS_RESID) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += b[i] + r[i];
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  } else if (MODE == B200REC_EPI_RESBLOCK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += b[i];
    if (p.C2) store4_dt(p.C2, p.c2_dtype, epi_offset(p, m, n, p.ldc2, p.c2_split_stride), v);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = r[i] + silu_fast_f(v[i]);
    store4_dt(p.C, p.c_dtype, epi_offset(p, m, n, p.ldc, p.c_split_stride), v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast coalesced phase of the tcgen05 epilogue for one 32 x 32 chunk that lies fully inside N with vector access legal
// (the common case: every chunk of config B).  Everything that does not depend on the row -- output dtype, split-column
// offsets, bias, residual column, bounds -- is resolved ONCE per chunk instead of once per 4 elements, the 8 row passes
// are fully unrolled (independent dependency chains: the two epilogue warps of a scheduler cannot hide fixed ALU latency
// by themselves) and row-dependent global loads are all issued before the first use.  The generic epi_apply_vec4 path
// remains for ragged chunks.  ncu of the SiLU-dual uvqk GEMM before this path: ~800 instructions per chunk-warp, warps
// stalled on fixed-latency dependencies ("wait") most of the time, the kernel epilogue-bound at 0.54 of the peak.
template <int MODE, typename TC, typename TC2>
__device__ __forceinline__ void epi_chunk_fast_t(const EpiParams& p, uint32_t stg, int mrow0, int ncol, int sub, int cg) {
  TC* const cbase = (TC*)p.C + epi_offset(p, 0, ncol, p.ldc, p.c_split_stride);
  TC2* c2base = nullptr;
  if ((MODE == B200REC_EPI_SILU_DUAL || MODE == B200REC_EPI_RESBLOCK) && p.C2 != nullptr)
    c2base = (TC2*)p.C2 + epi_offset(p, 0, ncol, p.ldc2, p.c2_split_stride);
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if ((MODE == B200REC_EPI_BIAS_RESID || MODE == B200REC_EPI_RESBLOCK) && p.bias != nullptr) load4<float>(p.bias + ncol, bias4);
  const float* rbase = nullptr;
  int64_t ldr = p.ldr;
  if (MODE == B200REC_EPI_BIAS_RESID && p.resid != nullptr) rbase = p.resid + ncol;
  if (MODE == B200REC_EPI_RESBLOCK) rbase = p.resid + (p.n_split > 0 ? ncol % p.n_split : ncol);
  if (MODE == B200REC_EPI_ACCUM) { rbase = (const float*)cbase; ldr = p.ldc; }
  float pre[8][4];
  if (MODE == B200REC_EPI_ACCUM || MODE == B200REC_EPI_BIAS_RESID || MODE == B200REC_EPI_RESBLOCK) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = mrow0 + 4 * j + sub;
      pre[j][0] = pre[j][1] = pre[j][2] = pre[j][3] = 0.f;
      if (rbase != nullptr && m < p.M) load4<float>(rbase + (int64_t)m * ldr, pre[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int row = 4 * j + sub;
    const int m = mrow0 + row;
    const uint32_t addr = stg + (uint32_t)row * 128u + (uint32_t)((cg ^ (row & 7)) << 4);
    float x[4];
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                 : "r"(addr)
                 : "memory");
    if (m >= p.M) continue;
    TC* const c = cbase + (int64_t)m * p.ldc;
    if (MODE == B200REC_EPI_NCE_EXP) {
      store4<TC>(c, x);
    } else if (MODE == B200REC_EPI_STORE) {
      const float a = p.row_scale ? p.alpha * __ldg(p.row_scale + m) : p.alpha;
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] *= a;
      store4<TC>(c, x);
    } else if (MODE == B200REC_EPI_ACCUM) {
      const float a = p.row_scale ? p.alpha * __ldg(p.row_scale + m) : p.alpha;
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = fmaf(a, x[i], pre[j][i]);
      store4<TC>(c, x);
    } else if (MODE == B200REC_EPI_SILU_DUAL) {
      store4<TC2>(c2base + (int64_t)m * p.ldc2, x);
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = silu_fast_f(x[i]);
      store4<TC>(c, x);
    } else if (MODE == B200REC_EPI_BIAS_RESID) {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] += bias4[i] + pre[j][i];
      store4<TC>(c, x);
    } else if (MODE == B200REC_EPI_RESBLOCK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] += bias4[i];
      if (c2base != nullptr) store4<TC2>(c2base + (int64_t)m * p.ldc2, x);
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = pre[j][i] + silu_fast_f(x[i]);
      store4<TC>(c, x);
    }
  }
}

template <int MODE>
__device__ __forceinline__ constexpr bool epi_has_fast_chunk() {
  return MODE == B200REC_EPI_STORE || MODE == B200REC_EPI_ACCUM || MODE == B200REC_EPI_SILU_DUAL ||
         MODE == B200REC_EPI_BIAS_RESID || MODE == B200REC_EPI_RESBLOCK || MODE == B200REC_EPI_NCE_EXP;
}

// dtype dispatch hoisted out of the row loop (one warp-uniform branch per chunk)
template <int MODE>
__device__ __forceinline__ void epi_chunk_fast(const EpiParams& p, uint32_t stg, int mrow0, int ncol, int sub, int cg) {
  if (MODE == B200REC_EPI_ACCUM) {
    epi_chunk_fast_t<MODE, float, float>(p, stg, mrow0, ncol, sub, cg);
  } else if (MODE == B200REC_EPI_SILU_DUAL || MODE == B200REC_EPI_RESBLOCK) {
    const bool c2_bf = p.c2_dtype == B200REC_BF16;
    if (p.c_dtype == B200REC_BF16) {
      if (c2_bf) epi_chunk_fast_t<MODE, bf16, bf16>(p, stg, mrow0, ncol, sub, cg);
      else epi_chunk_fast_t<MODE, bf16, float>(p, stg, mrow0, ncol, sub, cg);
    } else {
      if (c2_bf) epi_chunk_fast_t<MODE, float, bf16>(p, stg, mrow0, ncol, sub, cg);
      else epi_chunk_fast_t<MODE, float, float>(p, stg, mrow0, ncol, sub, cg);
    }
  } else {
    if (p.c_dtype == B200REC_BF16) epi_chunk_fast_t<MODE, bf16, float>(p, stg, mrow0, ncol, sub, cg);
    else epi_chunk_fast_t<MODE, float, float>(p, stg, mrow0, ncol, sub, cg);
  }
}
