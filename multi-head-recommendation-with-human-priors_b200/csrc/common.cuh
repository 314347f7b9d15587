// Shared helpers for libb200rec kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200rec.h"

typedef __nv_bfloat16 bf16;

void b200rec_set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)         \
  do {                                    \
    if (!(cond)) {                        \
      b200rec_set_error(__VA_ARGS__);     \
      return 1;                           \
    }                                     \
  } while (0)

#define B200_CUDA_OK(expr)                                                             \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      b200rec_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define B200_LAUNCH_OK() B200_CUDA_OK(cudaGetLastError())

static inline int ceil_div_i(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// pdl_trigger(): the NEXT kernel in the stream, if it was launched with programmatic stream serialization (our GEMMs),
// may be scheduled now; it runs its prologue (barriers, TMEM allocation, tensor-map prefetch) under this kernel's tail
// and blocks in pdl_wait() until this grid has completed and flushed.  A no-op for ordinary successors.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Host side: launch `kernel` with programmatic stream serialization.  The kernel MUST execute pdl_wait() before its first
// global-memory access; its launch latency and block scheduling then overlap the tail of the previous kernel (measured on
// the 84 GEMM launches of a config-B step: -2.4 us per launch inside the replayed graph).  B200REC_PDL=0 disables.
int b200rec_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b200rec_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dtype helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// sigmoid via ex2 + approximate reciprocal (2 MUFU ops, no IEEE-division slow path): rel. error ~1e-6
__device__ __forceinline__ float sigmoid_f(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float silu_f(float z) { return z * sigmoid_f(z); }
// one-MUFU SiLU for the bf16 tensor-core epilogues: z * sigmoid(z) = z/2 * (1 + tanh(z/2)); tanh.approx has
// ~2^-11 relative error, far below bf16 resolution (the fp32 verification path keeps silu_f)
__device__ __forceinline__ float silu_fast_f(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return 0.5f * z * (1.f + t);
}
// one-MUFU sigmoid / silu' for the bf16 attention kernels, where the pointwise SiLU of every score is the bound (two MUFU
// per score cap dh = 64 attention at 1/4 of the tensor peak): sigmoid(z) = (1 + tanh(z/2)) / 2
__device__ __forceinline__ float sigmoid_fast_f(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float silu_grad_fast_f(float z) {
  const float s = sigmoid_fast_f(z);
  return s * fmaf(z, 1.f - s, 1.f);
}
__device__ __forceinline__ float silu_grad_f(float z) {
  float s = sigmoid_f(z);
  return s * (1.f + z * (1.f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result broadcast to every thread.  `red` is >= 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// 16-byte vector access helpers for rows of floats / bf16 (callers guarantee alignment).
struct alignas(16) f32x4 { float x, y, z, w; };
struct alignas(8) bf16x4 { bf16 x, y, z, w; };

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  typedef f32x4 type;
};
template <> struct Vec4<bf16> {
  typedef bf16x4 type;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]) {
  typename Vec4<T>::type v = *reinterpret_cast<const typename Vec4<T>::type*>(p);
  o[0] = to_f32(v.x); o[1] = to_f32(v.y); o[2] = to_f32(v.z); o[3] = to_f32(v.w);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&o)[4]) {
  typename Vec4<T>::type v;
  v.x = from_f32<T>(o[0]); v.y = from_f32<T>(o[1]); v.z = from_f32<T>(o[2]); v.w = from_f32<T>(o[3]);
  *reinterpret_cast<typename Vec4<T>::type*>(p) = v;
}

// ---- counter-based RNG for dropout: Philox4x32-10 (Salmon et al. 2011), stateless -------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// keep-mask of 4 consecutive elements: element group `idx4` of stream (seed, step, layer)
struct DropCfg {
  float p;                 // drop probability (0 disables)
  uint32_t seed, layer;
  const int64_t* step;     // device counter advanced once per forward (graph-replay safe); may be null
};
__device__ __forceinline__ void dropout4(const DropCfg& d, uint64_t idx4, float (&v)[4]) {
  if (d.p <= 0.f) return;
  const uint64_t st = d.step ? (uint64_t)*d.step : 0ull;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)st, (uint32_t)(st >> 32)),
                                make_uint2(d.seed, d.layer));
  const uint32_t thr = (uint32_t)fminf(d.p * 4294967296.f, 4294967295.f);
  const float sc = 1.f / (1.f - d.p);
  v[0] = r.x >= thr ? v[0] * sc : 0.f;
  v[1] = r.y >= thr ? v[1] * sc : 0.f;
  v[2] = r.z >= thr ? v[2] * sc : 0.f;
  v[3] = r.w >= thr ? v[3] * sc : 0.f;
}

#define DISPATCH_ACT(dtype, T, ...)                       \
  do {                                                    \
    if ((dtype) == B200REC_F32) {                         \
      typedef float T;                                    \
      __VA_ARGS__                                         \
    } else if ((dtype) == B200REC_BF16) {                 \
      typedef bf16 T;                                     \
      __VA_ARGS__                                         \
    } else {                                              \
      b200rec_set_error("bad dtype %d", (int)(dtype));    \
      return 1;                                           \
    }                                                     \
  } while (0)
