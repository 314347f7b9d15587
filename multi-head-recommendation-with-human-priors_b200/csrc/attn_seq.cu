// HSTU attention for SHORT sequences (max_len <= 64, bf16): one CTA per (sequence, head), the whole
// [n, n] score block lives in registers (hstu.py:137-160, backward App. D.1).
//
//   A = silu(q k^T) / n_pad * [key valid & j <= i],  out = A v
//
// With L = 50 (Pixel8M) a 128-row tcgen05 tile spans 2-3 sequences and 60-80 % of every tile is
// masked; the per-layer work (~0.2 GFLOP) is so small that the kernel is bound by its dependency
// chain (TMA -> MMA -> TMEM load -> ...), not by math or HBM.  Here every warp owns 16 query (or key)
// rows, issues warp-level mma.sync.m16n8k16 on fragments loaded with ldmatrix from one padded
// shared-memory copy of q/k/v/d_out, and keeps S, A and dS in registers: no barriers after the
// initial load, no atomics (one CTA owns the sequence, so dK/dV need no cross-CTA reduction), and
// backward is ONE launch.  Causal and beyond-length 16x16 blocks are skipped.
//
// Sequences longer than 64 tokens are only legal when they carry no valid key (the all-padding
// dummy row of static-shape mode): their outputs / gradients are written as zeros.
#include "common.cuh"

namespace {

constexpr int SEQ_MAX = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int DH>
struct Tile {
  bf16 m[SEQ_MAX][DH + 8];  // +16 bytes per row: the 8 row addresses of an ldmatrix hit 8 different bank groups
};

// rows [t0, t0+n) x DH columns starting at src -> padded tile; rows >= n are zero.  Asynchronous 16-byte
// copies (LDGSTS): every load of the CTA is in flight before the single wait in load_wait().
template <int DH>
__device__ __forceinline__ void load_rows(Tile<DH>& dst, const bf16* __restrict__ src, int64_t ld, int n) {
  constexpr int V = DH / 8;
  for (int idx = threadIdx.x; idx < SEQ_MAX * V; idx += blockDim.x) {
    int r = idx / V, c = idx - r * V;
    if (r < n) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&dst.m[r][c * 8])),
                   "l"(src + (int64_t)r * ld + c * 8)
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(&dst.m[r][c * 8]) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
__device__ __forceinline__ void load_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

// acc[nt] (16 x 8 blocks nt = 0..7 over 64 columns) = X[r0.., :] * Y[:, :]^T, both [row][DH] tiles.
// Column blocks of 16 with (cb_lo <= cb < cb_hi) only.
template <int DH>
__device__ __forceinline__ void mm_xyT(const Tile<DH>& X, int r0, const Tile<DH>& Y, int cb_lo, int cb_hi,
                                       float (&acc)[8][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  const int a_row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, a_col = ((lane >> 4) & 1) * 8;
  const int b_row = (lane & 7) + ((lane >> 4) & 1) * 8, b_col = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    uint32_t a[4];
    ldsm_x4(smem_u32(&X.m[a_row][kk * 16 + a_col]), a);
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      if (cb >= cb_lo && cb < cb_hi) {
        uint32_t b[4];
        ldsm_x4(smem_u32(&Y.m[cb * 16 + b_row][kk * 16 + b_col]), b);
        mma_bf16(acc[2 * cb], a, b[0], b[1]);
        mma_bf16(acc[2 * cb + 1], a, b[2], b[3]);
      }
    }
  }
}

// o[nt] (16 x 8 blocks over DH columns) = P (register A-fragments, 16 x 64, k-blocks of 16) * Y[k][:]
template <int DH>
__device__ __forceinline__ void mm_pY(const uint32_t (&p)[4][4], const Tile<DH>& Y, int kb_lo, int kb_hi,
                                      float (&o)[DH / 8][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
  const int b_row = (lane & 7) + ((lane >> 3) & 1) * 8, b_col = ((lane >> 4) & 1) * 8;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    if (kb >= kb_lo && kb < kb_hi) {
#pragma unroll
      for (int dp = 0; dp < DH / 16; ++dp) {
        uint32_t b[4];
        ldsm_x4_t(smem_u32(&Y.m[kb * 16 + b_row][dp * 16 + b_col]), b);
        mma_bf16(o[2 * dp], p[kb], b[0], b[1]);
        mma_bf16(o[2 * dp + 1], p[kb], b[2], b[3]);
      }
    }
  }
}

template <int DH>
__device__ __forceinline__ void zero_long_rows(float* out32, bf16* out16, int64_t ld, int n) {
  // n > SEQ_MAX: all-padding sequence
  for (int idx = threadIdx.x; idx < n * (DH / 2); idx += blockDim.x) {
    int r = idx / (DH / 2), c = (idx - r * (DH / 2)) * 2;
    if (out32) *reinterpret_cast<float2*>(out32 + (int64_t)r * ld + c) = make_float2(0.f, 0.f);
    if (out16) *reinterpret_cast<uint32_t*>(out16 + (int64_t)r * ld + c) = 0u;
  }
}

// ------------------------------------------------------------------------------------ forward
template <int DH>
__global__ void __launch_bounds__(128) attn_seq_fwd_kernel(const bf16* __restrict__ act, int ld,
                                                           const int32_t* __restrict__ seq_off,
                                                           const uint8_t* __restrict__ key_valid, int D, float inv_n,
                                                           float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) Tile<DH> sq, sk, sv;
  __shared__ uint8_t kv[SEQ_MAX];
  const int h = blockIdx.x, b = blockIdx.y;
  const int t0 = seq_off[b], n = seq_off[b + 1] - t0;
  if (n <= 0) return;
  float* o_base = out + (int64_t)t0 * D + h * DH;
  if (n > SEQ_MAX) {
    zero_long_rows<DH>(o_base, nullptr, D, n);
    return;
  }
  const bf16* base = act + (int64_t)t0 * ld + h * DH;
  load_rows<DH>(sv, base + D, ld, n);
  load_rows<DH>(sq, base + 2 * D, ld, n);
  load_rows<DH>(sk, base + 3 * D, ld, n);
  if (threadIdx.x < SEQ_MAX) kv[threadIdx.x] = threadIdx.x < n ? key_valid[t0 + threadIdx.x] : 0;
  load_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int i0 = 16 * w;
  if (i0 >= n) return;
  const int cb_hi = min(w + 1, (n + 15) >> 4);  // key blocks 0..w (causal), below the length
  float s[8][4];
  mm_xyT<DH>(sq, i0, sk, 0, cb_hi, s);
  uint32_t p[4][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int i = i0 + g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
      a[e] = (j <= i && kv[j]) ? silu_f(s[nt][e]) * inv_n : 0.f;
    }
    p[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(a[0], a[1]);
    p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(a[2], a[3]);
  }
  float o[DH / 8][4];
  mm_pY<DH>(p, sv, 0, cb_hi, o);
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
    int c = nt * 8 + 2 * t;
    if (i0 + g < n) *reinterpret_cast<float2*>(o_base + (int64_t)(i0 + g) * D + c) = make_float2(o[nt][0], o[nt][1]);
    if (i0 + g + 8 < n)
      *reinterpret_cast<float2*>(o_base + (int64_t)(i0 + g + 8) * D + c) = make_float2(o[nt][2], o[nt][3]);
  }
}

// ------------------------------------------------------------------------------------ backward
// d_pre[r, c] = grad[r, c] * silu'(pre[r, c]) for the 16 x DH accumulator block of this warp
template <int DH>
__device__ __forceinline__ void store_dpre(const float (&acc)[DH / 8][4], const bf16* __restrict__ pre,
                                           bf16* __restrict__ dpre, int64_t ld, int r0, int n) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    int r = r0 + g + half * 8;
    if (r >= n) continue;
    uint32_t z[DH / 8];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) z[nt] = __ldg(reinterpret_cast<const uint32_t*>(pre + (int64_t)r * ld + nt * 8 + 2 * t));
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      __nv_bfloat162 zz = *reinterpret_cast<__nv_bfloat162*>(&z[nt]);
      float lo = acc[nt][half * 2 + 0] * silu_grad_fast_f(__low2float(zz));
      float hi = acc[nt][half * 2 + 1] * silu_grad_fast_f(__high2float(zz));
      *reinterpret_cast<uint32_t*>(dpre + (int64_t)r * ld + nt * 8 + 2 * t) = pack_bf16(lo, hi);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(128) attn_seq_bwd_kernel(const bf16* __restrict__ act, const bf16* __restrict__ pre,
                                                           int ld, const int32_t* __restrict__ seq_off,
                                                           const uint8_t* __restrict__ key_valid, int D, float inv_n,
                                                           const bf16* __restrict__ d_out, bf16* __restrict__ d_pre) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) Tile<DH> sq, sk, sv, sd;
  __shared__ uint8_t kv[SEQ_MAX];
  const int h = blockIdx.x, b = blockIdx.y;
  const int t0 = seq_off[b], n = seq_off[b + 1] - t0;
  if (n <= 0) return;
  const int64_t row0 = (int64_t)t0 * ld + h * DH;
  if (n > SEQ_MAX) {
    for (int part = 1; part < 4; ++part) zero_long_rows<DH>(nullptr, d_pre + row0 + part * D, ld, n);
    return;
  }
  load_rows<DH>(sv, act + row0 + D, ld, n);
  load_rows<DH>(sq, act + row0 + 2 * D, ld, n);
  load_rows<DH>(sk, act + row0 + 3 * D, ld, n);
  load_rows<DH>(sd, d_out + (int64_t)t0 * D + h * DH, D, n);
  if (threadIdx.x < SEQ_MAX) kv[threadIdx.x] = threadIdx.x < n ? key_valid[t0 + threadIdx.x] : 0;
  load_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = 16 * w;
  if (r0 >= n) return;
  const int nb = (n + 15) >> 4;
  float x[8][4], y[8][4];
  uint32_t p[4][4];
  float o[DH / 8][4];
  {
    // ---- rows = queries i: dS[i, j] = (d_out v^T)[i, j] * silu'(S[i, j]) / n_pad * mask ; dq = dS k
    const int cb_hi = min(w + 1, nb);
    mm_xyT<DH>(sq, r0, sk, 0, cb_hi, x);
    mm_xyT<DH>(sd, r0, sv, 0, cb_hi, y);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float a[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int i = r0 + g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
        a[e] = (j <= i && kv[j]) ? y[nt][e] * silu_grad_fast_f(x[nt][e]) * inv_n : 0.f;
      }
      p[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(a[0], a[1]);
      p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(a[2], a[3]);
    }
    mm_pY<DH>(p, sk, 0, cb_hi, o);
    store_dpre<DH>(o, pre + row0 + 2 * D, d_pre + row0 + 2 * D, ld, r0, n);
  }
  {
    // ---- rows = keys j: S^T = k q^T, dA^T = v d_out^T over query blocks w..nb-1
    //      dv = A^T d_out ; dk = dS^T q
    mm_xyT<DH>(sk, r0, sq, w, nb, x);
    mm_xyT<DH>(sv, r0, sd, w, nb, y);
    uint32_t ps[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float a[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int j = r0 + g + (e >> 1) * 8, i = nt * 8 + 2 * t + (e & 1);
        bool on = (j <= i) && kv[j] && (i < n);
        float sg = sigmoid_fast_f(x[nt][e]);
        a[e] = on ? x[nt][e] * sg * inv_n : 0.f;
        ds[e] = on ? y[nt][e] * sg * (1.f + x[nt][e] * (1.f - sg)) * inv_n : 0.f;
      }
      p[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(a[0], a[1]);
      p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(a[2], a[3]);
      ps[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
      ps[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    mm_pY<DH>(p, sd, w, nb, o);
    store_dpre<DH>(o, pre + row0 + D, d_pre + row0 + D, ld, r0, n);
    mm_pY<DH>(ps, sq, w, nb, o);
    store_dpre<DH>(o, pre + row0 + 3 * D, d_pre + row0 + 3 * D, ld, r0, n);
  }
}

}  // namespace

extern "C" int b200rec_hstu_attn_seq_fwd(const void* act, int ld, const int32_t* seq_off, const uint8_t* key_valid,
                                         int B, int T, int n_heads, int dh, float inv_n, int max_len, float* out,
                                         void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld == 4 * n_heads * dh, "attn_seq: ld must be 4*D");
  B200_CHECK_ARG(max_len <= SEQ_MAX, "attn_seq: max_len %d > %d (use b200rec_hstu_attn_tc_fwd)", max_len, SEQ_MAX);
  B200_CHECK_ARG(((uintptr_t)act & 15) == 0 && ((uintptr_t)out & 7) == 0, "attn_seq: alignment");
  dim3 grid(n_heads, B);
  if (dh == 64)
    B200_CUDA_OK(launch_pdl(attn_seq_fwd_kernel<64>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const bf16*)act, ld, seq_off, key_valid,
                                                                     n_heads * dh, inv_n, out));
  else if (dh == 32)
    B200_CUDA_OK(launch_pdl(attn_seq_fwd_kernel<32>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const bf16*)act, ld, seq_off, key_valid,
                                                                     n_heads * dh, inv_n, out));
  else {
    b200rec_set_error("attn_seq: head dim %d not supported (32 or 64)", dh);
    return 1;
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200rec_hstu_attn_seq_bwd(const void* act, const void* pre, int ld, const int32_t* seq_off,
                                         const uint8_t* key_valid, int B, int T, int n_heads, int dh, float inv_n,
                                         int max_len, const void* d_out, void* d_pre, void* stream) {
  if (T == 0 || B == 0) return 0;
  B200_CHECK_ARG(ld == 4 * n_heads * dh, "attn_seq: ld must be 4*D");
  B200_CHECK_ARG(max_len <= SEQ_MAX, "attn_seq: max_len %d > %d (use b200rec_hstu_attn_tc_bwd)", max_len, SEQ_MAX);
  B200_CHECK_ARG(((uintptr_t)act & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && ((uintptr_t)pre & 3) == 0 &&
                     ((uintptr_t)d_pre & 3) == 0,
                 "attn_seq: alignment");
  dim3 grid(n_heads, B);
  if (dh == 64)
    B200_CUDA_OK(launch_pdl(attn_seq_bwd_kernel<64>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, 
        (const bf16*)act, (const bf16*)pre, ld, seq_off, key_valid, n_heads * dh, inv_n, (const bf16*)d_out, (bf16*)d_pre));
  else if (dh == 32)
    B200_CUDA_OK(launch_pdl(attn_seq_bwd_kernel<32>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, 
        (const bf16*)act, (const bf16*)pre, ld, seq_off, key_valid, n_heads * dh, inv_n, (const bf16*)d_out, (bf16*)d_pre));
  else {
    b200rec_set_error("attn_seq: head dim %d not supported (32 or 64)", dh);
    return 1;
  }
  B200_LAUNCH_OK();
  return 0;
}
