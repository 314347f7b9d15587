// ComiRec-SA multi-interest readout on the HSTU body (reference REC/model/IDNet/comirec.py:232-300 and
// its autograd backward).  The reference materialises, for every position l, the window of the L positions ending at
// l ([B, L, L, D]) and runs a masked softmax pooling over each window; a window holds exactly the valid positions
// <= l, so interest k at token t is a causal prefix softmax average over the token's own sequence:
//     u[t, k, :] = sum_{t' <= t} exp(a[t', k]) y[t', :] / sum_{t' <= t} exp(a[t', k])
// computed here in ONE pass per (sequence, interest) with a running maximum (online softmax), O(L D) instead of
// O(L^2 D), on the jagged token layout (valid positions only).  All fp32: ComiRec is a small-width baseline
// (d = 64 in the paper), the kernels are HBM / latency bound.
#include "common.cuh"

// h = tanh(z) in place; dz = dh * (1 - h^2) in place on dh
__global__ void __launch_bounds__(256) comi_tanh_kernel(float* __restrict__ z, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) z[i] = tanhf(z[i]);
}
__global__ void __launch_bounds__(256) comi_tanh_bwd_kernel(const float* __restrict__ h, float* __restrict__ dh, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dh[i] *= 1.f - h[i] * h[i];
}

extern "C" int b200rec_comi_tanh(float* z, int64_t n, void* stream) {
  if (n == 0) return 0;
  comi_tanh_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(z, n);
  B200_LAUNCH_OK();
  return 0;
}
extern "C" int b200rec_comi_tanh_bwd(const float* h, float* dh, int64_t n, void* stream) {
  if (n == 0) return 0;
  comi_tanh_bwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(h, dh, n);
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- pooling forward
// grid (B, K); thread d-strided over D.  Saves the running maximum M[t, k] and denominator S[t, k] (relative to M).
__global__ void __launch_bounds__(256) comi_pool_fwd_kernel(const float* __restrict__ a, const float* __restrict__ y,
                                                            const int32_t* __restrict__ seq_off, int K, int D,
                                                            float* __restrict__ u, float* __restrict__ Mo,
                                                            float* __restrict__ So) {
  const int b = blockIdx.x, k = blockIdx.y;
  const int t0 = seq_off[b], t1 = seq_off[b + 1];
  constexpr int MAXE = 8;                                   // D <= 2048 with 256 threads
  float acc[MAXE];
#pragma unroll
  for (int e = 0; e < MAXE; ++e) acc[e] = 0.f;
  float m = -INFINITY, s = 0.f;
  for (int t = t0; t < t1; ++t) {
    const float av = a[(int64_t)t * K + k];
    const float mn = fmaxf(m, av);
    const float r = __expf(m - mn);                         // 0 on the first token (m = -inf)
    const float w = __expf(av - mn);
    s = s * r + w;
    m = mn;
    const float inv = 1.f / s;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      const int d = threadIdx.x + e * 256;
      if (d < D) {
        acc[e] = acc[e] * r + w * y[(int64_t)t * D + d];
        u[((int64_t)t * K + k) * D + d] = acc[e] * inv;
      }
    }
    if (threadIdx.x == 0) {
      Mo[(int64_t)t * K + k] = m;
      So[(int64_t)t * K + k] = s;
    }
  }
}

extern "C" int b200rec_comi_pool_fwd(const float* a, const float* y, const int32_t* seq_off, int B, int K, int D, float* u,
                                     float* M, float* S, void* stream) {
  B200_CHECK_ARG(D >= 1 && D <= 2048 && K >= 1, "comi_pool_fwd: bad D=%d K=%d", D, K);
  if (B == 0) return 0;
  comi_pool_fwd_kernel<<<dim3(B, K), 256, 0, (cudaStream_t)stream>>>(a, y, seq_off, K, D, u, M, S);
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- hard readout
// one warp per (token t, offset p): sim_k = <u[t, k], target(b, pos + 1 + p)>, first arg-max over k (comirec.py:283-291:
// the softmax in front of the arg-max is monotone), hd[t, p, :] = u[t, best, :].
__global__ void __launch_bounds__(256) comi_select_fwd_kernel(const float* __restrict__ u, const float* __restrict__ traw,
                                                              const int32_t* __restrict__ tok_b,
                                                              const int32_t* __restrict__ tok_pos, int T, int LP, int P,
                                                              int K, int D, float* __restrict__ hd,
                                                              int32_t* __restrict__ sel) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (int64_t)T * P) return;
  const int t = (int)(wid / P), p = (int)(wid - (int64_t)t * P);
  const int64_t r = (int64_t)tok_b[t] * LP + tok_pos[t] + 1 + p;
  const float* tg = traw + r * D;
  float best = -INFINITY;
  int bk = 0;
  for (int k = 0; k < K; ++k) {
    const float* uk = u + ((int64_t)t * K + k) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(uk[d], tg[d], acc);
    acc = warp_sum(acc);
    if (acc > best) { best = acc; bk = k; }                 // strict: the first maximum wins (torch.argmax)
  }
  const float* ub = u + ((int64_t)t * K + bk) * D;
  float* o = hd + ((int64_t)t * P + p) * D;
  for (int d = lane; d < D; d += 32) o[d] = ub[d];
  if (lane == 0) sel[(int64_t)t * P + p] = bk;
}

extern "C" int b200rec_comi_select_fwd(const float* u, const float* traw, const int32_t* tok_b, const int32_t* tok_pos,
                                       int T, int LP, int P, int K, int D, float* hd, int32_t* sel, void* stream) {
  if (T == 0) return 0;
  const int64_t warps = (int64_t)T * P;
  comi_select_fwd_kernel<<<(unsigned)ceil_div_i(warps, 8), 256, 0, (cudaStream_t)stream>>>(u, traw, tok_b, tok_pos, T, LP,
                                                                                         P, K, D, hd, sel);
  B200_LAUNCH_OK();
  return 0;
}

// du[t, k, :] = sum_p [sel[t, p] == k] d_hd[t, p, :]   (fixed p order: deterministic)
__global__ void __launch_bounds__(256) comi_select_bwd_kernel(const float* __restrict__ d_hd, const int32_t* __restrict__ sel,
                                                              int T, int P, int K, int D, float* __restrict__ du) {
  const int64_t n = (int64_t)T * K * D;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int d = (int)(i % D);
    const int64_t tk = i / D;
    const int k = (int)(tk % K);
    const int64_t t = tk / K;
    float acc = 0.f;
    for (int p = 0; p < P; ++p)
      if (sel[t * P + p] == k) acc += d_hd[(t * P + p) * D + d];
    du[i] = acc;
  }
}

extern "C" int b200rec_comi_select_bwd(const float* d_hd, const int32_t* sel, int T, int P, int K, int D, float* du,
                                       void* stream) {
  if (T == 0) return 0;
  const int64_t n = (int64_t)T * K * D;
  comi_select_bwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      d_hd, sel, T, P, K, D, du);
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- pooling backward
// With Z[t] = S[t] e^{M[t]}:  dy[t'] += e^{a[t']} G[t'],  da[t'] = e^{a[t']} (<G[t'], y[t']> - H[t']),
//   G[t'] = sum_{t >= t'} du[t] / Z[t]   (vector),   H[t'] = sum_{t >= t'} <du[t], u[t]> / Z[t]   (scalar),
// accumulated from the last token backwards relative to M[t'] (M is non-decreasing, every factor is <= 1).
// One CTA per sequence, the K interests one after the other: dy receives its K contributions in a fixed order.
__global__ void __launch_bounds__(256) comi_pool_bwd_kernel(const float* __restrict__ du, const float* __restrict__ u,
                                                            const float* __restrict__ y, const float* __restrict__ a,
                                                            const float* __restrict__ Mi, const float* __restrict__ Si,
                                                            const int32_t* __restrict__ seq_off, int K, int D,
                                                            float* __restrict__ dy, float* __restrict__ da) {
  __shared__ float red[40];
  const int b = blockIdx.x;
  const int t0 = seq_off[b], t1 = seq_off[b + 1];
  constexpr int MAXE = 8;
  for (int k = 0; k < K; ++k) {
    float G[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) G[e] = 0.f;
    float H = 0.f, m_next = 0.f;
    for (int t = t1 - 1; t >= t0; --t) {
      const int64_t tk = (int64_t)t * K + k;
      const float m = Mi[tk], inv_s = 1.f / Si[tk];
      const float r = (t == t1 - 1) ? 0.f : __expf(m - m_next);
      float dot_u = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int d = threadIdx.x + e * 256;
        if (d < D) {
          const float g = du[tk * D + d];
          G[e] = G[e] * r + g * inv_s;
          dot_u = fmaf(g, u[tk * D + d], dot_u);
        }
      }
      dot_u = block_sum(dot_u, red);
      H = H * r + dot_u * inv_s;
      const float w = __expf(a[tk] - m);
      float dot_y = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int d = threadIdx.x + e * 256;
        if (d < D) {
          const float yv = y[(int64_t)t * D + d];
          dot_y = fmaf(G[e], yv, dot_y);
          dy[(int64_t)t * D + d] += w * G[e];
        }
      }
      dot_y = block_sum(dot_y, red);
      if (threadIdx.x == 0) da[tk] = w * (dot_y - H);
      m_next = m;
    }
  }
}

extern "C" int b200rec_comi_pool_bwd(const float* du, const float* u, const float* y, const float* a, const float* M,
                                     const float* S, const int32_t* seq_off, int B, int K, int D, float* dy, float* da,
                                     void* stream) {
  B200_CHECK_ARG(D >= 1 && D <= 2048 && K >= 1, "comi_pool_bwd: bad D=%d K=%d", D, K);
  if (B == 0) return 0;
  comi_pool_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(du, u, y, a, M, S, seq_off, K, D, dy, da);
  B200_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- REMI routing regulariser
// remi.py:156-196, 356-372 (and its autograd): one thread per (sequence, interest) runs the scans of remi_core.cuh
// over the sequence's tokens (the work is O(T K) scalars: latency-bound by design, a few microseconds).  Sequences
// b >= B_real (the dummy sequence of static-token mode) are not visited: their var2 / da entries stay as the caller
// initialised them (zero).  The mean is over the seq_off[B_real] valid positions of the batch, read on the device.
#include "remi_core.cuh"

__global__ void __launch_bounds__(128) comi_rr_kernel(const float* __restrict__ a, const int32_t* __restrict__ seq_off,
                                                      int B_real, int K, int D, float* __restrict__ var2,
                                                      float* __restrict__ scratch, float* __restrict__ da) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B_real * K) return;
  const int b = idx / K, k = idx % K;
  const float inv_n = 1.f / fmaxf((float)seq_off[B_real], 1.f);
  remi_rr_scan(a, K, k, seq_off[b], seq_off[b + 1], 1.f / (float)D, inv_n, var2, scratch, da);
}

extern "C" int b200rec_comi_rr(const float* a, const int32_t* seq_off, int B_real, int K, int D, float* var2,
                               float* scratch, float* da, void* stream) {
  B200_CHECK_ARG(K >= 1 && D >= 1 && B_real >= 0, "comi_rr: bad B=%d K=%d D=%d", B_real, K, D);
  B200_CHECK_ARG(da == nullptr || scratch != nullptr, "comi_rr: the backward needs the [T, K, 3] scratch");
  if (B_real == 0) return 0;
  comi_rr_kernel<<<(B_real * K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, seq_off, B_real, K, D, var2, scratch, da);
  B200_LAUNCH_OK();
  return 0;
}
