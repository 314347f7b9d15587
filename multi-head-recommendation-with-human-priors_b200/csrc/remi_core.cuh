// REMI's routing regulariser on the jagged token layout (reference REC/model/IDNet/remi.py:156-196,
// 356-372).  Shared by the CUDA kernel (comirec.cu) and a host build (tests/remi_host.cpp compiles this header with g++
// to check the closed forms against the reference's dense formulation in the CPU test tier).
//
// The reference materialises, for every position l, the routing weights A[l, k, l'] = softmax over the valid window
// positions l' <= l of the attention-net logit a[l', k] and takes, per (l, k), the variance of those weights over the n
// valid window positions divided by D: var = (sum_l' A^2 - 1/n) / D  (the weights sum to 1, their mean is 1/n).  The
// logits do not depend on l, so with running sums relative to a running maximum M,
//     S1 = sum_{l' <= l} e^(a - M),  S2 = sum_{l' <= l} e^(2 (a - M)),   sum A^2 = S2 / S1^2
// one pass per (sequence, interest) gives every window: O(L) instead of O(L^2).  The loss is
//     rr = (1 / #valid positions) sum_{(b, l) valid} sum_k var[b, l, k]^2
// and its gradient w.r.t. a[j, k] a suffix sum over the later positions of the same sequence, accumulated relative to
// the running maxima (every factor e^(M_j - M_t) <= 1).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define REMI_HD __host__ __device__ __forceinline__
#else
#define REMI_HD inline
#endif

// One (sequence, interest): tokens [t0, t1) of the sequence, column k of a [T, K].
//   var2[t, k]  = var^2 (the caller sums and divides by the number of valid positions)
//   da[t, k]    = d rr / d a[t, k] (nullable: forward only), with inv_n = 1 / #valid positions of the batch
//   scratch     = [T, K, 3] floats (M, S1, S2 per token; only touched when da != NULL)
REMI_HD void remi_rr_scan(const float* a, int K, int k, int t0, int t1, float inv_D, float inv_n, float* var2,
                          float* scratch, float* da) {
  float m = -INFINITY, s1 = 0.f, s2 = 0.f;
  for (int t = t0; t < t1; ++t) {
    const int64_t i = (int64_t)t * K + k;
    const float av = a[i];
    const float mn = fmaxf(m, av);
    const float r = expf(m - mn);                      // 0 on the first token (m = -inf)
    const float w = expf(av - mn);
    s1 = s1 * r + w;
    s2 = s2 * r * r + w * w;
    m = mn;
    const float var = (s2 / (s1 * s1) - 1.f / (float)(t - t0 + 1)) * inv_D;
    var2[i] = var * var;
    if (da != nullptr) {
      scratch[i * 3 + 0] = m;
      scratch[i * 3 + 1] = s1;
      scratch[i * 3 + 2] = s2;
    }
  }
  if (da == nullptr) return;
  // d rr / d r_t = 2 var_t inv_n inv_D with r_t = S2 / S1^2;  d r_t / d a_j = 2 e^2 / S1^2 - 2 S2 e / S1^3,
  // e = e^(a_j - M_t) = e^(a_j - M_j) e^(M_j - M_t)
  float A = 0.f, Bc = 0.f, m_next = 0.f;
  for (int t = t1 - 1; t >= t0; --t) {
    const int64_t i = (int64_t)t * K + k;
    const float mt = scratch[i * 3 + 0], c_s1 = scratch[i * 3 + 1], c_s2 = scratch[i * 3 + 2];
    const float var = (c_s2 / (c_s1 * c_s1) - 1.f / (float)(t - t0 + 1)) * inv_D;
    const float g = 2.f * var * inv_n * inv_D;
    const float c1 = 2.f * g / (c_s1 * c_s1);
    const float c2 = -2.f * g * c_s2 / (c_s1 * c_s1 * c_s1);
    if (t == t1 - 1) {
      A = c1;
      Bc = c2;
    } else {
      const float d = expf(mt - m_next);               // <= 1: the running maximum never decreases
      A = c1 + A * d * d;
      Bc = c2 + Bc * d;
    }
    m_next = mt;
    const float e = expf(a[i] - mt);
    da[i] = e * e * A + e * Bc;
  }
}
