// fp32 SIMT GEMM: the verification-mode contraction (tight-tolerance parity against the CPU
// oracle) with the same epilogues and operand-major options as the tcgen05 kernel.
#include "gemm_epilogue.cuh"

#define ST_BM 64
#define ST_BN 64
#define ST_BK 16

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int64_t lda, int a_major,
                                                        const float* __restrict__ B, int64_t ldb, int b_major,
                                                        int M, int N, int K, EpiParams ep) {
  if (ep.alpha_dev && (ep.mode == B200REC_EPI_STORE || ep.mode == B200REC_EPI_ACCUM)) ep.alpha *= *ep.alpha_dev;
  __shared__ float sA[ST_BK][ST_BM + 4];
  __shared__ float sB[ST_BK][ST_BN + 4];
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int m0 = blockIdx.y * ST_BM, n0 = blockIdx.x * ST_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += ST_BK) {
    // 64x16 elements per operand, 256 threads -> 4 each
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int mm, kk;
      if (a_major == 0) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < K) v = a_major == 0 ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
      sA[kk][mm] = v;
      int nn;
      if (b_major == 0) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      int gn = n0 + nn;
      gk = k0 + kk;
      v = 0.f;
      if (gn < N && gk < K) v = b_major == 0 ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < ST_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) epi_apply_scalar(ep, m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

int gemm_simt_launch(const b200rec_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  if (a->epilogue == B200REC_EPI_GT_BITS)
    B200_CUDA_OK(cudaMemsetAsync(a->C, 0, (size_t)a->M * a->ldc * 4, st));
  dim3 grid(ceil_div_i(a->N, ST_BN), ceil_div_i(a->M, ST_BM));
  gemm_simt_kernel<<<grid, 256, 0, st>>>((const float*)a->A, a->lda, a->a_major, (const float*)a->B, a->ldb,
                                         a->b_major, a->M, a->N, a->K, ep);
  B200_LAUNCH_OK();
  return 0;
}
