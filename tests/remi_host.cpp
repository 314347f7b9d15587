// TEST-ONLY host build of csrc/remi_core.cuh (the routing-regulariser scan shared with the CUDA kernel in comirec.cu):
// tests/test_remi_core_cpu.py compiles this file with g++ and checks the closed forms against the dense reference
// formulation (remi.py:156-196) and its autograd in the CPU tier.  Same thread mapping as comi_rr_kernel: one (sequence,
// interest) per call of remi_rr_scan.
#include "../multi-head-recommendation-with-human-priors_b200/csrc/remi_core.cuh"

extern "C" void remi_rr_host(const float* a, const int32_t* seq_off, int B_real, int K, int D, float* var2, float* scratch,
                             float* da) {
  const int nv = seq_off[B_real];
  const float inv_n = 1.f / fmaxf((float)nv, 1.f);
  for (int idx = 0; idx < B_real * K; ++idx) {
    const int b = idx / K, k = idx % K;
    remi_rr_scan(a, K, k, seq_off[b], seq_off[b + 1], 1.f / (float)D, inv_n, var2, scratch, da);
  }
}
