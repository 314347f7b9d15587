"""One-wave tcgen05 GEMMs (148 tiles of 128x256) at growing K: intercept = launch + prologue + epilogue,
slope = mainloop cost per 64-wide k-block.  GPU-side durations from the profiler (not host-bound)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from b200rec import _lib as L

dev = torch.device("cuda:0")
bf = torch.bfloat16
for ctas in (1, 2):
    L.lib().b200rec_gemm_force_ctas(ctas if ctas == 1 else 0)
    for cdt in (torch.float32, bf):
        for K in (64, 256, 1024, 2048, 4096):
            M, N = 128 * 148, 256
            A = torch.randn(M, K, device=dev).to(bf)
            B = torch.randn(N, K, device=dev).to(bf)
            C = torch.empty(M, N, dtype=cdt, device=dev)
            for _ in range(3):
                L.gemm(A, B, C, M, N, K, lda=K, ldb=K, ldc=N)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(5):
                    L.gemm(A, B, C, M, N, K, lda=K, ldb=K, ldc=N)
                    torch.cuda.synchronize()
            ev = [e for e in prof.key_averages() if "gemm_tc" in e.key]
            us = sum(e.device_time_total for e in ev) / sum(e.count for e in ev)
            print(f"ctas={ctas} c={str(cdt)[6:]:9s} K={K:5d}: {us:7.2f} us/launch", flush=True)
