"""Times the false-negative filter GEMM variants (GT_BITS epilogue) in isolation: full K, K = 64 prefix without and with
the rank-1 bound.  Usage (GPU box): python scripts/gt_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200rec import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
M, N, D = 7424, 8192, 1024
g = torch.Generator().manual_seed(0)
a = torch.randn(M, D, generator=g)
a = (a / a.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(dev)
b = torch.randn(N, D, generator=g)
b = (b / b.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(dev)
nw = N // 32
bits = torch.empty((M, nw), dtype=torch.int32, device=dev)
ra = torch.zeros(M, dtype=torch.uint8, device=dev)
ta, tb = torch.ones(M, device=dev) * 0.9, torch.ones(N, device=dev) * 0.9


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, fn in (
    ("full K=1024", lambda: L.gemm(a, b, bits, M, N, D, lda=D, ldb=D, ldc=nw, epilogue=L.EPI_GT_BITS, alpha=0.99, C2=ra)),
    ("K=64 plain", lambda: L.gemm(a, b, bits, M, N, 64, lda=D, ldb=D, ldc=nw, epilogue=L.EPI_GT_BITS, alpha=0.99)),
    ("K=64 bound", lambda: L.gemm(a, b, bits, M, N, 64, lda=D, ldb=D, ldc=nw, epilogue=L.EPI_GT_BITS, alpha=0.99, gt=(ta, tb))),
    ("K=128 bound", lambda: L.gemm(a, b, bits, M, N, 128, lda=D, ldb=D, ldc=nw, epilogue=L.EPI_GT_BITS, alpha=0.99, gt=(ta, tb))),
    ("K=64 store f32 (no GT)", lambda: L.gemm(a, b, torch.empty((M, 64), device=dev), M, 64, 64, lda=D, ldb=D, ldc=64)),
):
    print(f"{name:28s} {timeit(fn):8.1f} us")
