"""Times the tcgen05 GEMM on the shapes of config B (CUDA events, L2-cold-ish: operands rotate)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200rec import _lib as L

dev = torch.device("cuda:0")
T, D, NN, H = int(os.environ.get("T", "5248")), 1024, 8192, 12
bf = torch.bfloat16
cases = [
    # name, M, N, K, a_major, b_major, epilogue, c dtype
    ("uvqk_fwd silu_dual bf16", T, 4 * D, D, 0, 1, L.EPI_SILU_DUAL, bf),
    ("oproj_fwd bias_resid f32", T, D, D, 0, 0, L.EPI_BIAS_RESID, torch.float32),
    ("heads_fwd resblock f32", T, H * D, D, 0, 0, L.EPI_RESBLOCK, torch.float32),
    ("nce_logits store f32", T, NN, D, 0, 0, L.EPI_STORE, torch.float32),
    ("nce_bits gt_bits", 7424, NN, D, 0, 0, L.EPI_GT_BITS, torch.int32),
    ("nce_dq (G@N) f32", T, D, NN, 0, 1, L.EPI_STORE, torch.float32),
    ("nce_dn (G^T@Q) f32", NN, D, T, 1, 1, L.EPI_STORE, torch.float32),
    ("d_oin (dx@Wo) bf16", T, D, D, 0, 1, L.EPI_STORE, bf),
    ("dWo (dx^T@oin) f32", D, D, T, 1, 1, L.EPI_STORE, torch.float32),
    ("dWuvqk (n^T@dpre) f32", D, 4 * D, T, 1, 1, L.EPI_STORE, torch.float32),
    ("dn (dpre@W^T) bf16", T, D, 4 * D, 0, 0, L.EPI_STORE, bf),
    ("d_y (dz@Wcat) accum f32", T, D, H * D, 0, 1, L.EPI_ACCUM, torch.float32),
    ("dWcat (dz^T@yb) f32", H * D, D, T, 1, 1, L.EPI_STORE, torch.float32),
    ("eval scores f32", 3072, 450000, D, 0, 0, L.EPI_STORE, torch.float32),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
sweep = os.environ.get("SWEEP")
res = []
configs = [(0, 0)] if not sweep else [(1, 256), (2, 256), (1, 128), (2, 128)]
for name, M, N, K, am, bm, epi, cdt in [c + () for c in cases for _ in configs]:
    cfg_i = len([r for r in res if r[0] == name])
    ctas_f, bn_f = configs[cfg_i % len(configs)]
    L.lib().b200rec_gemm_force_ctas(ctas_f if ctas_f == 1 else 0)
    L.lib().b200rec_gemm_force_bn(bn_f)
    if only and only not in name:
        res.append((name,))
        continue
    A = torch.randn((M, K) if am == 0 else (K, M), device=dev).to(bf)
    B = torch.randn((N, K) if bm == 0 else (K, N), device=dev).to(bf)
    ldc = (N + 31) // 32 if epi == L.EPI_GT_BITS else N
    C = torch.zeros((M, ldc), dtype=cdt, device=dev)
    C2 = torch.empty((M, N), dtype=bf, device=dev) if epi in (L.EPI_SILU_DUAL, L.EPI_RESBLOCK) else None
    bias = torch.randn(N, device=dev)
    resid = torch.randn(M, D if epi == L.EPI_RESBLOCK else N, device=dev) if epi in (L.EPI_BIAS_RESID, L.EPI_RESBLOCK) else None
    kw = dict(lda=A.shape[1], ldb=B.shape[1], ldc=ldc, a_major=am, b_major=bm, epilogue=epi, alpha=0.99 if epi == L.EPI_GT_BITS else 1.0)
    if C2 is not None:
        kw.update(C2=C2, ldc2=N)
    if name.startswith("dWo") or name.startswith("dWuvqk"):
        kw.update(splitk_ws=torch.empty(8 * M * N, device=dev))
    if resid is not None:
        kw.update(bias=bias, resid=resid, ldr=resid.shape[1], n_split=D if epi == L.EPI_RESBLOCK else 0)
    for _ in range(2):
        L.gemm(A, B, C, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        L.gemm(A, B, C, M, N, K, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * N * K / ms / 1e9
    res.append((name, M, N, K, ms, tf))
    print(f"{name:28s} M={M:6d} N={N:6d} K={K:6d}  {ms*1e3:9.1f} us  {tf:7.1f} TF/s  ctas={ctas_f} bn={bn_f}", flush=True)
    del A, B, C, C2
