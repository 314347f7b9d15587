"""Times the tcgen05 attention kernels (forward, dQ, dK/dV) on a config-C shaped jagged batch (B = 64 sequences,
60 % of length L = 400, the rest U[1, L], 16 heads x 64) with CUDA events; `ncu -k regex:attn_tc` on this script gives
the per-kernel counters.  Env: L, B, NH, DH, REPS; B200REC_ATTN_PIPE=0 selects the un-pipelined kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200rec import _lib as L

dev = torch.device("cuda:0")
Lmax, B, nh, dh = int(os.environ.get("L", 400)), int(os.environ.get("B", 64)), int(os.environ.get("NH", 16)), int(os.environ.get("DH", 64))
reps = int(os.environ.get("REPS", 10))
g = torch.Generator().manual_seed(3)
lens = [Lmax if torch.rand(1, generator=g).item() < 0.6 else int(torch.randint(1, Lmax + 1, (1,), generator=g)) for _ in range(B)]
T, D = sum(lens), nh * dh
seq_off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
key_valid = torch.ones(T, dtype=torch.uint8, device=dev)
pre = (torch.randn(T, 4 * D, generator=g) * 0.7).to(dev).to(torch.bfloat16)
act = torch.nn.functional.silu(pre.float()).to(torch.bfloat16)
out = torch.empty(T, D, device=dev)
gr = torch.randn(T, D, generator=g).to(dev).to(torch.bfloat16)
d_pre = torch.zeros(T, 4 * D, dtype=torch.bfloat16, device=dev)
flops_causal = sum(2.0 * n * (n + 1) / 2 * dh * 2 for n in lens) * nh      # QK^T + PV on the causal half


def fwd():
    L.call("b200rec_hstu_attn_tc_fwd", act.data_ptr(), 4 * D, seq_off.data_ptr(), key_valid.data_ptr(), B, T, nh, dh,
           1.0 / Lmax, out.data_ptr(), L.stream())


def bwd():
    L.call("b200rec_hstu_attn_tc_bwd", act.data_ptr(), pre.data_ptr(), 4 * D, seq_off.data_ptr(), key_valid.data_ptr(),
           B, T, nh, dh, 1.0 / Lmax, gr.data_ptr(), d_pre.data_ptr(), L.stream())


for name, fn, mult in (("fwd", fwd, 1.0), ("bwd (dq + dkv)", bwd, 2.5)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:16s} T={T} tiles={-(-T // 128) * nh}  {us:8.1f} us   causal-useful {flops_causal * mult / us / 1e6:7.1f} TF/s", flush=True)
