"""Torch-only reproduction (no reference / repo code): under gloo the backward of torch.distributed.nn.functional.all_gather
does not return the sum over ranks of the slot gradients when those are element-wise non-uniform (prints a non-zero
`diff` against the analytic gradient).  This is why tests/golden/multi2_*.pt pin the item-table gradient with the NCCL
branch semantics (`grads_rs`) instead of the gloo run.  Usage: python scripts/torch_gloo_allgather_backward_repro.py"""
import os, sys, socket, torch
import torch.multiprocessing as mp
W=2
def worker(rank, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from torch.distributed import nn as dnn
    torch.manual_seed(0)
    cs = [torch.randn(W,6,3,32) for _ in range(W)]       # cs[rank][slot]
    x = torch.nn.Parameter(torch.randn(6,3,32))
    xn = x / x.norm(dim=-1, keepdim=True)
    ys = torch.stack(dnn.functional.all_gather(xn), 0).reshape(-1, 32)
    yt = ys.T.contiguous()
    q = cs[rank].reshape(-1, 32)[:20]
    logits = q @ yt
    fix = q @ yt
    logits.masked_fill_(fix > 3.0, torch.finfo(logits.dtype).min)
    loss = torch.nn.functional.cross_entropy(logits*20, torch.zeros(20, dtype=torch.long))
    loss.backward()
    g = x.grad.clone()
    # expected: emulate by computing both ranks' losses locally wrt a full differentiable copy
    x2 = x.detach().clone().requires_grad_(True)
    tot = 0
    xs = [torch.empty_like(x) for _ in range(W)]
    dist.all_gather(xs, x.detach())
    leaves = [t.clone().requires_grad_(True) for t in xs]
    for r in range(W):
        ys2 = torch.stack([l / l.norm(dim=-1, keepdim=True) for l in leaves], 0).reshape(-1, 32)
        q2 = cs[r].reshape(-1, 32)[:20]
        lg = q2 @ ys2.T
        lg = lg.masked_fill((q2 @ ys2.T) > 3.0, torch.finfo(lg.dtype).min)
        tot = tot + torch.nn.functional.cross_entropy(lg*20, torch.zeros(20, dtype=torch.long))
    tot.backward()
    print(rank, "diff", (g - leaves[rank].grad).abs().max().item(), "gnorm", g.norm().item(), flush=True)
    dist.barrier(); dist.destroy_process_group()
if __name__ == "__main__":
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    ps=[ctx.Process(target=worker,args=(r,port)) for r in range(W)]
    [p.start() for p in ps]; [p.join() for p in ps]
