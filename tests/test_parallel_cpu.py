"""World-size-2 gloo test of the data-parallel gradient exchange protocol (host logic only; the CUDA
sorted-segment merge is replaced by a torch reference so the exchange / padding / ordering can be
checked on CPU)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _ref_merge(ids_list, rows_list, D, scale):
    ids = torch.cat(ids_list)
    rows = torch.cat(rows_list) * scale
    keep = ids > 0
    uniq, inv = torch.unique(ids[keep], return_inverse=True)
    out = torch.zeros(uniq.numel(), D).index_add_(0, inv, rows[keep])
    return uniq, out, torch.tensor([uniq.numel()], dtype=torch.int32)


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.item_embedding = torch.nn.Embedding(50, 8)
        self.lin = torch.nn.Linear(8, 8)
        self.unused = torch.nn.Parameter(torch.zeros(3))
        self.emb_grad = None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200rec import parallel
    parallel.merge_compact_rows = _ref_merge
    torch.manual_seed(0)
    m = _Toy()
    g = torch.Generator().manual_seed(100 + rank)
    m.lin.weight.grad = torch.randn(8, 8, generator=g)
    m.lin.bias.grad = torch.randn(8, generator=g)
    k = 3 + rank                                  # ragged per-rank unique-row counts
    ids = torch.tensor([5, 7, 9, 11][:k]) + rank  # overlapping ids across ranks
    rows = torch.randn(k, 8, generator=g)
    pad = 6
    uid = torch.cat([ids, torch.full((pad - k,), -7)])
    urows = torch.cat([rows, torch.zeros(pad - k, 8)])
    m.emb_grad = (uid, urows, torch.tensor([k], dtype=torch.int32))
    dp = parallel.DataParallel(m)
    dp.sync_gradients()
    uniq, merged, nu = m.emb_grad
    q.put((rank, m.lin.weight.grad.clone(), m.lin.bias.grad.clone(), uniq.clone(), merged.clone(), ids, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, b0, u0, m0, ids0, rows0), (r1, w1, b1, u1, m1, ids1, rows1) = res
    # dense grads: identical on both ranks and equal to the mean of the per-rank grads
    assert torch.equal(w0, w1) and torch.equal(b0, b1)
    g0 = torch.Generator().manual_seed(100)
    g1 = torch.Generator().manual_seed(101)
    ew = (torch.randn(8, 8, generator=g0) + torch.randn(8, 8, generator=g1)) / 2
    assert torch.allclose(w0, ew, atol=1e-6)
    # table grads: same merged rows on both ranks == mean over ranks of the dense scatter
    assert torch.equal(u0, u1) and torch.equal(m0, m1)
    dense = torch.zeros(50, 8)
    dense.index_add_(0, ids0, rows0)
    dense.index_add_(0, ids1, rows1)
    dense /= 2
    assert torch.allclose(m0, dense[u0], atol=1e-6)
    assert set(u0.tolist()) == set(ids0.tolist()) | set(ids1.tolist())
