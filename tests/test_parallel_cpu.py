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
    # numpy payloads are pickled by value; torch tensors travel as shared-memory handles that can vanish with the worker
    q.put((rank,) + tuple(t.detach().clone().numpy() for t in (m.lin.weight.grad, m.lin.bias.grad, uniq, merged, ids, rows)))
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
    res = [(r[0],) + tuple(torch.from_numpy(a) for a in r[1:]) for r in res]
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


# ---------------------------------------------------------------------------- sharded table routing
def _torch_row_gather(table, idx):
    return table[idx].clone()


def _torch_segment_reduce(ids, rows):
    keep = ids > 0
    uniq, inv = torch.unique(ids[keep], return_inverse=True)
    out = torch.zeros(uniq.numel(), rows.shape[1]).index_add_(0, inv, rows[keep])
    return uniq, out, torch.tensor([uniq.numel()], dtype=torch.int32)


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200rec import parallel
    N, D = 23, 4
    full = torch.arange(N * D, dtype=torch.float32).view(N, D)
    st = parallel.ShardedTable(parallel.ShardedTable.shard_of(full, world, rank), N, None,
                               row_gather=_torch_row_gather, segment_reduce=_torch_segment_reduce)
    g = torch.Generator().manual_seed(7 + rank)
    ids = torch.unique(torch.randint(0, N, (11,), generator=g))
    seen = []
    st.pre_gather = lambda idx: seen.append(idx.clone())          # hook of the lazy table update
    cache = torch.full((N, D), -1.0)
    rows = st.fetch(ids, out=cache)                                # rows land in the caller's static cache
    ok_fetch = torch.equal(rows, full[ids]) and rows.data_ptr() == cache.data_ptr() \
        and torch.equal(cache[ids.numel():], torch.full((N - ids.numel(), D), -1.0))
    # the owner was told exactly which local rows it is about to read (requests of BOTH ranks, local indices)
    ok_fetch = ok_fetch and len(seen) == 1 and bool((seen[0] * world + rank < N).all())
    grads = torch.randn(ids.numel(), D, generator=g)
    lid, lrows, nu = st.push_grads(grads, scale=0.5)
    negs = parallel.gather_negative_ids(torch.full((2, 3, 2), rank, dtype=torch.int64))
    q.put((rank, ok_fetch) + tuple(t.detach().clone().numpy() for t in (ids, grads, lid[: int(nu)], lrows[: int(nu)], negs)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_table_fetch_and_push_world2():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    res = [r[:2] + tuple(torch.from_numpy(a) for a in r[2:]) for r in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, D, W = 23, 4, 2
    dense = torch.zeros(N, D)
    for rank, ok, ids, grads, *_ in res:
        assert ok
        dense.index_add_(0, ids, grads)
    dense *= 0.5
    dense[0] = 0                                     # padding id 0 never gets a gradient
    for rank, ok, ids, grads, lid, lrows, negs in res:
        got = torch.zeros(N, D)
        got[lid * W + rank] = lrows                  # local row r of rank k is global id r*W + k
        want = torch.zeros(N, D)
        want[rank::W] = dense[rank::W]
        assert torch.allclose(got, want, atol=1e-6)
        assert negs.shape == (4, 3, 2) and torch.equal(negs[:2], torch.zeros(2, 3, 2, dtype=torch.int64)) \
            and torch.equal(negs[2:], torch.ones(2, 3, 2, dtype=torch.int64))


def test_merge_topk_tie_rule():
    from b200rec import parallel
    v0 = torch.tensor([[5.0, 3.0, float("-inf")]])
    i0 = torch.tensor([[4, 2, 0]])
    v1 = torch.tensor([[5.0, 4.0, 3.0]])
    i1 = torch.tensor([[1, 7, 9]])
    h = torch.zeros(1, 3, dtype=torch.int32)
    idx, val, _ = parallel.merge_topk([v0, v1], [i0, i1], [h, h], 4)
    assert idx.tolist() == [[1, 4, 7, 2]] and val.tolist() == [[5.0, 5.0, 4.0, 3.0]]


# ---------------------------------------------------------------------------- trainer metric reduction
def _reduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200rec import trainer as T

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(2))
            self.sharded_table = None

    cfg = dict(optim_args=dict(learning_rate=1e-2, weight_decay=0.0), total_iters=10, metrics_pred_len_list=[0],
               eval_pred_len=1, metric_decimal_place=4)
    tr = T.Trainer(cfg, M(), optimizer=object(), use_graph=False)
    # per-rank SUMS over users (metrics.py) + per-category (sum, count) tuples -> global means (trainer.py:1097-1123)
    res = {"recall@10": 3.0 + rank, "cat0-recall@10": (1.0 + rank, 2 + rank)}
    q.put((rank, tr._reduce(res, num_total=10 + 10 * rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_metric_reduction_world2():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        assert out == {"recall@10": round(7.0 / 30.0, 4), "cat0-recall@10": round(3.0 / 5.0, 4)}
