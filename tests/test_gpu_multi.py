"""2-GPU NCCL test (skipped with fewer devices): row-sharded table + all-gathered negatives + dense
all-reduce reproduce the rank-averaged gradients of the same step emulated on one GPU, and sharded
eval returns the single-GPU top-K."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from b200rec import synth, parallel
    from b200rec.hstu import HSTU
    fx = load_golden("prior_additive")
    cfg = synth.Config(fx["cfg"])
    cfg["sparse_embedding_grad"] = True
    dl = synth.Dataload(cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    model = HSTU(cfg, dl, compute_dtype=torch.float32)
    model.load_state_dict(fx["state_dict"])
    model = model.to(dev).eval()
    model.shard_item_table()
    batch = tuple(t.to(dev) for t in synth.make_train_batch(cfg, seed=40 + rank, item_tags=fx["item_tags"], zipf=False))
    out = model(batch)
    out["loss"].backward()
    parallel.DataParallel(model).sync_gradients()
    lid, lrows, nu = model.emb_grad
    k = int(nu.item())
    ev = fx["eval_batch"]
    C = cfg["eval_num_cats"]
    tags = fx["item_tags"].t().contiguous().to(dev)[:, rank::world].contiguous()
    feat = model.compute_item_all()
    hu, hi = ev["history_index"]
    idx, val, hs = model.predict_topk(ev["item_seq"].to(dev), feat, tags, ev["target_tags"].to(dev),
                                      history_index=(hu.to(dev), hi.to(dev)), K=max(cfg["topk"]))
    q.put((rank, float(out["loss"]), model._hstu._attention_layers[0]._uvqk.grad.cpu(), lid[:k].cpu(), lrows[:k].cpu(),
           idx.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_training_and_eval():
    from b200rec import synth
    from b200rec.hstu import HSTU
    from oracle import hstu_oracle as orc
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    W = 2
    procs = [ctx.Process(target=_worker, args=(r, W, port, q)) for r in range(W)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(W)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-GPU emulation: rank r's batch against the concatenated negatives of both ranks
    fx = load_golden("prior_additive")
    cfg = synth.Config(fx["cfg"])
    dl = synth.Dataload(cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    dev = torch.device("cuda:0")
    batches = [synth.make_train_batch(cfg, seed=40 + r, item_tags=fx["item_tags"], zipf=False) for r in range(W)]
    neg_all = torch.cat([b[1] for b in batches], dim=0)
    g_uvqk, g_emb, losses = 0, 0, []
    for r in range(W):
        m = HSTU(cfg, dl, compute_dtype=torch.float32)
        m.load_state_dict(fx["state_dict"])
        m = m.to(dev).eval()
        items, _, mask, tags = batches[r]
        out = m((items.to(dev), neg_all.to(dev), mask.to(dev), tags.to(dev)))
        out["loss"].backward()
        losses.append(float(out["loss"]))
        g_uvqk = g_uvqk + m._hstu._attention_layers[0]._uvqk.grad.cpu() / W
        g_emb = g_emb + m.item_embedding.weight.grad.cpu() / W
    for r, loss, gu, lid, lrows, idx in res:
        assert abs(loss - losses[r]) < 1e-5 * max(1.0, abs(losses[r]))
        assert (gu - g_uvqk).abs().max().item() < 2e-4 * g_uvqk.abs().max().item()
        got = torch.zeros_like(g_emb)
        got[lid * W + r] = lrows
        want = torch.zeros_like(g_emb)
        want[r::W] = g_emb[r::W]
        assert (got - want).abs().max().item() < 2e-4 * g_emb.abs().max().item()
    ref_idx, _, _ = orc.collect_topk(fx["scores"], max(cfg["topk"]), cfg["split_mode"])
    for r, *_, idx in res:
        assert (idx.numpy() == ref_idx).all()


def _graph_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    q.put((rank, _graphed_vs_eager(dev, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def _graphed_vs_eager(dev, rank, world, steps=3):
    """Max relative parameter difference after `steps` AdamW steps: eager sharded step vs pre/graph/post."""
    from b200rec import synth, parallel
    from b200rec.hstu import HSTU
    from b200rec.optim import FusedAdamW
    from b200rec.graphed import GraphedShardedStep
    fx = load_golden("prior_additive")
    cfg = synth.Config(fx["cfg"])
    cfg["sparse_embedding_grad"] = True
    dl = synth.Dataload(cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    batches = [tuple(t.to(dev) for t in synth.make_train_batch(cfg, seed=70 + 10 * i + rank, item_tags=fx["item_tags"],
                                                                 zipf=False)) for i in range(steps)]
    Lc = cfg["MAX_ITEM_LIST_LENGTH"]
    models = []
    for mode in ("eager", "graph"):
        model = HSTU(cfg, dl, compute_dtype=torch.float32)
        model.load_state_dict(fx["state_dict"])
        model = model.to(dev).eval()
        model.shard_item_table()
        opt = FusedAdamW(model, lr=1e-2, weight_decay=0.01)
        if mode == "eager":
            dp = parallel.DataParallel(model, opt)
            for b in batches:
                opt.zero_grad()
                model(b)["loss"].backward()
                dp.sync_gradients()
                opt.step()
        else:
            stepper = GraphedShardedStep(model, opt, batches[0], bucket=32)
            for b in batches:
                stepper(b, int(b[2][:, :Lc].sum()))
            stepper.flush()
        models.append(model)
    torch.cuda.synchronize()
    worst = 0.0
    for (n, a), (_, b) in zip(models[0].named_parameters(), models[1].named_parameters()):
        worst = max(worst, (a - b).abs().max().item() / max(a.abs().max().item(), 1e-6))
    return worst


def test_graphed_sharded_step_single_gpu():
    assert _graphed_vs_eager(torch.device("cuda:0"), 0, 1) < 2e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_graphed_sharded_step_two_gpus():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    W = 2
    procs = [ctx.Process(target=_graph_worker, args=(r, W, port, q)) for r in range(W)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(W)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, worst in res:
        assert worst < 2e-4
