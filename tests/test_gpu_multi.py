"""World-size-2 parity of the multi-GPU path against a TWO-RANK run of the unmodified reference
(tests/golden/multi2_*.pt, written by tests/golden/make_golden_multi.py under gloo on CPU).

Never skipped: with >= 2 visible GPUs the ranks use NCCL on cuda:0 / cuda:1; on a one-GPU box both ranks share
cuda:0 and talk over gloo (collectives on CUDA tensors are staged through the host by a test-only shim), which
exercises exactly the same routing: row-sharded table (all-to-all lookups / gradient rows), all-gathered negative
ids, dense all-reduce, sharded eval with the cross-shard top-K merge."""
import os
import socket
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

from conftest import load_golden, ROOT

pytestmark = pytest.mark.gpu

W = 2


# ------------------------------------------------------------------------------------------------- launch helpers
def _stage_collectives_through_host():
    """gloo transport for CUDA tensors (two ranks on ONE GPU cannot form an NCCL communicator)."""
    import torch.distributed as dist

    class _Done(object):
        def wait(self):
            return True

    o_ag, o_a2a, o_ar = dist.all_gather, dist.all_to_all_single, dist.all_reduce

    def all_gather(out_list, t, group=None, async_op=False):
        if not t.is_cuda:
            return o_ag(out_list, t, group=group, async_op=async_op)
        host = [torch.empty(o.shape, dtype=o.dtype) for o in out_list]
        o_ag(host, t.cpu(), group=group)
        for o, h in zip(out_list, host):
            o.copy_(h)
        return _Done()

    def all_to_all_single(output, input, output_split_sizes=None, input_split_sizes=None, group=None, async_op=False):
        if not input.is_cuda:
            return o_a2a(output, input, output_split_sizes, input_split_sizes, group=group, async_op=async_op)
        host = torch.empty(output.shape, dtype=output.dtype)
        o_a2a(host, input.cpu().contiguous(), output_split_sizes, input_split_sizes, group=group)
        output.copy_(host)
        return _Done()

    def all_reduce(t, op=dist.ReduceOp.SUM, group=None, async_op=False):
        if not t.is_cuda:
            return o_ar(t, op=op, group=group, async_op=async_op)
        host = t.cpu()
        o_ar(host, op=op, group=group)
        t.copy_(host)
        return _Done()

    dist.all_gather, dist.all_to_all_single, dist.all_reduce = all_gather, all_to_all_single, all_reduce


def _entry(rank, port, fn_name, out_dir, nccl):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=W, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=W)
        _stage_collectives_through_host()
    res = globals()[fn_name](dev, rank)
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _run_two_ranks(fn_name):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    nccl = torch.cuda.device_count() >= W
    out_dir = tempfile.mkdtemp()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_entry, args=(r, port, fn_name, out_dir, nccl)) for r in range(W)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0, f"rank process failed ({'nccl' if nccl else 'gloo, shared cuda:0'})"
    return [torch.load(os.path.join(out_dir, f"r{r}.pt"), weights_only=False) for r in range(W)]


def _model(name, dev, sparse=True, dtype=torch.float32):
    from b200rec import synth
    from b200rec.hstu import HSTU
    fx = load_golden(name)
    cfg = synth.Config(fx["cfg"])
    cfg["sparse_embedding_grad"] = sparse
    dl = synth.Dataload(cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    model = HSTU(cfg, dl, compute_dtype=dtype)
    model.load_state_dict(fx["state_dict"])
    return fx, cfg, model.to(dev).eval()


def _multi(name):
    return torch.load(os.path.join(ROOT, "tests", "golden", f"multi2_{name}.pt"), weights_only=False)


# ------------------------------------------------------------------------------------------------- rank bodies
def _rank_train_eval(dev, rank, name="prior_additive", sharded=True):
    from b200rec import parallel
    fx, cfg, model = _model(name, dev)
    mfx = _multi(name)
    if sharded:
        model.shard_item_table()
    batch = tuple(t.to(dev) for t in mfx["batches"][rank])
    out = model(batch)
    out["loss"].backward()
    parallel.DataParallel(model).sync_gradients()
    uid, urows, nu = model.emb_grad
    k = int(nu.item())
    res = dict(logs={kk: float(v) for kk, v in out.items()},
               grads={n: p.grad.cpu() for n, p in model.named_parameters() if p.grad is not None},
               emb_ids=uid[:k].cpu(), emb_rows=urows[:k].cpu())
    if sharded:
        ev = fx["eval_batch"]
        tags = fx["item_tags"].t().contiguous().to(dev)[:, rank::W].contiguous()
        feat = model.compute_item_all()
        hu, hi = ev["history_index"]
        # ranks hold different user counts (ragged last batch): rank 1 drops its last user
        nb = ev["item_seq"].shape[0] - rank
        keep = hu < nb
        idx, val, hs = model.predict_topk(ev["item_seq"][:nb].to(dev), feat, tags, ev["target_tags"][:nb].to(dev),
                                          history_index=(hu[keep].to(dev), hi[keep].to(dev)), K=max(cfg["topk"]))
        res["topk"] = idx.cpu()
    return res


def _rank_prior_additive(dev, rank):
    return _rank_train_eval(dev, rank, "prior_additive")


def _rank_nce_pred4(dev, rank):
    return _rank_train_eval(dev, rank, "nce_pred4")


def _rank_prior_mult(dev, rank):
    return _rank_train_eval(dev, rank, "prior_mult")


def _rank_replicated(dev, rank):
    return _rank_train_eval(dev, rank, "prior_additive", sharded=False)


def _check_against_reference(name, res, sharded=True):
    from oracle import hstu_oracle as orc
    fx, mfx = load_golden(name), _multi(name)
    g_ref = mfx["grads"]
    # item table: the NCCL branch's reduce-scatter SUM semantics (`grads_rs`, see make_golden_multi.py: torch's gloo
    # emulation of the all_gather backward mis-delivers element-wise gradients, so the gloo run pins everything but this)
    emb_ref = mfx["grads_rs"]["item_embedding.weight"]
    for r in range(W):
        want = mfx["logs"][r]
        got = res[r]["logs"]
        assert abs(got["loss"] - want["loss"]) <= 2e-5 * max(1.0, abs(want["loss"])), (r, got["loss"], want["loss"])
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (r, k, got[k], v)
        for k, g in g_ref.items():
            if g is None or k == "item_embedding.weight":
                continue
            mine = res[r]["grads"][k]
            assert (mine - g).abs().max().item() < 2e-4 * max(1e-6, g.abs().max().item()), (r, k)
        got_emb = torch.zeros_like(emb_ref)
        ids = res[r]["emb_ids"]
        if sharded:                                   # compact gradient of THIS rank's shard, local row indices
            got_emb[ids * W + r] = res[r]["emb_rows"]
            want_emb = torch.zeros_like(emb_ref)
            want_emb[r::W] = emb_ref[r::W]
        else:                                         # replicated table: every rank holds the full averaged gradient
            got_emb[ids] = res[r]["emb_rows"]
            want_emb = emb_ref
        assert (got_emb - want_emb).abs().max().item() < 2e-4 * emb_ref.abs().max().item(), r
        assert set(torch.nonzero(got_emb.abs().sum(1) > 0).flatten().tolist()) == \
            set(torch.nonzero(want_emb.abs().sum(1) > 0).flatten().tolist())           # exact gradient row set
        if sharded:
            ref_idx, _, _ = orc.collect_topk(fx["scores"], max(fx["cfg"]["topk"]), fx["cfg"]["split_mode"])
            nb = ref_idx.shape[0] - r
            assert (res[r]["topk"].numpy() == ref_idx[:nb]).all(), r


@pytest.mark.parametrize("name", ["prior_additive", "nce_pred4", "prior_mult"])
def test_two_rank_sharded_training_and_eval_match_two_rank_reference(name):
    _check_against_reference(name, _run_two_ranks(f"_rank_{name}"))


def test_two_rank_replicated_table_matches_two_rank_reference():
    _check_against_reference("prior_additive", _run_two_ranks("_rank_replicated"), sharded=False)


# ------------------------------------------------------------------------------------------------- steppers
def _graphed_vs_eager(dev, rank, world, steps=3):
    """Max relative parameter difference after `steps` AdamW steps: eager sharded step vs pre/graph/post."""
    from b200rec import synth, parallel
    from b200rec.optim import FusedAdamW
    from b200rec.graphed import GraphedShardedStep
    fx = load_golden("prior_additive")
    cfg = synth.Config(fx["cfg"])
    batches = [tuple(t.to(dev) for t in synth.make_train_batch(cfg, seed=70 + 10 * i + rank, rank=rank, world_size=world,
                                                                 item_tags=fx["item_tags"], zipf=False))
               for i in range(steps)]
    Lc = cfg["MAX_ITEM_LIST_LENGTH"]
    models = []
    for mode in ("eager", "graph"):
        _, _, model = _model("prior_additive", dev)
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


def _rank_graphed(dev, rank):
    return _graphed_vs_eager(dev, rank, W)


def test_graphed_sharded_step_single_gpu():
    assert _graphed_vs_eager(torch.device("cuda:0"), 0, 1) < 2e-4


def test_graphed_sharded_step_two_ranks():
    for worst in _run_two_ranks("_rank_graphed"):
        assert worst < 2e-4


def _rank_trainer_replicated(dev, rank):
    """Trainer with world > 1 and a replicated table must synchronise gradients (eager step), not
    replay the single-GPU graph: after a few steps both ranks hold identical parameters."""
    import torch.distributed as dist
    from b200rec import synth
    from b200rec.trainer import Trainer
    fx, cfg, model = _model("prior_additive", dev)
    cfg.update(total_iters=3, optim_args=dict(learning_rate=1e-2, weight_decay=0.0), eval_freq=0)
    tr = Trainer(cfg, model, use_graph=True)
    assert tr.use_graph is False

    def dense_flat():
        return torch.cat([p.detach().reshape(-1) for n, p in model.named_parameters() if "item_embedding" not in n])

    before = dense_flat().clone()
    batches = [synth.make_train_batch(cfg, seed=300 + 7 * i + rank, rank=rank, world_size=W, item_tags=fx["item_tags"],
                                      zipf=False) for i in range(3)]
    tr.fit(batches, None, saved=False)
    flat = dense_flat()
    other = [torch.empty_like(flat) for _ in range(W)]
    dist.all_gather(other, flat)
    return dict(diff=(other[0] - other[1]).abs().max().item(), moved=(flat - before).abs().max().item())


def test_trainer_two_ranks_replicated_table_stays_in_sync():
    for r in _run_two_ranks("_rank_trainer_replicated"):
        assert r["diff"] == 0.0 and r["moved"] > 0.0
