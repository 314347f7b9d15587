"""Generates tests/golden/multi2_*.pt: a TWO-RANK run of the LIVE, UNMODIFIED reference under gloo on CPU
(torch.distributed.nn.functional.all_gather forward and backward work with gloo).
Usage:  python tests/golden/make_golden_multi.py        (build container only: needs /root/reference)

Each rank builds the reference HSTU from the state dict of the single-rank fixture, runs its own seeded batch
(negatives per sample = ceil(num_negatives / W / B), trainset.py:58-60), the model all-gathers the normalised
negatives of both ranks (basemodel.py:11-22 via hstu.py:673,755), and the per-rank gradients are averaged the way
DDP / ZeRO-2 do (trainer.py:434-453).  The fixture stores per-rank batches, per-rank losses / logging scalars and
the rank-averaged gradient of every parameter (`grads`).

Caveat found while building this fixture: under GLOO, torch emulates the backward of the autograd all_gather with
scatter calls (torch/distributed/nn/functional.py `_AllGather.backward`, non-NCCL branch) and that emulation does NOT
return sum_i dL_i/d(slot) for element-wise non-uniform gradients (a 30-line torch-only reproduction, no reference code
involved, disagrees with the analytic gradient; finite differences of the 2-rank reference losses agree with the
analytic one).  Losses, logging scalars and every DENSE gradient of the 2-rank run are unaffected (they do not pass
through that backward) and are pinned by `grads`.  For the item table the fixture additionally stores `grads_rs`: the
gradient with the NCCL branch's semantics (reduce-scatter SUM, what the reference computes on GPUs), produced by the
unmodified reference model in a single process, rank by rank, with both ranks' negative ids in its negative slot
(the same global negative SET; the loss is invariant to the order of negatives), then averaged over ranks.
"""
import os
import socket
import sys
import tempfile

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ["prior_additive", "nce_pred4", "prior_mult"]
W = 2


def worker(rank, port, name, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=W)
    torch.set_num_threads(2)
    from oracle import ref_harness as rh
    from b200rec import synth
    from conftest import load_golden
    fx = load_golden(name)
    cfg = synth.Config(fx["cfg"])
    model = rh.build_reference_model(dict(cfg), cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    model.load_state_dict(fx["state_dict"])
    model.eval()
    batch = synth.make_train_batch(cfg, seed=40 + rank, rank=rank, world_size=W, item_tags=fx["item_tags"], zipf=False)
    out = model(batch)
    out["loss"].backward()
    grads = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            grads[k] = None
            continue
        g = p.grad.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        grads[k] = g / W
    torch.save((rank, batch, {k: float(v) for k, v in out.items()}, grads if rank == 0 else None),
               os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def reduce_scatter_semantics(name, batches):
    """Rank-averaged gradients with sum-over-ranks semantics for the gathered negatives, from the reference model run
    in ONE process (world size 1: all_gather is the identity, basemodel.py:21-22)."""
    from oracle import ref_harness as rh
    from b200rec import synth
    from conftest import load_golden
    fx = load_golden(name)
    cfg = synth.Config(fx["cfg"])
    neg_all = torch.cat([b[1] for b in batches], dim=0)                     # [W*B, sets, n]
    B = batches[0][0].shape[0]
    # every sample carries W*n negative ids; flattened over the batch this is the global set of the 2-rank run
    neg = neg_all.view(W, B, neg_all.shape[1], neg_all.shape[2]).permute(1, 2, 0, 3).reshape(B, neg_all.shape[1], -1)
    tot, losses = {}, []
    for r in range(W):
        model = rh.build_reference_model(dict(cfg), cfg["item_num"], fx["category_counts"], fx["category_to_int"])
        model.load_state_dict(fx["state_dict"])
        model.eval()
        items, _, mask, tags = batches[r]
        out = model((items, neg, mask, tags))
        out["loss"].backward()
        losses.append(float(out["loss"]))
        for k, p in model.named_parameters():
            if p.grad is not None:
                tot[k] = tot.get(k, 0) + p.grad / W
    return tot, losses


def main():
    from oracle import ref_harness as rh
    assert rh.available(), "needs /root/reference"
    for name in CASES:
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        ctx = mp.get_context("spawn")
        tmp = tempfile.mkdtemp()
        procs = [ctx.Process(target=worker, args=(r, port, name, tmp)) for r in range(W)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
            assert p.exitcode == 0
        res = [torch.load(os.path.join(tmp, f"r{r}.pt"), weights_only=False) for r in range(W)]
        grads_rs, losses_rs = reduce_scatter_semantics(name, [r[1] for r in res])
        for r in range(W):                          # same forward: the single-process run reproduces each rank's loss
            assert abs(losses_rs[r] - res[r][2]["loss"]) < 1e-5 * abs(res[r][2]["loss"]), (losses_rs, res[r][2]["loss"])
        for k, g in res[0][3].items():              # dense gradients agree between the two computations
            if g is not None and k != "item_embedding.weight":
                assert (g - grads_rs[k]).abs().max().item() < 1e-4 * max(1e-6, g.abs().max().item()), k
        fx = dict(world=W, base_fixture=name, batches=[r[1] for r in res], logs=[r[2] for r in res], grads=res[0][3],
                  grads_rs={"item_embedding.weight": grads_rs["item_embedding.weight"]})
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"multi2_{name}.pt")
        torch.save(fx, path)
        print("wrote", path, os.path.getsize(path), "bytes; losses", [r[2]["loss"] for r in res])


if __name__ == "__main__":
    main()
