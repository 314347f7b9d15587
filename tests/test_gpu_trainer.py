"""Train/eval harness end to end on one B200: a tiny HSTU learns a deterministic next-item chain
through the CUDA-graph step with a cosine schedule, evaluation runs the fused predict_topk -> collector ->
Recall/NDCG path, early stopping bookkeeping works, and a checkpoint restores weights + optimizer state exactly."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

N_ITEMS, L_CTX, P = 199, 8, 1


def _chain(start, n):
    seq = [int(start)]
    for _ in range(n - 1):
        seq.append((seq[-1] * 7 + 3) % (N_ITEMS - 1) + 1)     # a permutation of 1..N-1: the next item is learnable
    return seq


def _train_batches(cfg, n_batches, B, gen):
    out = []
    for _ in range(n_batches):
        starts = torch.randint(1, N_ITEMS, (B,), generator=gen)
        items = torch.tensor([_chain(s, L_CTX + P) for s in starts], dtype=torch.int64)
        neg = torch.randint(1, N_ITEMS, (B, 1, 16), generator=gen)
        mask = torch.ones(B, L_CTX + P, dtype=torch.int64)
        tags = torch.empty(B, L_CTX + P, 0, dtype=torch.int64)
        out.append((items, neg, mask, tags))
    return out


def _eval_batches(n_batches, B, gen, C=1):
    out = []
    for _ in range(n_batches):
        starts = torch.randint(1, N_ITEMS, (B,), generator=gen)
        full = torch.tensor([_chain(s, L_CTX + 1) for s in starts], dtype=torch.int64)
        seq, target = full[:, :L_CTX].contiguous(), full[:, L_CTX:].contiguous()
        u, p = torch.nonzero(seq, as_tuple=True)
        out.append(dict(item_seq=seq, item_target=target, target_tags=torch.ones(B, 1, C, dtype=torch.int64),
                        history_index=(u, seq[u, p]), positive_u=torch.arange(B).unsqueeze(1)))
    return out


def _build(tmp_path, total_iters=240):
    from b200rec import synth
    from b200rec.hstu import HSTU
    cfg = synth.make_config("A", n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
                            MAX_ITEM_LIST_LENGTH=L_CTX, train_batch_size=64, num_negatives=16 * 64, item_num=N_ITEMS + 1,
                            eval_batch_size=64)
    cfg["topk"] = [1, 5, 10]
    cfg.update(optim_args=dict(learning_rate=5e-3, weight_decay=0.0), scheduler_args=dict(type="cosine", warmup=0.05),
               total_iters=total_iters, eval_freq=80, stopping_step=5, valid_metric="recall@10", valid_metric_bigger=True,
               checkpoint_dir=str(tmp_path))
    torch.manual_seed(2020)
    model = HSTU(cfg, synth.make_dataload(cfg), compute_dtype=torch.bfloat16).cuda()
    return cfg, model


def test_fit_evaluate_checkpoint_roundtrip(tmp_path):
    from b200rec.trainer import Trainer
    cfg, model = _build(tmp_path)
    gen = torch.Generator().manual_seed(11)
    train = _train_batches(cfg, 40, 64, gen)
    valid = _eval_batches(3, 64, gen, cfg["eval_num_cats"])
    logs = []
    tr = Trainer(cfg, model, log=logs.append)
    before = tr.evaluate(valid)["pred_0"]["recall@10"]
    best, best_result = tr.fit(train, valid)
    assert tr.train_step == 240 and len(tr.results) == 3 and set(best_result) == {"pred_0"}
    after = best_result["pred_0"]
    assert set(after) >= {"recall@1", "recall@5", "recall@10", "ndcg@10"}
    assert after["recall@10"] > 0.8 > 0.3 > before, (before, after)
    assert best == after["recall@10"]
    # checkpoint round trip: weights and optimizer moments come back bit-exactly
    path = tr.save_checkpoint(os.path.join(str(tmp_path), "ckpt.pth"))
    ref_sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref_m = tr.optimizer._st(model.item_embedding.weight)[0].clone()
    res1 = tr.evaluate(valid)
    with torch.no_grad():
        for p_ in model.parameters():
            p_.add_(0.05 * torch.randn_like(p_))
    model.invalidate_shadows()
    tr.load_checkpoint(path)
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    assert torch.equal(tr.optimizer._st(model.item_embedding.weight)[0], ref_m)
    assert tr.evaluate(valid) == res1
    # training continues from the restored state through the same captured graph
    tr.total_iters += 10
    tr.fit(train, None, saved=False)
    assert tr.train_step == 250


def test_early_stopping_stops_fit(tmp_path):
    """trainer.py:596-609,683-689: stop once the valid score failed to reach the best `stopping_step + 1` times."""
    from b200rec.trainer import Trainer
    cfg, model = _build(tmp_path, total_iters=400)
    cfg["eval_freq"], cfg["stopping_step"] = 10, 1
    gen = torch.Generator().manual_seed(12)
    tr = Trainer(cfg, model)
    scores = iter([0.5, 0.4, 0.3, 0.9])
    tr.evaluate = lambda data: {"pred_0": {"recall@10": next(scores)}}
    best, _ = tr.fit(_train_batches(cfg, 5, 64, gen), [None], saved=False)
    assert tr.train_step == 30 and tr.no_improve_times == 2 and best == 0.5
