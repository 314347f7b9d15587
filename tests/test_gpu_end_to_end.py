"""The pieces a reference user touches, together, on one B200: interaction data in HBM -> GPU batch construction
(`GpuTrainBatcher`, `GpuEvalBatcher`) -> `Trainer.fit` (CUDA-graph step, cosine schedule, lazy table AdamW, periodic
full-catalogue evaluation with early-stopping bookkeeping) -> Recall/NDCG of the held-out next item."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _TrainLoader(object):
    """Epoch = a shuffled pass over the training windows (what DataLoader + DistributedSampler do in the reference)."""

    def __init__(self, batcher, n_samples, batch_size, seed=0):
        self.batcher, self.n, self.B, self.seed, self.step = batcher, n_samples, batch_size, seed, 0

    def __iter__(self):
        order = np.random.default_rng(self.seed + self.step).permutation(self.n)
        for i in range(0, self.n - self.B + 1, self.B):
            (batch, _n_tok) = self.batcher.batch(order[i:i + self.B], step=self.step)
            self.step += 1
            yield batch


def test_batcher_trainer_evaluator_learn_a_next_item_rule(tmp_path):
    from b200rec import synth
    from b200rec.batcher import GpuEvalBatcher, GpuTrainBatcher, InteractionData
    from b200rec.hstu import HSTU
    from b200rec.trainer import Trainer
    N, L, n_users = 300, 10, 400
    rng = np.random.default_rng(3)
    user_seq, train_len = [[]], [0]
    for _ in range(n_users):                       # next item = (7 * item + 3) mod (N - 1) + 1: learnable from the last item
        s = [int(rng.integers(1, N))]
        for _ in range(int(rng.integers(6, 22))):
            s.append((s[-1] * 7 + 3) % (N - 1) + 1)
        user_seq.append(s)
        train_len.append(len(s) - 1)               # the last item is the validation target
    cfg = synth.make_config("A", n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
                            MAX_ITEM_LIST_LENGTH=L, train_batch_size=64, num_negatives=64 * 20, item_num=N,
                            eval_batch_size=100)
    cfg["topk"] = [1, 5, 10]
    cfg.update(optim_args=dict(learning_rate=5e-3, weight_decay=0.0), scheduler_args=dict(type="cosine", warmup=0.05),
               total_iters=300, eval_freq=100, stopping_step=5, valid_metric="recall@10", valid_metric_bigger=True,
               checkpoint_dir=str(tmp_path), pad_random_sample=True)
    data = InteractionData(user_seq, train_len, N, L)
    torch.manual_seed(2020)
    model = HSTU(cfg, synth.make_dataload(cfg), compute_dtype=torch.bfloat16).cuda()
    train = _TrainLoader(GpuTrainBatcher(data, cfg), len(data), 64)
    ev = GpuEvalBatcher(data, cfg)
    uids = np.arange(1, n_users + 1)
    valid = [ev.batch(uids[i:i + 100], "valid") for i in range(0, n_users, 100)]
    for b in valid:                                # category-free config: one all-ones tag column per target
        b["target_tags"] = torch.ones(b["item_seq"].shape[0], cfg["eval_pred_len"], cfg["eval_num_cats"],
                                      dtype=torch.int64, device="cuda")
    tr = Trainer(cfg, model)
    before = tr.evaluate(valid)["pred_0"]
    best, best_result = tr.fit(train, valid, saved=False)
    after = best_result["pred_0"]
    assert tr.train_step == 300 and tr.optimizer.lazy_table
    assert after["recall@10"] > 0.8 and after["ndcg@10"] > 0.5 and before["recall@10"] < 0.2, (before, after)
