"""Host logic of the train/eval harness (b200rec.trainer): LR schedules against torch's LambdaLR driven by the
reference formulas (utils/lr_scheduler.py:44-116), early stopping and valid-score selection (utils/utils.py:60-122)."""
import math

import torch

from b200rec import trainer as T


def _lambda_lr_values(fn, base_lr, steps):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base_lr)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, fn)
    vals = []
    for _ in range(steps):
        vals.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    return vals


def test_schedules_match_lambda_lr():
    warm, total, base = 12.5, 100, 3e-3

    def ref_cos(step):                                   # lr_scheduler.py:105-113
        if step < warm:
            return float(step) / float(max(1, warm))
        progress = float(step - warm) / float(max(1, total - warm))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 0.5 * 2.0 * progress)))

    def ref_lin(step):                                   # lr_scheduler.py:66-74
        if step < warm:
            return float(step) / float(max(1, warm))
        return max(0.0, float(total - step) / float(max(1, total - warm)))

    for ref, mine in ((ref_cos, T.cosine_schedule_with_warmup), (ref_lin, T.linear_schedule_with_warmup)):
        want = _lambda_lr_values(ref, base, total)
        got = [base * mine(s, warm, total) for s in range(total)]
        assert max(abs(a - b) for a, b in zip(want, got)) < 1e-12


def test_trainer_lr_at_uses_warmup_fraction():
    cfg = dict(optim_args=dict(learning_rate=1e-2, weight_decay=0.0), scheduler_args=dict(type="cosine", warmup=0.1),
               total_iters=50, metrics_pred_len_list=[0], eval_pred_len=1)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(2))
            self.sharded_table = None

    tr = T.Trainer(cfg, M(), optimizer=object(), use_graph=False)
    assert tr.lr_at(0) == 0.0 and abs(tr.lr_at(5) - 1e-2) < 1e-15
    assert abs(tr.lr_at(50) - 0.0) < 1e-12 and tr.lr_at(20) < tr.lr_at(10)


def test_early_stopping_contract():
    best, cur, stop, upd = T.early_stopping(0.5, -float("inf"), 0, max_step=1, bigger=True)
    assert (best, cur, stop, upd) == (0.5, 0, False, True)
    best, cur, stop, upd = T.early_stopping(0.4, best, cur, max_step=1, bigger=True)
    assert (best, cur, stop, upd) == (0.5, 1, False, False)
    best, cur, stop, upd = T.early_stopping(0.3, best, cur, max_step=1, bigger=True)
    assert stop and not upd and cur == 2
    best, cur, stop, upd = T.early_stopping(0.2, 0.3, 0, max_step=3, bigger=False)
    assert upd and best == 0.2


def test_calculate_valid_score_picks_last_offset():
    res = {"pred_0": {"recall@10": 0.1, "ndcg@10": 0.05}, "pred_3": {"recall@10": 0.4, "ndcg@10": 0.2}}
    assert T.calculate_valid_score(res, 4, "NDCG@10") == 0.2
    assert T.calculate_valid_score(res, 4, None) == 0.4


def test_interaction_data_host_layout():
    """b200rec.batcher.InteractionData (host side, no kernels): CSR layout, per-category pools without id 0 and the
    reference's training-window locations (dataload.py:165-194 with max_item_list_len = L + 1)."""
    from b200rec.batcher import InteractionData
    L = 4
    user_seq = [[], list(range(1, 3)), list(range(1, 6)), list(range(1, 12)), list(range(1, 14)), [7]]
    train_len = [0, 2, 5, 11, 13, 1]
    tags = torch.tensor([[1, 0], [1, 0], [0, 1], [1, 1]] + [[0, 1]] * 10, dtype=torch.bool)   # item 0 tagged: must not enter pools
    d = InteractionData(user_seq, train_len, 14, L, item_tags=tags, device="cpu", include_empty_context=True)
    assert d.user_off.tolist() == [0, 0, 2, 7, 18, 31, 32] and d.user_seq[2:7].tolist() == [1, 2, 3, 4, 5]
    # n=2 -> (1, 1); n=5 <= L+1 -> (2, 4); n=11 -> offset 0: 0, 5, 10; n=13 -> offset 2: 2, 7, 12; n=1 -> none
    want = [(1, 1), (2, 4), (3, 0), (3, 5), (3, 10), (4, 2), (4, 7), (4, 12)]
    assert list(zip(d.h_sample_uid.tolist(), d.h_sample_end.tolist())) == want and len(d) == 8
    assert d.cat_off.tolist() == [0, 2, 14] and d.cat_items[:2].tolist() == [1, 3] and d.cat_items[2:].tolist() == list(range(2, 14))
    d3 = InteractionData(user_seq, train_len, 14, L, device="cpu")            # default: the empty-context window is dropped
    assert list(zip(d3.h_sample_uid.tolist(), d3.h_sample_end.tolist())) == [w for w in want if w[1] > 0]
    d2 = InteractionData(user_seq, train_len, 14, L, device="cpu", sample_last_only=True, pred_len=2)
    assert list(zip(d2.h_sample_uid.tolist(), d2.h_sample_end.tolist())) == [(1, 1), (2, 3), (3, 9), (4, 11)]


def test_get_model_resolves_like_the_reference():
    """REC/utils/utils.py:38-57: module <name.lower()>, attribute <name>; unknown names raise ValueError."""
    import pytest
    from b200rec.trainer import get_model
    from b200rec.hstu import HSTU
    from b200rec.comirec import ComiRec, REMI
    assert get_model("HSTU") is HSTU and get_model("ComiRec") is ComiRec and get_model("REMI") is REMI
    with pytest.raises(ValueError):
        get_model("SASRec")
