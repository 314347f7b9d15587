"""GPU batch construction (b200rec.batcher) against a plain-Python restatement of the reference's per-sample logic
(trainset.py:99-177, evalset.py:81-155, collate_fn.py:59-90): everything deterministic is checked exactly; the
sampled parts (random pads, negatives) are checked against the sampling law the reference implements."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(n_users=60, N=400, L=10, C=4, seed=5, events=False):
    from b200rec.batcher import InteractionData
    rng = np.random.default_rng(seed)
    user_seq, train_len, ev = [[]], [0], [[]]
    for _ in range(n_users):
        n = int(rng.integers(4, 40))
        user_seq.append(rng.choice(np.arange(1, N), size=n, replace=False).tolist())
        ev.append(rng.integers(0, C, size=n).tolist())
        train_len.append(n - 2)                       # last two items held out (valid / test)
    tags = torch.rand(N, C, generator=torch.Generator().manual_seed(seed)) < 0.4
    tags[torch.arange(N), torch.randint(0, C, (N,), generator=torch.Generator().manual_seed(seed + 1))] = True
    d = InteractionData(user_seq, train_len, N, L, item_tags=None if events else tags, event_seq=ev if events else None)
    return d, user_seq, train_len, tags, ev


def _cfg(L, P, C, loss="prior", by_cat=True, category_by="item", B=16, n_neg_total=16 * 12):
    return dict(MAX_ITEM_LIST_LENGTH=L, pred_len=P, eval_pred_len=2, loss=loss, neg_sample_by_cat=by_cat,
                category_by=category_by, num_negatives=n_neg_total, train_batch_size=B, eval_num_cats=C, seed=9,
                pad_random_sample=True)


@pytest.mark.parametrize("events", [False, True], ids=["item-tags", "event-tags"])
def test_train_batch_layout_and_sampling_law(events):
    from b200rec.batcher import GpuTrainBatcher
    L, P, C, N = 10, 3, 4, 400
    d, user_seq, train_len, tags, ev = _data(N=N, L=L, C=C, events=events)
    cfg = _cfg(L, P, C, category_by="event" if events else "item", by_cat=not events)
    bt = GpuTrainBatcher(d, cfg)
    idx = np.arange(len(d))[:48]
    (items, neg, mask, tg), n_tok = bt.batch(idx, step=3)
    (items2, neg2, _, _), _ = bt.batch(idx, step=3)
    (items3, neg3, _, _), _ = bt.batch(idx, step=4)
    assert torch.equal(items, items2) and torch.equal(neg, neg2)            # reproducible per (seed, step)
    assert not torch.equal(neg, neg3)                                        # a new stream per step
    items, neg, mask, tg = items.cpu(), neg.cpu(), mask.cpu(), tg.cpu()
    n_sets = (C + 1) if not events else 1
    assert neg.shape == (48, n_sets, 12) and n_tok == int(mask[:, :L].sum())
    for r, s in enumerate(idx):
        uid, end = int(d.h_sample_uid[s]), int(d.h_sample_end[s])
        start = max(0, end - L)
        pad = L - (end - start)
        pred = min(train_len[uid] - end, P)
        real = user_seq[uid][start:end + pred]                                # trainset.py:155-160
        row = items[r].tolist()
        assert row[pad:pad + len(real)] == real
        assert mask[r].tolist() == [0] * pad + [1] * len(real) + [0] * (P - pred)
        pads = row[:pad] + row[pad + len(real):]
        assert all(1 <= x < N for x in pads) and not (set(pads) & set(real)) and len(set(pads)) == len(pads)
        if events:
            want = torch.zeros(L + P, C, dtype=torch.int64)
            for j, e in enumerate(ev[uid][start:end + pred]):
                want[pad + j, e] = 1                                          # trainset.py:147-153
        else:
            want = tags[items[r]].to(torch.int64)                             # trainset.py:165-167 (pads included)
        assert torch.equal(tg[r], want)
        for s_ in range(n_sets):
            ns = neg[r, s_].tolist()
            assert len(set(ns)) == len(ns) and not (set(ns) & set(row))        # distinct, outside the padded row
            assert all(1 <= x < N for x in ns)
            if not events and s_ < C:
                assert all(bool(tags[x, s_]) for x in ns)                     # category pool s_
    # uniformity of the global set: every item id is hit with roughly equal frequency over many steps
    cnt = torch.zeros(N)
    for step in range(40):
        (_, ng, _, _), _ = bt.batch(idx, step=100 + step)
        cnt += torch.bincount(ng[:, -1].reshape(-1).cpu(), minlength=N).float()
    mean = float(cnt[1:].mean())                      # ~58 draws per item; binomial sd ~ sqrt(mean)
    assert cnt[0] == 0 and float((cnt[1:] - mean).abs().max()) < 5.0 * mean ** 0.5


@pytest.mark.parametrize("phase", ["valid", "test"])
def test_eval_batch_matches_reference_collate(phase):
    from b200rec.batcher import GpuEvalBatcher
    L, C, N, Pe = 10, 4, 400, 2
    d, user_seq, train_len, tags, _ = _data(N=N, L=L, C=C)
    cfg = _cfg(L, 3, C)
    uids = np.arange(1, 41)
    out = GpuEvalBatcher(d, cfg).batch(uids, phase)
    hu, hi = out["history_index"]
    seqs, tgts, hist_u, hist_i = [], [], [], []
    for r, u in enumerate(uids):
        s = user_seq[u]
        hist = s[:train_len[u]] if phase == "valid" else s[:-Pe]             # evalset.py:83-91
        tgt = s[train_len[u]:train_len[u] + Pe] if phase == "valid" else s[-Pe:]
        w = hist[-L:]
        seqs.append([0] * (L - len(w)) + w)                                   # left zero padding
        tgts.append(tgt)
        hist_u += [r] * len(hist)
        hist_i += hist
    assert out["item_seq"].cpu().tolist() == seqs and out["item_target"].cpu().tolist() == tgts
    assert hu.cpu().tolist() == hist_u and hi.cpu().tolist() == hist_i       # collate_fn.py:76-77
    assert torch.equal(out["target_tags"].cpu(), tags[torch.tensor(tgts)].to(torch.int64))
    assert out["positive_u"].tolist() == [[r] * Pe for r in range(len(uids))]


def test_batcher_feeds_the_model():
    from b200rec import synth
    from b200rec.batcher import GpuTrainBatcher, InteractionData
    from b200rec.hstu import HSTU
    cfg = synth.make_config("B", n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
                            MAX_ITEM_LIST_LENGTH=12, train_batch_size=8, num_negatives=8 * 6, item_num=400)
    C = cfg["eval_num_cats"]
    rng = np.random.default_rng(1)
    user_seq = [[]] + [rng.choice(np.arange(1, 400), size=int(rng.integers(6, 50)), replace=False).tolist() for _ in range(30)]
    train_len = [0] + [len(s) - 2 for s in user_seq[1:]]
    tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
    d = InteractionData(user_seq, train_len, 400, cfg["MAX_ITEM_LIST_LENGTH"], item_tags=tags)
    cfg["pad_random_sample"] = True
    batch, n_tok = GpuTrainBatcher(d, cfg).batch(np.arange(8), step=0)
    model = HSTU(cfg, synth.make_dataload(cfg), compute_dtype=torch.float32).cuda().eval()
    out = model(batch)
    out["loss"].backward()
    assert torch.isfinite(out["loss"]) and model.item_embedding.weight.grad.abs().sum() > 0
    assert batch[1].shape[1] == C + 1


# ---------------------------------------------------------------------------------------------------------------
# Against the LIVE reference datasets: tests/golden/data_layer.pt holds every row SEQTrainDataset.__getitem__
# (trainset.py:99-177) and SeqEvalDataset + seq_eval_collate (evalset.py:81-155, collate_fn.py:59-90) produced for
# a synthetic interaction set (tests/golden/make_golden_data.py).  Deterministic fields must be identical; the
# sampled fields (random pads, negatives: numpy RNG there, Philox here) must obey the same law on both sides.
def _fixture():
    import os
    from conftest import ROOT
    return torch.load(os.path.join(ROOT, "tests", "golden", "data_layer.pt"), weights_only=False)["data_layer"]


def _fixture_data(d, variant):
    from b200rec.batcher import InteractionData
    ev = variant == "event"
    return InteractionData(d["user_seq"], d["train_seq_len"], d["N"], d["L"], item_tags=None if ev else d["item_tags"],
                           event_seq=d["event_seq"] if ev else None, pred_len=d["P"], include_empty_context=True)


def _check_sampling_law(items, neg, mask, tags_table, N, by_cat):
    real = mask.bool()
    for r in range(items.shape[0]):
        row = items[r].tolist()
        seen = set(items[r][real[r]].tolist())
        pads = items[r][~real[r]].tolist()
        assert all(1 <= x < N for x in pads) and not (set(pads) & seen)          # trainset.py:111-122
        for s_ in range(neg.shape[1]):
            ns = neg[r, s_].tolist()
            assert len(set(ns)) == len(ns) and all(1 <= x < N for x in ns)        # rng.choice(replace=False)
            assert not (set(ns) & set(row))                                       # blacklist = the padded row (:126-131)
            if by_cat and s_ < neg.shape[1] - 1:
                assert all(bool(tags_table[x, s_]) for x in ns)                   # category pool s_


@pytest.mark.parametrize("variant", ["item_bycat", "event", "nce"])
def test_train_batch_matches_reference_dataset_rows(variant):
    from b200rec.batcher import GpuTrainBatcher
    d = _fixture()
    ref = d["train"][variant]
    cfg = dict(ref["config"])
    cfg["seed"] = 9
    data = _fixture_data(d, variant)
    assert len(data) == ref["items"].shape[0]
    (items, neg, mask, tg), n_tok = GpuTrainBatcher(data, cfg).batch(np.arange(len(data)), step=0)
    items, neg, mask, tg = items.cpu(), neg.cpu(), mask.cpu(), tg.cpu()
    assert torch.equal(mask, ref["mask"])                                          # trainset.py:133-134
    real = mask.bool()
    assert torch.equal(items[real], ref["items"][real])                            # the user's own items, in place
    assert neg.shape == ref["neg"].shape
    assert n_tok == int(mask[:, :d["L"]].sum())
    by_cat = variant == "item_bycat"
    for side in ((items, neg), (ref["items"], ref["neg"])):                        # same law on both sides
        _check_sampling_law(side[0], side[1], mask, d["item_tags"], d["N"], by_cat)
    if variant == "event":
        assert torch.equal(tg, ref["tags"])                                        # one-hot events, zero on pads (:147-153)
    elif variant == "item_bycat":
        assert torch.equal(tg[real], ref["tags"][real])
        assert torch.equal(tg, d["item_tags"][items].to(torch.int64))              # pads carry their own item's tags (:165-167)
        assert torch.equal(ref["tags"], d["item_tags"][ref["items"]].to(torch.int64))
    # marginal law of the global negative set: both samplers are uniform over [1, N)
    for ng in (neg[:, -1], ref["neg"][:, -1]):
        cnt = torch.bincount(ng.reshape(-1), minlength=d["N"]).float()
        mean = float(cnt[1:].mean())
        assert cnt[0] == 0 and float((cnt[1:] - mean).abs().max()) < 6.0 * max(1.0, mean) ** 0.5


@pytest.mark.parametrize("variant", ["item_bycat", "event"])
@pytest.mark.parametrize("phase", ["valid", "test"])
def test_eval_batch_matches_reference_collate_fixture(variant, phase):
    from b200rec.batcher import GpuEvalBatcher
    d = _fixture()
    ref = d["eval"][(variant, phase)]
    cfg = dict(d["train"][variant]["config"])
    data = _fixture_data(d, variant)
    out = GpuEvalBatcher(data, cfg).batch(ref["user_ids"].numpy(), phase)
    assert torch.equal(out["item_seq"].cpu(), ref["item_seq"])
    assert torch.equal(out["item_target"].cpu(), ref["item_target"])
    assert torch.equal(out["history_index"][0].cpu(), ref["history_u"])
    assert torch.equal(out["history_index"][1].cpu(), ref["history_i"])
    assert torch.equal(out["positive_u"], ref["positive_u"])
    assert torch.equal(out["target_tags"].cpu(), ref["target_tags"].to(torch.int64))
