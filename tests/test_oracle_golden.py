"""Pins the restated oracle against fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import hstu_oracle as orc
from b200rec import synth


def _oracle(fx, requires_grad=False):
    cfg = synth.Config(fx["cfg"])
    sd = {}
    for k, v in fx["state_dict"].items():
        t = v.clone()
        if t.is_floating_point() and requires_grad:
            t.requires_grad_(True)
        sd[k] = t
    return cfg, sd, orc.OracleHSTU(cfg, sd, fx["category_counts"], fx["category_to_int"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_train_loss_grads_logs(name):
    fx = load_golden(name)
    cfg, sd, model = _oracle(fx, requires_grad=True)
    out = model.forward(fx["train_batch"])
    assert abs(float(out["loss"]) - fx["loss"]) <= 2e-6 * max(1.0, abs(fx["loss"]))
    out["loss"].backward()
    for k, g in fx["grads"].items():
        if g is None:  # rel-bias params: never used by the reference
            assert sd[k].grad is None
            continue
        assert torch.allclose(sd[k].grad, g, rtol=1e-4, atol=1e-7), k
    for k, v in fx["logs"].items():
        if k != "loss":
            assert abs(float(out[k]) - v) <= 1e-5, k


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_eval_scores_topk_metrics(name):
    fx = load_golden(name)
    cfg, sd, model = _oracle(fx)
    ev = fx["eval_batch"]
    C = cfg["eval_num_cats"]
    if cfg["category_by"] == "item":
        all_item_tags = fx["item_tags"].t().contiguous()
    else:
        all_item_tags = torch.ones(C, cfg["item_num"], dtype=torch.bool)
    feat = model.compute_item_all()
    assert torch.allclose(feat, fx["item_feature"], rtol=1e-6, atol=1e-7)
    scores, logs, _, _ = model.predict(ev["item_seq"], None, feat, all_item_tags, ev["target_tags"])
    scores = orc.post_mask_scores(scores, ev["history_index"])
    ref = fx["scores"]
    assert torch.equal(torch.isinf(scores), torch.isinf(ref))
    fin = torch.isfinite(ref)
    assert torch.allclose(scores[fin], ref[fin], rtol=1e-5, atol=1e-6)
    assert logs["num_samples"] == fx["num_samples"]
    # top-K / hit matrix / metrics from the REFERENCE scores (removes fp noise from the id comparison)
    K = max(cfg["topk"])
    idx, vals, src = orc.collect_topk(ref, K, cfg["split_mode"])
    hits = orc.hit_matrices(idx, ev["item_target"].numpy(), cfg["metrics_pred_len_list"])
    for p in cfg["metrics_pred_len_list"]:
        assert np.array_equal(hits[p], fx["rec_topk"][p].numpy()), p
        m = orc.recall_ndcg_sums(hits[p], cfg["topk"])
        for k, v in fx["metrics"][p].items():
            assert abs(m[k] - v) <= 1e-9 * max(1.0, abs(v)), (p, k)


def test_combine_equals_max_over_heads():
    """SURVEY A.5 identity used by the fused kernel: the reference's per-head top-K + sort +
    dedupe merge equals top-K of the max over heads (tie-free scores)."""
    g = torch.Generator().manual_seed(0)
    s = torch.randn(4, 5, 400, generator=g)
    s[:, :, 0] = float("-inf")
    s[:, 2, 100:300] = float("-inf")
    idx, vals, src = orc.collect_topk(s, 50, "combine")
    mx, am = s.max(dim=1)
    v2, i2 = torch.topk(mx, 50, dim=-1)
    assert np.array_equal(idx, i2.numpy())
    assert np.array_equal(src, torch.gather(am, 1, i2).numpy())


@pytest.mark.parametrize("name", ["comirec_p1", "comirec_p4", "remi_p1", "remi_p3_beta4", "remi_p2_beta0"])
def test_comirec_oracle_matches_live_reference_fixture(name):
    """oracle/comirec_oracle.py against tests/golden/comirec_*.pt / remi_*.pt (live reference ComiRec / REMI classes,
    make_golden_comirec.py)."""
    from oracle.comirec_oracle import OracleComiRec, OracleREMI
    fx = load_golden(name)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["state_dict"].items()}
    model = (OracleREMI if name.startswith("remi") else OracleComiRec)(fx["cfg"], sd)
    out = model.forward(fx["train_batch"])
    assert abs(float(out["loss"].detach()) - fx["loss"]) <= 2e-6 * max(1.0, abs(fx["loss"]))
    out["loss"].backward()
    for k, g in fx["grads"].items():
        if g is None:
            assert sd[k].grad is None, k
            continue
        err = (sd[k].grad - g).abs().max().item() / max(g.abs().max().item(), 1e-12)
        assert err < 2e-3, (k, err)
    for k, v in fx["logs"].items():
        if k != "loss":
            assert abs(float(out[k]) - v) <= 1e-5, k
    scores = model.predict(fx["eval_batch"]["item_seq"], fx["item_feature"])
    assert torch.allclose(scores, fx["scores"], rtol=1e-5, atol=1e-6)
