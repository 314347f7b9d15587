"""Parity at the REAL widths / depths / head layouts / negative counts of BASELINE.json's configs:

  A  HSTU-Pixel8M-base small      full size (B=64, L=20, D=64, 2 blocks, 512 negatives, 10 k items)
  B  HSTU-Pixel8M-prior           D=1024, 16 blocks, 16 attention heads, 4+8 additive heads, P=8, 8192 negatives x 9 sets
  C  HSTU-MerRec-prior            D=1024, 16 blocks, L=400, 6 event heads, fixed temperature, 4096 negatives
  D  HSTU-EBNerd-prior-mult       D=256, 8 blocks, dh=32, 7 multiplicative heads, P=8, 8192 negatives x 7 sets

B/C/D run at batch 4-8 (negatives per SET as in the scripts) and a 100 k-item catalogue so the dense CPU oracle finishes
in seconds; every kernel sees the production tile shapes (K = 1024 / 256 contractions, 16-deep residual stream, 8192-wide
logit rows, L = 400 attention).  Checked against `oracle.OracleHSTU` (fp32, CPU): loss, every parameter gradient, every
logging scalar, eval top-K ids.  fp32 verification mode: tight.  bf16 production mode: asserted at the measured level and
the per-tensor cosines are written to gpurun_out/bf16_parity_<config>.json (copied to profiles/ per round)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

from b200rec import synth  # noqa: E402
from b200rec.hstu import HSTU  # noqa: E402

DEV = torch.device("cuda:0")

CASES = {
    "A": ("A", dict()),
    "B": ("B", dict(train_batch_size=4, num_negatives=8192, item_num=100000, hidden_dropout_prob=0.0)),
    "C": ("C", dict(train_batch_size=4, num_negatives=4096, item_num=100000, hidden_dropout_prob=0.0)),
    "D": ("D", dict(train_batch_size=8, num_negatives=8192, item_num=100000, hidden_dropout_prob=0.0)),
}
# bf16 production mode, asserted just below what the kernels achieve (measured values and the reference's own bf16-mixed
# numbers live in profiles/r02_bf16_parity.md): minimum per-tensor gradient cosine, relative loss error, top-200 overlap.
# Measured r02: A 0.99996 / 1.1e-4 / 0.996, B 0.9818 / 4.5e-5 / 0.889, C 0.9871 / 6.6e-5 / 0.964, D 0.99885 / 2.3e-5 / 0.967.
# Named cause of B / C < 0.999 (SURVEY D.4's proposal): the body output of the 16-block stacks differs by 8 % between
# fp32 and bf16 activations (scripts/bf16_error_trace.py: the random-init residual stream grows x60 over 16 blocks and
# every block re-rounds n, uvqk, P and the gate input to bf16), the NCE gradient at the top is within 0.3 %.  The
# UNMODIFIED reference under its own production precision (torch.autocast bf16, `bf16-mixed`) loses MORE against its
# fp32 self on the same batch: min cosine 0.9703 at config B (scripts/reference_bf16_autocast_error.py).
# Top-K overlap is bounded by the random-init catalogue: 100 k untrained items put neighbouring top-200 scores ~1e-4
# apart, below the 2^-9 relative rounding of bf16 operands.
BF16_BAR = {"A": (0.9995, 1e-3, 0.98), "B": (0.975, 1e-3, 0.85), "C": (0.98, 1e-3, 0.93), "D": (0.998, 1e-3, 0.94)}

_cache = {}


def _setup(name):
    if name in _cache:
        return _cache[name]
    from oracle import hstu_oracle as orc
    preset, over = CASES[name]
    cfg = synth.make_config(preset, **over)
    dl = synth.make_dataload(cfg)
    torch.manual_seed(2020)
    host = HSTU(cfg, dl, compute_dtype=torch.float32)
    params = dict(host.named_parameters())             # `logit_scale` is a buffer (no gradient) under fix_temp
    sd = {k: v.detach().clone().requires_grad_(k in params and "_rel_attn_bias" not in k)
          for k, v in host.state_dict().items()}
    item_tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
    batch = synth.make_train_batch(cfg, seed=11, item_tags=item_tags)
    torch.set_num_threads(os.cpu_count() or 1)
    oracle = orc.OracleHSTU(cfg, sd, dl.category_counts, dl.category_to_int)
    ref = oracle.forward(batch)
    ref["loss"].backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    logs = {k: float(v) for k, v in ref.items()}
    # eval: 8 users against the full catalogue
    ev = synth.make_eval_batch(cfg, seed=6, batch_size=8, item_tags=item_tags)
    C = cfg["eval_num_cats"]
    tags_cn = item_tags.t().contiguous() if cfg["category_by"] == "item" else \
        torch.ones(C, cfg["item_num"], dtype=torch.bool)
    with torch.no_grad():
        o2 = orc.OracleHSTU(cfg, {k: v.detach() for k, v in sd.items()}, dl.category_counts, dl.category_to_int)
        scores, _, _, _ = o2.predict(ev["item_seq"], None, o2.compute_item_all(), tags_cn, ev["target_tags"])
        scores = orc.post_mask_scores(scores, ev["history_index"])
        K = 200
        ref_idx, ref_val, _ = orc.collect_topk(scores, K, "combine")
    out = dict(cfg=cfg, dl=dl, state=host.state_dict(), batch=batch, logs=logs, grads=grads, ev=ev, tags_cn=tags_cn,
               ref_idx=np.asarray(ref_idx), ref_val=np.asarray(ref_val), K=K)
    _cache.clear()                      # one config resident at a time (B holds ~1.5 GB of host tensors)
    _cache[name] = out
    return out


def _run_gpu(s, dtype):
    cfg = synth.Config(s["cfg"])
    cfg["sparse_embedding_grad"] = True
    m = HSTU(cfg, s["dl"], compute_dtype=dtype)
    m.load_state_dict(s["state"])
    m = m.to(DEV).eval()
    out = m(tuple(t.to(DEV) for t in s["batch"]))
    out["loss"].backward()
    grads = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    uid, urows, nu = m.emb_grad
    k = int(nu.item())
    emb = torch.zeros((cfg["item_num"], cfg["item_embedding_size"]))
    emb[uid[:k].cpu()] = urows[:k].cpu()
    grads["item_embedding.weight"] = emb
    logs = {kk: float(v) for kk, v in out.items()}
    ev = s["ev"]
    hu, hi = ev["history_index"]
    with torch.no_grad():
        idx, val, _ = m.predict_topk(ev["item_seq"].to(DEV), m.compute_item_all(), s["tags_cn"].to(DEV),
                                     ev["target_tags"].to(DEV), history_index=(hu.to(DEV), hi.to(DEV)), K=s["K"])
    return logs, grads, idx.cpu().numpy(), val.cpu().numpy()


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


@pytest.fixture(scope="module", params=["A", "B", "C", "D"])
def case(request):
    """Module-scoped so pytest runs the fp32 and bf16 checks of one config back to back (one oracle pass each)."""
    return request.param, _setup(request.param)


def test_fp32_mode_matches_oracle_at_real_shapes(case):
    name, s = case
    logs, grads, idx, val = _run_gpu(s, torch.float32)
    ref = s["logs"]
    assert abs(logs["loss"] - ref["loss"]) <= 2e-5 * max(1.0, abs(ref["loss"])), (logs["loss"], ref["loss"])
    for k, v in ref.items():
        assert abs(logs[k] - v) <= 2e-4 * max(1.0, abs(v)), (k, logs[k], v)
    for k, g_ref in s["grads"].items():
        g = grads[k]
        # 16-block fp32 chains accumulate in a different order than the CPU oracle: relative-to-max tolerance
        err = (g - g_ref).abs().max().item() / max(1e-12, g_ref.abs().max().item())
        assert err < 1e-3, (k, err)
        if g_ref.numel() > 1:
            assert _cos(g, g_ref) > 0.99999, (k, _cos(g, g_ref))
    # exact gradient row set of the item table
    got_rows = set(torch.nonzero(grads["item_embedding.weight"].abs().sum(1) > 0).flatten().tolist())
    want_rows = set(torch.nonzero(s["grads"]["item_embedding.weight"].abs().sum(1) > 0).flatten().tolist())
    assert got_rows == want_rows
    # top-K ids: identical wherever the oracle's neighbouring scores are separated by more than fp32 noise
    ref_idx, ref_val = s["ref_idx"], s["ref_val"]
    for b in range(ref_idx.shape[0]):
        gap = np.abs(np.diff(ref_val[b]))
        tie_near = np.zeros(ref_idx.shape[1], dtype=bool)
        tie_near[:-1] |= gap < 3e-6
        tie_near[1:] |= gap < 3e-6
        fin = np.isfinite(ref_val[b])
        ok = (idx[b] == ref_idx[b]) | tie_near | ~fin
        assert ok.all(), (b, np.nonzero(~ok)[0][:5])
        assert np.allclose(val[b][fin], ref_val[b][fin], atol=2e-5)


def test_bf16_mode_within_stated_tolerance_at_real_shapes(case):
    name, s = case
    logs, grads, idx, val = _run_gpu(s, torch.bfloat16)
    ref = s["logs"]
    cos_bar, loss_bar, ov_bar = BF16_BAR[name]
    report = {"config": name, "loss": logs["loss"], "loss_ref": ref["loss"],
              "loss_rel_err": abs(logs["loss"] - ref["loss"]) / abs(ref["loss"]), "cosine": {}, "rel_max_err": {}}
    for k, g_ref in s["grads"].items():
        if g_ref.numel() < 2:
            continue
        report["cosine"][k] = _cos(grads[k], g_ref)
        report["rel_max_err"][k] = (grads[k] - g_ref).abs().max().item() / max(1e-12, g_ref.abs().max().item())
    ov = [len(set(a.tolist()) & set(b.tolist())) / float(len(b)) for a, b in zip(idx, s["ref_idx"])]
    report["topk_overlap_mean"], report["topk_overlap_min"] = float(np.mean(ov)), float(np.min(ov))
    report["top10_overlap_mean"] = float(np.mean([len(set(a[:10].tolist()) & set(b[:10].tolist())) / 10.0
                                                  for a, b in zip(idx, s["ref_idx"])]))
    worst = min(report["cosine"].items(), key=lambda kv: kv[1])
    report["cosine_min"] = {"tensor": worst[0], "value": worst[1]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bf16_parity_{name}.json"), "w") as f:
        json.dump(report, f, indent=1)
    assert report["loss_rel_err"] <= loss_bar, report["loss_rel_err"]
    assert worst[1] >= cos_bar, worst
    assert report["topk_overlap_mean"] >= ov_bar, report["topk_overlap_mean"]
    for k in ("nce_top1_acc", "nce_top10_acc", "nce_samples"):
        if k in ref:
            assert abs(logs[k] - ref[k]) <= 2e-2 * max(1.0, abs(ref[k])), (k, logs[k], ref[k])


def test_bf16_recall_ndcg_parity_on_10k_users():
    """BASELINE metric: "Recall@10 parity".  10 240 synthetic users, 20 k items, config-B head layout (4+8 additive
    heads with item-category masks); every user's target is drawn from the fp32 oracle's own top-30 so that
    Recall@10 ~ 1/3 and rank swaps around the cut-off move the metric.  |delta| <= 1e-3 on Recall@10 / NDCG@10
     between bf16 CUDA predict_topk + Collector / Evaluator and the fp32 CPU oracle pipeline."""
    from oracle import hstu_oracle as orc
    from b200rec.evaluator import Collector, Evaluator
    cfg = synth.make_config("B", n_layers=4, n_heads=4, item_embedding_size=256, hstu_embedding_size=256,
                            item_num=20000, eval_pred_len=1, pred_len=8, hidden_dropout_prob=0.0, topk=[10, 50])
    cfg["metrics_pred_len_list"] = [0]
    dl = synth.make_dataload(cfg)
    torch.manual_seed(2020)
    host = HSTU(cfg, dl, compute_dtype=torch.float32)
    sd = {k: v.detach() for k, v in host.state_dict().items()}
    item_tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
    tags_cn = item_tags.t().contiguous()
    U, Bs = 10240, 512
    torch.set_num_threads(os.cpu_count() or 1)
    oracle = orc.OracleHSTU(cfg, sd, dl.category_counts, dl.category_to_int)
    m = HSTU(cfg, dl, compute_dtype=torch.bfloat16)
    m.load_state_dict(host.state_dict())
    m = m.to(DEV).eval()
    feat_gpu = m.compute_item_all()
    c_ref, c_gpu = Collector(cfg), Collector(cfg)
    gen = torch.Generator().manual_seed(77)
    with torch.no_grad():
        feat = oracle.compute_item_all()
        for b0 in range(0, U, Bs):
            ev = synth.make_eval_batch(cfg, seed=100 + b0, batch_size=Bs, item_tags=item_tags)
            scores, _, _, _ = oracle.predict(ev["item_seq"], None, feat, tags_cn, ev["target_tags"])
            scores = orc.post_mask_scores(scores, ev["history_index"])
            # collector merge == top-K of the max over heads (proven on the fixtures in
            # test_gpu_model.py::test_eval_matches_reference); fp32 scores of distinct items do not tie
            fold, rhead = scores.max(dim=1)
            rval, ridx = torch.topk(fold, 50, dim=1)
            rhead = rhead.gather(1, ridx)
            pick = torch.randint(0, 30, (Bs,), generator=gen)
            target = ridx[torch.arange(Bs), pick].unsqueeze(1)                     # [Bs, 1]
            pos_u = torch.arange(Bs).unsqueeze(1)
            top_ref = (ridx.to(DEV), rval.to(DEV), rhead.to(DEV))
            c_ref.eval_batch_collect(None, pos_u, target.to(DEV), None, topk=top_ref)
            hu, hi = ev["history_index"]
            top = m.predict_topk(ev["item_seq"].to(DEV), feat_gpu, tags_cn.to(DEV), ev["target_tags"].to(DEV),
                                 history_index=(hu.to(DEV), hi.to(DEV)), K=50)
            c_gpu.eval_batch_collect(None, pos_u, target.to(DEV), None, topk=top)
    e = Evaluator(cfg)
    r_ref, r_gpu = e.evaluate(c_ref.get_data_struct(0), 0), e.evaluate(c_gpu.get_data_struct(0), 0)
    report = {}
    for k in ("recall@10", "ndcg@10", "recall@50", "ndcg@50"):
        a = (r_ref[k][0] if isinstance(r_ref[k], tuple) else r_ref[k]) / U
        b = (r_gpu[k][0] if isinstance(r_gpu[k], tuple) else r_gpu[k]) / U
        report[k] = {"oracle_fp32": a, "cuda_bf16": b, "delta": b - a}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bf16_recall_parity.json"), "w") as f:
        json.dump(report, f, indent=1)
    assert 0.2 < report["recall@10"]["oracle_fp32"] < 0.5
    for k in ("recall@10", "ndcg@10"):
        assert abs(report[k]["delta"]) <= 1e-3, report
