"""Generates tests/golden/*.pt from the LIVE, UNMODIFIED reference (run in the build
container only: needs /root/reference).  Usage:  python tests/golden/make_golden.py

Each fixture holds, for one tiny config: the config, the reference-initialised state dict,
one seeded train batch with the reference's loss / logging scalars / every parameter
gradient, and one eval batch with the reference's predict() scores pushed through
trainer.py:724-726 masks, Collector.eval_batch_collect and Evaluator.evaluate.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from b200rec import synth  # noqa: E402

TINY = dict(n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
            MAX_ITEM_LIST_LENGTH=12, train_batch_size=6, num_negatives=30, item_num=300,
            eval_batch_size=5)
CASES = {
    "nce_single": ("A", dict(TINY, n_heads=1)),
    "nce_2attn": ("A2", dict(TINY)),
    "prior_additive": ("B", dict(TINY)),
    "prior_mult": ("D", dict(TINY)),
    "prior_event_given": ("C", dict(TINY, MAX_ITEM_LIST_LENGTH=16)),
    "nce_pred4": ("A2", dict(TINY, pred_len=4, eval_pred_len=4, medusa_num_layers=1, num_segment_head=2)),
    # item_id_proj_tower (table width 24 -> HSTU width 32) under the additive prior heads
    "tower_additive": ("B", dict(TINY, item_embedding_size=24)),
    # weight-tied two-layer decode heads (`[ResBlock] * 2`, hstu.py:486-493)
    "mult_2layers": ("D", dict(TINY, medusa_num_layers=2)),
    # hierarchical heads seg[c][s](cat[c](x)) (hstu.py:443-483, 652-663), 2 segments x 7 categories, 2 layers each
    "hier_2x7": ("D", dict(TINY, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
                           medusa_num_layers=2)),
    # every hierarchical-head option at once (hstu.py:444-483): LayerNorm inside the ResBlocks, bottleneck MLP (32 -> 16 -> 32)
    # in front of the category block, ONE segment block shared by all (category, segment), learned segment offsets
    "hier_options": ("D", dict(TINY, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
                               medusa_num_layers=2, head_norm=True, cat_bottleneck=True, share_seg_weights=True,
                               segment_embed=True)),
    # prior-switch aux heads (hstu.py:512-544, 731-805): weighted BCE over all positions / ASL on the last position with a
    # master switch, both used at test time to switch prior heads off
    "switch_bce": ("B", dict(TINY, prior_switch="in", prior_switch_loss_weight=0.7, use_prior_switch_test=True)),
    "switch_asl_master": ("D", dict(TINY, prior_switch="in", prior_switch_loss_weight=1.3, asym_switch_loss=True,
                                    switch_last_only=True, master_switch=True, use_prior_switch_test=True)),
}
TOPK = [1, 5, 10, 20]


def run_case(name, preset, over):
    cfg = synth.make_config(preset, **over)
    cfg["topk"] = TOPK
    dl = synth.make_dataload(cfg)
    ref = rh.build_reference_model(dict(cfg), cfg["item_num"], dl.category_counts, dl.category_to_int)
    ref.eval()
    item_tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
    batch = synth.make_train_batch(cfg, seed=3, item_tags=item_tags, zipf=False)
    out = ref(batch)
    out["loss"].backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in ref.named_parameters()}
    logs = {k: float(v) for k, v in out.items()}
    # ---- eval through the reference trainer/collector/evaluator sequence
    _, Collector, Evaluator = rh.load()
    ev = synth.make_eval_batch(cfg, seed=5, item_tags=item_tags)
    C = cfg["eval_num_cats"]
    if cfg["category_by"] == "item":
        all_item_tags = item_tags.t().contiguous().to(torch.int64)   # [C, N] (trainer.py:824)
        all_tags_NC = item_tags.to(torch.int64)
    else:
        all_item_tags = torch.ones(C, cfg["item_num"], dtype=torch.int64)  # batchset.py:36-38
        all_tags_NC = torch.ones(cfg["item_num"], C, dtype=torch.int64)
    with torch.no_grad():
        feat = ref.compute_item_all()
        scores, plogs, _, _ = ref.predict(ev["item_seq"], None, feat, all_item_tags, ev["target_tags"])
        scores[:, :, 0] = float("-inf")                               # trainer.py:724
        hu, hi = ev["history_index"]
        scores[hu, :, hi] = float("-inf")                             # trainer.py:725-726
    ccfg = rh.RefConfig(dict(cfg))
    ccfg["device"] = "cpu"
    collector = Collector(ccfg)
    collector.set_all_tags(all_tags_NC)
    final_scores = scores.clone()
    collector.eval_batch_collect(scores, ev["positive_u"], ev["item_target"], ev["target_tags"], None, False)
    evaluator = Evaluator(ccfg)
    rec_topk, metrics = {}, {}
    for p in cfg["metrics_pred_len_list"]:
        struct = collector.get_data_struct(p)
        rec_topk[p] = struct.get("rec.topk").clone()
        res = evaluator.evaluate(struct, p)
        metrics[p] = {k: (float(v[0]) if isinstance(v, tuple) else float(v)) for k, v in res.items() if "-" not in k}
    fx = dict(
        name=name, cfg=dict(cfg), category_counts=dl.category_counts, category_to_int=dl.category_to_int,
        state_dict={k: v.clone() for k, v in ref.state_dict().items()},
        item_tags=item_tags, train_batch=batch, loss=float(out["loss"]), logs=logs, grads=grads,
        eval_batch=ev, item_feature=feat, scores=final_scores, rec_topk=rec_topk, metrics=metrics,
        num_samples=plogs["num_samples"],
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.pt")
    torch.save(fx, path)
    print(f"{name}: loss={fx['loss']:.6f} heads={scores.shape[1]} -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (preset, over) in CASES.items():
        if not only or name in only:
            run_case(name, preset, over)
