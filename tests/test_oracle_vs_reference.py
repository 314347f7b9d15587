"""Runs the restated oracle against the UNMODIFIED reference modules.  Only possible in the
build container (needs /root/reference); skipped elsewhere — the committed golden fixtures
(test_oracle_golden.py) carry the same pin to the GPU box."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle import hstu_oracle as orc
from b200rec import synth

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not mounted")

SMALL = dict(n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
             MAX_ITEM_LIST_LENGTH=10, train_batch_size=5, num_negatives=20, item_num=200)


@pytest.mark.parametrize("preset,over", [
    ("A", dict(train_batch_size=16, num_negatives=64, item_num=500)),
    ("B", SMALL), ("D", SMALL), ("C", SMALL),
    ("D", dict(SMALL, num_segment_head=2, pred_len=4, eval_pred_len=4)),
    ("B", dict(SMALL, item_embedding_size=24)),                                     # item_id_proj_tower
    ("D", dict(SMALL, medusa_num_layers=2)),                                        # weight-tied 2-layer heads
    ("D", dict(SMALL, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
               medusa_num_layers=2)),                                              # seg[c][s](cat[c](x))
    ("D", dict(SMALL, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
               medusa_num_layers=2, head_norm=True, cat_bottleneck=True, share_seg_weights=True, segment_embed=True)),
    ("D", dict(SMALL, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
               medusa_num_layers=1, head_norm=True, cat_bottleneck=True, cat_bottleneck_dim=8)),
    ("D", dict(SMALL, head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
               medusa_num_layers=1, share_seg_weights=True, segment_embed=True)),
], ids=["nce", "additive", "mult", "event", "mult-2seg", "tower", "tied-2layer", "hierarchical", "hier-all-options",
        "hier-norm-bottleneck", "hier-shared-segembed"])
def test_train_step_matches_reference(preset, over):
    cfg = synth.make_config(preset, **over)
    dl = synth.make_dataload(cfg)
    ref = rh.build_reference_model(dict(cfg), cfg["item_num"], dl.category_counts, dl.category_to_int)
    ref.eval()
    batch = synth.make_train_batch(cfg, seed=11, zipf=False)
    out = ref(batch)
    out["loss"].backward()
    sd = orc.state_dict_from_module(ref, requires_grad=True)
    o2 = orc.OracleHSTU(cfg, sd, dl.category_counts, dl.category_to_int).forward(batch)
    o2["loss"].backward()
    assert abs(float(out["loss"].detach()) - float(o2["loss"].detach())) < 2e-6 * max(1, abs(float(out["loss"].detach())))
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert sd[k].grad is None, k
        else:
            assert torch.allclose(p.grad, sd[k].grad, rtol=1e-4, atol=1e-7), k
    for k in out:
        if k != "loss":
            assert abs(float(out[k]) - float(o2[k])) < 1e-5, k


def test_relative_position_bias_matches_live_module():
    """the bias the reference BUILDS (and never applies, hstu.py:221-290) — the oracle's restatement of its
    position part equals RelativeBucketedTimeAndPositionBasedBias.forward (hstu.py:99-134) for equal timestamps."""
    import torch
    from oracle import ref_harness as rh, hstu_oracle as orc
    if not rh.available():
        import pytest
        pytest.skip("reference tree not mounted")
    rh.load()
    from REC.model.IDNet.hstu import RelativeBucketedTimeAndPositionBasedBias as RB
    for N, L in ((24, 12), (100, 50)):
        m = RB(max_seq_len=N, num_buckets=128,
               bucketization_fn=lambda x: (torch.log(torch.abs(x).clamp(min=1)) / 0.301).long())
        ref = m(torch.zeros(2, N, dtype=torch.long))[0, :L, :L]
        assert torch.equal(ref, orc.rel_pos_bias(m._pos_w.data, m._ts_w.data, L))


@pytest.mark.parametrize("over", [
    dict(),
    dict(head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4, medusa_num_layers=2),
    dict(head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4, medusa_num_layers=2,
         head_norm=True, cat_bottleneck=True, share_seg_weights=True, segment_embed=True),
], ids=["mult", "hierarchical", "hier-all-options"])
def test_same_seed_init_is_bit_identical_to_the_reference(over):
    """parameters are created in the reference's order with the reference's initialisers, so the same seed gives
    the same weights (incl. the hierarchical-head options: segment_emb, bottleneck, per-block LayerNorm, shared block)."""
    from b200rec.hstu import HSTU
    cfg = synth.make_config("D", **dict(SMALL, **over))
    dl = synth.make_dataload(cfg)
    ref = rh.build_reference_model(dict(cfg), cfg["item_num"], dl.category_counts, dl.category_to_int, seed=2020)
    torch.manual_seed(2020)
    mine = HSTU(cfg, dl, compute_dtype=torch.float32)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd) == list(msd)
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k
