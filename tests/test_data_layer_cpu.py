"""Host-side logic either side of the hot path, pinned against tests/golden/data_layer.pt — outputs of the LIVE
reference data layer / schedules written by tests/golden/make_golden_data.py:
  * prior dictionaries -> heads (dataload.py:347-371, 226-246, 312-327)   [SURVEY x3]
  * training-window locations (dataload.py:165-194)                        [N2]
  * LR schedules (utils/lr_scheduler.py:44-116)                            [N1]
Where /root/reference is mounted the same checks also run against the live modules."""
import os

import pytest
import torch

from conftest import ROOT

from b200rec import priors
from b200rec import trainer as T

REF_DATA = "/root/reference/code/REC/data"


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(ROOT, "tests", "golden", "data_layer.pt"), weights_only=False)


def test_prior_spec_from_dict_matches_reference_build(fx):
    it = fx["item_features"]
    spec = priors.PriorSpec.from_source({"v1": dict(tag_to_category=it["tag_to_category"],
                                                    category_counts=it["category_counts"])}, "item", "v1")
    want = fx["prior_dicts"][("Pixel8M_tag_dict", "v1")]
    assert spec.category_to_int == want["category_to_int"]
    assert spec.int_to_category == want["int_to_category"]
    assert spec.category_counts == want["category_counts"]
    cfg = spec.apply_to_config({})
    assert cfg["int_to_category"] == want["int_to_category"] and cfg["eval_num_cats"] == 8
    w = spec.prior_loss_weight()
    tot = sum(want["category_counts"].values())
    assert w == [want["category_counts"][want["int_to_category"][i]] / tot for i in range(8)]   # hstu.py:503-510


def test_item_tag_table_matches_reference_item_features(fx):
    it = fx["item_features"]
    spec = priors.PriorSpec.from_source({"v1": dict(tag_to_category=it["tag_to_category"],
                                                    category_counts=it["category_counts"])}, "item", "v1")
    tags, pools = priors.item_tag_table(it["raw_tags"], spec)
    assert torch.equal(tags, it["tag_category"])                       # item_to_info[i]['tag_category']
    assert not tags[0].any()
    # the reference's per-category pools come out of a pandas groupby in parquet row order: same SETS of items
    for c in range(spec.num_categories):
        assert sorted(pools[c].tolist()) == sorted(it["int_category_to_item_id"][c])
    unmapped = [i for i, t in enumerate(it["raw_tags"]) if t == "tag-with-no-mapping"]
    assert unmapped and not tags[unmapped].any()


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the shipped prior dictionaries live in /root/reference")
def test_every_shipped_prior_dictionary_loads_like_the_reference(fx):
    for (mod, ver), want in fx["prior_dicts"].items():
        spec = priors.PriorSpec.from_source(os.path.join(REF_DATA, mod + ".py"), want["category_by"], ver)
        assert spec.category_to_int == want["category_to_int"], (mod, ver)
        assert spec.int_to_category == want["int_to_category"], (mod, ver)
        assert spec.category_counts == want["category_counts"], (mod, ver)
        assert len(spec.tag_to_category or {}) == want["n_tags"]
    with pytest.raises(KeyError):                                       # the EB-NeRD scripts ask for v3 / v16: not shipped
        priors.PriorSpec.from_source(os.path.join(REF_DATA, "eb_nerd_512_tag_dict.py"), "item", "v3")


def test_prior_spec_builds_the_model_heads(fx):
    """The adapter's output is what HSTU(config, dataload) consumes: head count, head order, loss weights."""
    from b200rec import synth
    from b200rec.hstu import HSTU
    want = fx["prior_dicts"][("merrec_2000_tag_dict", None)]
    spec = priors.PriorSpec(want["category_counts"], want["category_to_int"])
    cfg = synth.make_config("C", n_layers=1, n_heads=1, item_embedding_size=16, hstu_embedding_size=16,
                            MAX_ITEM_LIST_LENGTH=8, item_num=50, num_prior_head=spec.num_categories)
    spec.apply_to_config(cfg)
    model = HSTU(cfg, spec.dataload(50))
    assert model.medusa_num_heads == 6 and cfg["int_to_category"][5] == "buy_comp"
    assert abs(sum(model.prior_loss_weight) - 1.0) < 1e-12
    assert model.prior_loss_weight == spec.prior_loss_weight()


def test_training_windows_match_reference(fx):
    from b200rec.batcher import InteractionData
    d = fx["data_layer"]
    data = InteractionData(d["user_seq"], d["train_seq_len"], d["N"], d["L"], item_tags=d["item_tags"], device="cpu",
                           pred_len=d["P"], include_empty_context=True)
    got = list(zip(data.h_sample_uid.tolist(), data.h_sample_end.tolist()))
    assert got == [tuple(x) for x in d["valid_sample_locations"]]
    # default: windows without a single context position are dropped (they carry no loss term)
    data2 = InteractionData(d["user_seq"], d["train_seq_len"], d["N"], d["L"], item_tags=d["item_tags"], device="cpu",
                            pred_len=d["P"])
    kept = [tuple(x) for x in d["valid_sample_locations"] if x[1] > 0]
    assert list(zip(data2.h_sample_uid.tolist(), data2.h_sample_end.tolist())) == kept and len(kept) < len(got)


def test_schedules_match_reference_lambda_lr(fx):
    fns = {"cosine": T.cosine_schedule_with_warmup, "linear": T.linear_schedule_with_warmup}
    for (name, warm, total), want in fx["schedules"].items():
        got = [3e-3 * fns[name](s, warm, total) for s in range(len(want))]
        assert max(abs(a - b) for a, b in zip(want, got)) < 1e-15, (name, warm, total)


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="needs /root/reference")
def test_schedules_match_live_reference_module():
    import sys
    from oracle import ref_harness as rh
    rh.load()
    from REC.utils import lr_scheduler
    for name, ref_fn, mine in (("cosine", lr_scheduler.get_cosine_schedule_with_warmup, T.cosine_schedule_with_warmup),
                               ("linear", lr_scheduler.get_linear_schedule_with_warmup, T.linear_schedule_with_warmup)):
        for warm, total in ((7.5, 60), (30.0, 300)):                     # trainer.py:457-458: warmup = total * 0.1 etc.
            p = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.SGD([p], lr=1e-2)
            sch = ref_fn(opt, num_warmup_steps=warm, num_training_steps=total)
            for s in range(total + 3):
                assert abs(opt.param_groups[0]["lr"] - 1e-2 * mine(s, warm, total)) < 1e-15, (name, s)
                opt.step()
                sch.step()
