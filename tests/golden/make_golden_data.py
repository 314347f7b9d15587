"""Generates tests/golden/data_layer.pt from the LIVE, UNMODIFIED reference data layer and schedules
(run in the build container only: needs /root/reference).  Usage:  python tests/golden/make_golden_data.py

Pins the "next" rows of SURVEY §8(f) that sit either side of the hot path:
  * N2 batch construction: REC.data.dataset.trainset.SEQTrainDataset.__getitem__ (trainset.py:99-177),
    REC.data.dataset.evalset.SeqEvalDataset.__getitem__ + collate_fn.seq_eval_collate (evalset.py:81-155,
    collate_fn.py:59-90), REC.data.dataload.InteractionData._get_valid_sample_loc_for_train (dataload.py:165-194)
    driven with a fake `dataload` object holding synthetic interaction lists;
  * x3 prior dictionaries: InteractionData.build (dataload.py:347-371) on every shipped `*_dict.py`, and
    InteractionData._load_item_feat (dataload.py:196-331) on a synthetic item parquet with real Pixel8M tags;
  * N1 LR schedules: REC.utils.lr_scheduler.get_{cosine,linear}_schedule_with_warmup through LambdaLR.
Only the modules that need absent third-party packages at import time are stubbed (polars, torch_geometric,
colorlog, colorama, tensorboardX, pytz); every function that produces fixture content is the reference's own.
"""
import importlib
import logging
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

DICTS = [("Pixel8M_tag_dict", "item", ["v1", "v2"]), ("Pixel8M_cluster_dict", "item", ["v0"]),
         ("eb_nerd_512_tag_dict", "item", ["v1", "v2"]), ("eb_nerd_512_cluster_dict", "item", ["v1"]),
         ("eb_nerd_512_user_cluster_dict", "user", ["v1"]), ("merrec_2000_tag_dict", "event", [None])]


def load_reference_modules():
    rh.load()                                        # stubs colorlog / colorama / tensorboardX / pytz, sys.path, gloo group
    for n, attrs in (("polars", {}), ("torch_geometric", {}), ("torch_geometric.utils", {"degree": None})):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__dict__.update(attrs)
            sys.modules[n] = m
    import REC.data.dataload as dataload
    dataload.SharedList = lambda x: x                # shared-memory IPC is out of scope
    from REC.data.dataset.trainset import SEQTrainDataset
    from REC.data.dataset.evalset import SeqEvalDataset
    from REC.data.dataset.collate_fn import seq_eval_collate
    from REC.utils import lr_scheduler
    return dataload, SEQTrainDataset, SeqEvalDataset, seq_eval_collate, lr_scheduler


def ns(**kw):
    return types.SimpleNamespace(**kw)


def prior_dicts(dataload):
    """InteractionData.build (dataload.py:347-371) with local_rank != 0 so nothing is loaded from disk."""
    out = {}
    for mod_name, category_by, versions in DICTS:
        for ver in versions:
            dataset = mod_name.replace("_user_cluster_dict", "").replace("_cluster_dict", "").replace("_tag_dict", "")
            cfg = rh.RefConfig(eval_num_cats=2, tag_version=ver)
            fake = ns(config=cfg, cluster_as_tag=mod_name.endswith("cluster_dict"), category_by=category_by,
                      dataset_name=dataset, local_rank=1, logger=logging.getLogger("golden"), user_seq=[[]],
                      timestamp_required=False, valid_sample_locations=[], train_seq_len=[],
                      id2token={"user_id": [], "item_id": []}, item_interact_weights=[0], item_weights_by_cat=[[0]],
                      int_category_to_item_id=[], user_cluster_list=[0], event_seq=[], item_to_info=[{}],
                      uid_field="user_id", iid_field="item_id", category_counts={}, tag_to_category={},
                      category_to_int={})
            dataload.InteractionData.build(fake)
            out[(mod_name, ver)] = dict(category_by=category_by, category_counts=dict(fake.category_counts),
                                        category_to_int=dict(fake.category_to_int),
                                        int_to_category=dict(cfg["int_to_category"]),
                                        n_tags=len(fake.tag_to_category))
    return out


def item_features(dataload, n_items=300, seed=11):
    """InteractionData._load_item_feat (dataload.py:196-331) on a synthetic parquet carrying real Pixel8M tags."""
    ttg = importlib.import_module("REC.data.Pixel8M_tag_dict").tag_to_general["v1"]
    rng = np.random.default_rng(seed)
    tag_names = sorted(ttg["tag_to_category"].keys()) + ["tag-with-no-mapping"]
    tokens = ["[PAD]"] + [f"item{i:04d}" for i in range(1, n_items)]
    raw_tags = [None] + [tag_names[int(rng.integers(0, len(tag_names)))] for _ in range(1, n_items)]
    order = rng.permutation(np.arange(1, n_items))                 # the parquet is not in id order
    df = pd.DataFrame({"item_id": [tokens[i] for i in order], "title": ["t"] * (n_items - 1),
                       "tag": [raw_tags[i] for i in order]})
    counts = ttg["category_counts"]
    c2i = {cat: idx for idx, cat in enumerate(sorted(counts.keys()))}
    i2c = {v: k for k, v in c2i.items()}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "items.parquet")
        df.to_parquet(path)
        cfg = rh.RefConfig(text_path=path, text_keys=["title"], use_image=False, use_image_online=False,
                           neg_sample_mode=None, int_to_category=i2c, eval_num_cats=len(i2c))
        fake = ns(config=cfg, tag_col="tag", category_by="item", id2token={"item_id": tokens, "user_id": []},
                  use_image=False, use_image_online=False, image_dir=None, eval_num_cats=len(i2c), cluster_as_tag=False,
                  tag_to_category=ttg["tag_to_category"], logger=logging.getLogger("golden"), item_to_info=[{}],
                  item_num=n_items, item_interact_weights=[0], item_weights_by_cat=[[0]])
        dataload.InteractionData._load_item_feat(fake)
    tag_category = [[False] * len(i2c)] + [list(map(bool, info["tag_category"])) if info else [False] * len(i2c)
                                           for info in fake.item_to_info[1:]]
    return dict(raw_tags=raw_tags, tag_version="v1", int_to_category=i2c, tag_to_category=dict(ttg["tag_to_category"]),
                category_counts=dict(counts),
                tag_category=torch.tensor(tag_category, dtype=torch.bool),
                int_category_to_item_id=[list(map(int, lst)) for lst in fake.int_category_to_item_id])


def synthetic_interactions(n_users, N, C, seed, L):
    rng = np.random.default_rng(seed)
    user_seq, ev = [[]], [[]]
    for u in range(n_users):
        # a mix of short users (one window) and long users (several non-overlapping windows, some with an empty
        # first context: (n - 1) % (L + 1) == 0)
        n = int(rng.integers(5, 4 * L)) if u % 5 else (L + 1) * int(rng.integers(1, 3)) + 1 + 4
        user_seq.append([int(x) for x in rng.choice(np.arange(1, N), size=n, replace=False)])
        ev.append([int(x) for x in rng.integers(0, C, size=n)])
    tags = torch.rand(N, C, generator=torch.Generator().manual_seed(seed)) < 0.4
    tags[torch.arange(N), torch.randint(0, C, (N,), generator=torch.Generator().manual_seed(seed + 1))] = True
    tags[0] = False
    return user_seq, ev, tags


def data_layer(dataload, SEQTrainDataset, SeqEvalDataset, seq_eval_collate):
    L, P, Pe, C, N, n_users, n_neg_total, B = 10, 3, 2, 4, 400, 60, 16 * 12, 16
    user_seq, ev_seq, tags = synthetic_interactions(n_users, N, C, 5, L)
    base = dict(MAX_ITEM_LIST_LENGTH=L, pred_len=P, eval_pred_len=Pe, eval_num_cats=C, num_negatives=n_neg_total,
                train_batch_size=B, pad_random_sample=True, neg_sample_mix_ratio=0, neg_sample_mode=None,
                int_to_category={i: f"cat{i}" for i in range(C)}, device="cpu", timestamp_required=False,
                outlier_user_metrics=None)
    # valid_sample_locations / train_seq_len from the reference (dataload.py:165-194)
    loc = ns(user_num=len(user_seq), user_seq=user_seq, config=rh.RefConfig(eval_pred_len=Pe), train_test_gap=0,
             subset_user=False, subset_user_rmd=0, sample_last_only=False, pred_len=P, max_item_list_len=L + 1,
             id2token={"user_id": [f"u{i}" for i in range(len(user_seq))]}, logger=logging.getLogger("golden"),
             train_seq_len=[], valid_sample_locations=[])
    dataload.InteractionData._get_valid_sample_loc_for_train(loc)
    out = dict(L=L, P=P, Pe=Pe, C=C, N=N, user_seq=user_seq, event_seq=ev_seq, item_tags=tags,
               train_seq_len=list(loc.train_seq_len), valid_sample_locations=list(loc.valid_sample_locations),
               base_config=dict(base), train={}, eval={})
    item_to_info = [{}] + [{"tag_category": tags[i].tolist()} for i in range(1, N)]
    pools = [torch.nonzero(tags[:, c]).flatten().tolist() for c in range(C)]
    for variant, over in (("item_bycat", dict(loss="prior", category_by="item", neg_sample_by_cat=True)),
                          ("event", dict(loss="prior", category_by="event", neg_sample_by_cat=False)),
                          ("nce", dict(loss="nce", category_by="item", neg_sample_by_cat=False))):
        cfg = rh.RefConfig(dict(base, **over))
        fake = ns(item_num=N, user_num=len(user_seq), user_seq=user_seq, event_seq=ev_seq,
                  train_seq_len=loc.train_seq_len, valid_sample_locations=loc.valid_sample_locations,
                  id2token={"item_id": list(range(N)), "user_id": list(range(len(user_seq)))},
                  item_interact_weights=[0], item_weights_by_cat=[[0]], int_category_to_item_id=pools,
                  category_counts={f"cat{i}": len(pools[i]) for i in range(C)}, item_to_info=item_to_info,
                  category_to_int={f"cat{i}": i for i in range(C)}, user_cluster_list=[0])
        ds = SEQTrainDataset(cfg, fake)
        ds.rng = np.random.default_rng(123)
        rows = [ds[i] for i in range(len(ds))]
        out["train"][variant] = dict(config=dict(cfg), items=torch.stack([r[0] for r in rows]),
                                     neg=torch.stack([r[1] for r in rows]), mask=torch.stack([r[2] for r in rows]),
                                     tags=torch.stack([r[3] for r in rows]))
        for phase in ("valid", "test"):
            es = SeqEvalDataset(cfg, fake, phase=phase)
            batch = seq_eval_collate([es[i] for i in range(len(es))])
            user_ids, item_seq, item_target, (hu, hi), positive_u, _time, target_tags, outlier = batch
            out["eval"][(variant, phase)] = dict(user_ids=user_ids, item_seq=item_seq, item_target=item_target,
                                                 history_u=hu, history_i=hi, positive_u=positive_u,
                                                 target_tags=target_tags)
    return out


def schedules(lr_scheduler):
    out = {}
    for name, fn in (("cosine", lr_scheduler.get_cosine_schedule_with_warmup),
                     ("linear", lr_scheduler.get_linear_schedule_with_warmup)):
        for warm, total in ((12.5, 100), (0, 40), (30.0, 30000 // 100)):
            p = torch.nn.Parameter(torch.zeros(1))
            opt = torch.optim.SGD([p], lr=3e-3)
            sch = fn(opt, num_warmup_steps=warm, num_training_steps=total)
            vals = []
            for _ in range(total + 5):
                vals.append(opt.param_groups[0]["lr"])
                opt.step()
                sch.step()
            out[(name, warm, total)] = vals
    return out


def main():
    assert rh.available(), "needs /root/reference"
    dataload, SEQTrainDataset, SeqEvalDataset, seq_eval_collate, lr_scheduler = load_reference_modules()
    fx = dict(prior_dicts=prior_dicts(dataload), item_features=item_features(dataload),
              data_layer=data_layer(dataload, SEQTrainDataset, SeqEvalDataset, seq_eval_collate),
              schedules=schedules(lr_scheduler))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_layer.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
