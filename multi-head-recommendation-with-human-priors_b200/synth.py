"""Synthetic configs and batches with the reference's input layouts.

Layouts follow the reference data layer (not ported):
  * train batch = (items i64[B,L+P], neg_items i64[B,C+1 | 1,n], attn_mask i64[B,L+P],
    pos_tag_categories i64[B,L+P,C])          -- trainset.py:124-177, hstu.py:631-636
  * eval batch  = item_seq i64[B,L] left zero-padded, item_target i64[B,P_e],
    target_tags [B,P_e,C], history (u,i)      -- evalset.py:81-155, collate_fn.py:59-90
All tensors are CPU tensors; callers move them.
"""
import math

import torch


class Config(dict):
    """dict with the reference Config lookup rules (configurator.py:142-153):
    missing key -> None; get(k, d) -> d when the stored value is None."""

    def __getitem__(self, k):
        return dict.get(self, k, None)

    def get(self, k, d=None):
        v = dict.get(self, k, None)
        return d if v is None else v

    def __contains__(self, k):
        return dict.get(self, k, None) is not None


class Dataload:
    """The three attributes HSTU.__init__ reads from `dataload` (hstu.py:346,504-507)."""

    def __init__(self, item_num, category_counts=None, category_to_int=None):
        self.item_num = item_num
        self.category_counts = category_counts or {}
        self.category_to_int = category_to_int or {}


_COMMON = dict(
    hidden_act="silu", enable_relative_attention_bias=True, medusa_lambda=0.99,
    split_mode="combine", pos_sample_mix_ratio=0, prior_switch=None, nce_thres=0.99,
    hidden_dropout_prob=0.0, attn_dropout_prob=0.0, topk=[5, 10, 50, 200],
    metrics=["Recall", "NDCG"], shared_metrics=["Entropy"], metric_decimal_place=7,
    outlier_user_metrics=None, eval_by_cat=True, category_by="item", seed=2020,
)

# name -> (model/config keys, data-shape keys).  See SURVEY.md §8 header for A-E.
PRESETS = {
    # A: small, CPU-runnable (BASELINE.json configs[0])
    "A": dict(n_layers=2, n_heads=1, item_embedding_size=64, hstu_embedding_size=64,
              MAX_ITEM_LIST_LENGTH=20, pred_len=1, eval_pred_len=1, medusa_num_layers=0,
              num_segment_head=1, num_prior_head=1, head_interaction="multiplicative",
              loss="nce", neg_sample_by_cat=False, weighted_prior_loss=False, eval_num_cats=2,
              train_batch_size=64, num_negatives=512, item_num=10000, eval_batch_size=64),
    "A2": dict(n_layers=2, n_heads=2, item_embedding_size=64, hstu_embedding_size=64,
               MAX_ITEM_LIST_LENGTH=20, pred_len=1, eval_pred_len=1, medusa_num_layers=0,
               num_segment_head=1, num_prior_head=1, head_interaction="multiplicative",
               loss="nce", neg_sample_by_cat=False, weighted_prior_loss=False, eval_num_cats=2,
               train_batch_size=64, num_negatives=512, item_num=10000, eval_batch_size=64),
    # B: Pixel8M-prior (reproduce/HSTU-Pixel8M-prior.slurm:34-60), 1 GPU
    "B": dict(n_layers=16, n_heads=16, item_embedding_size=1024, hstu_embedding_size=1024,
              MAX_ITEM_LIST_LENGTH=50, pred_len=8, eval_pred_len=8, medusa_num_layers=1,
              num_segment_head=4, num_prior_head=8, head_interaction="additive",
              loss="prior", neg_sample_by_cat=True, weighted_prior_loss=True, eval_num_cats=8,
              train_batch_size=128, num_negatives=8192, item_num=450000, eval_batch_size=256,
              given_prior_len=8, prior_given_at_test=False, hidden_dropout_prob=0.2),   # IDNet/hstu-size4.yaml
    # C: MerRec-prior (reproduce/HSTU-merrec-prior.slurm:33-66)
    "C": dict(n_layers=16, n_heads=16, item_embedding_size=1024, hstu_embedding_size=1024,
              MAX_ITEM_LIST_LENGTH=400, pred_len=1, eval_pred_len=1, medusa_num_layers=1,
              num_segment_head=1, num_prior_head=6, head_interaction="multiplicative",
              loss="prior", neg_sample_by_cat=False, weighted_prior_loss=True, eval_num_cats=6,
              train_batch_size=64, num_negatives=4096, item_num=5000000, eval_batch_size=256,
              fix_temp=True, category_by="event", prior_given_at_test=True, given_prior_len=1,
              hidden_dropout_prob=0.2),
    # D: EB-NeRD prior-mult (reproduce/HSTU-EBNerd-prior-mult.slurm:33-66)
    "D": dict(n_layers=8, n_heads=8, item_embedding_size=256, hstu_embedding_size=256,
              MAX_ITEM_LIST_LENGTH=50, pred_len=8, eval_pred_len=8, medusa_num_layers=1,
              num_segment_head=1, num_prior_head=7, head_interaction="multiplicative",
              loss="prior", neg_sample_by_cat=True, weighted_prior_loss=True, eval_num_cats=7,
              train_batch_size=128, num_negatives=8192, item_num=200000, eval_batch_size=256,
              given_prior_len=8, prior_given_at_test=False, eval_by_cat=False, hidden_dropout_prob=0.1),  # size2
}


def make_config(name, **overrides):
    cfg = Config(_COMMON)
    cfg.update(PRESETS[name])
    cfg.update(overrides)
    C = cfg["num_prior_head"]
    cfg["int_to_category"] = {i: f"cat{i}" for i in range(max(C, cfg["eval_num_cats"]))}
    cfg["metrics_pred_len_list"] = metrics_pred_len_list(cfg["eval_pred_len"])
    cfg["name"] = name
    return cfg


def metrics_pred_len_list(eval_pred_len):
    """0-based list after run.py:91-99: [1] + eval_pred_len (+ eval_pred_len//2 if > 0), minus one, sorted."""
    lst = [1]
    if eval_pred_len not in lst:
        lst.append(eval_pred_len)
    half = eval_pred_len // 2
    if half > 0 and half not in lst:
        lst.append(half)
    return sorted(x - 1 for x in lst)


def make_dataload(cfg):
    C = cfg["num_prior_head"]
    # skewed counts so weighted_prior_loss is not uniform
    counts = {f"cat{i}": 100 * (i + 1) for i in range(C)}
    c2i = {f"cat{i}": i for i in range(C)}
    return Dataload(cfg["item_num"], counts, c2i)


def negatives_per_sample(cfg, world_size=1):
    """trainset.py:58-60."""
    return math.ceil(cfg["num_negatives"] / world_size / cfg["train_batch_size"])


def make_item_tags(cfg, gen):
    """bool [N, C] item->category table (category_by == 'item')."""
    N, C = cfg["item_num"], cfg["eval_num_cats"]
    tags = torch.rand(N, C, generator=gen) < 0.35
    none = ~tags.any(dim=1)
    fill = torch.randint(0, C, (N,), generator=gen)
    tags[none, fill[none]] = True
    return tags


def _zipf_items(shape, N, gen, alpha=1.05):
    # inverse-CDF sampling of a truncated power law over [1, N)
    u = torch.rand(shape, generator=gen, dtype=torch.float64)
    if abs(alpha - 1.0) < 1e-9:
        x = torch.exp(u * math.log(N - 1))
    else:
        a = 1.0 - alpha
        x = ((u * ((N - 1) ** a - 1.0)) + 1.0) ** (1.0 / a)
    return x.floor().clamp_(1, N - 1).to(torch.int64)


def make_train_batch(cfg, seed=0, rank=0, world_size=1, batch_size=None, item_tags=None, zipf=True):
    """Returns (items, neg_items, attn_mask, pos_tag_categories) on CPU."""
    gen = torch.Generator().manual_seed(1000 * seed + rank + 17)
    B = batch_size or cfg["train_batch_size"]
    L, P, N = cfg["MAX_ITEM_LIST_LENGTH"], cfg["pred_len"], cfg["item_num"]
    C = cfg["eval_num_cats"]  # tag width of the data layer (== num_prior_head for prior losses)
    n = negatives_per_sample(cfg, world_size)
    if zipf:
        items = _zipf_items((B, L + P), N, gen)
    else:
        items = torch.randint(1, N, (B, L + P), generator=gen)
    # context length: full with prob .6 else U[1, L]; left padded (mask 0, random ids)
    full = torch.rand(B, generator=gen) < 0.6
    ell = torch.where(full, torch.full((B,), L), torch.randint(1, L + 1, (B,), generator=gen))
    # targets: all P valid with prob .9 else right padded to U[1, P]
    pfull = torch.rand(B, generator=gen) < 0.9
    pl = torch.where(pfull, torch.full((B,), P), torch.randint(1, P + 1, (B,), generator=gen))
    pos = torch.arange(L + P).unsqueeze(0)
    mask = ((pos >= (L - ell).unsqueeze(1)) & (pos < (L + pl).unsqueeze(1))).to(torch.int64)
    by_cat = bool(cfg["neg_sample_by_cat"]) and cfg["loss"] == "prior"
    if cfg["category_by"] == "item":
        if item_tags is None:
            item_tags = make_item_tags(cfg, torch.Generator().manual_seed(4242))
        tags = item_tags[items].to(torch.int64)  # [B, L+P, C]
    else:
        # event priors: one-hot per position, zero on padding (trainset.py:139-145)
        ev = torch.randint(0, C, (B, L + P), generator=gen)
        tags = torch.nn.functional.one_hot(ev, C).to(torch.int64) * mask.unsqueeze(-1)
    if by_cat:
        neg = torch.empty(B, C + 1, n, dtype=torch.int64)
        for c in range(C):
            pool = torch.nonzero(item_tags[1:, c]).squeeze(1) + 1 if cfg["category_by"] == "item" \
                else torch.arange(1, N)
            idx = torch.randint(0, pool.numel(), (B, n), generator=gen)
            neg[:, c] = pool[idx]
        neg[:, C] = torch.randint(1, N, (B, n), generator=gen)
    else:
        neg = torch.randint(1, N, (B, 1, n), generator=gen)
    return items, neg, mask, tags


def make_eval_batch(cfg, seed=0, batch_size=None, item_tags=None):
    """Returns dict(item_seq, item_target, target_tags, history_index, positive_u)."""
    gen = torch.Generator().manual_seed(7000 + seed)
    B = batch_size or cfg["eval_batch_size"]
    L, Pe, N, C = cfg["MAX_ITEM_LIST_LENGTH"], cfg["eval_pred_len"], cfg["item_num"], cfg["eval_num_cats"]
    seq = _zipf_items((B, L), N, gen)
    full = torch.rand(B, generator=gen) < 0.6
    ell = torch.where(full, torch.full((B,), L), torch.randint(1, L + 1, (B,), generator=gen))
    pos = torch.arange(L).unsqueeze(0)
    seq = torch.where(pos >= (L - ell).unsqueeze(1), seq, torch.zeros_like(seq))
    target = _zipf_items((B, Pe), N, gen)
    if cfg["category_by"] == "item":
        if item_tags is None:
            item_tags = make_item_tags(cfg, torch.Generator().manual_seed(4242))
        target_tags = item_tags[target].to(torch.int64)
    else:
        ev = torch.randint(0, C, (B, Pe), generator=gen)
        target_tags = torch.nn.functional.one_hot(ev, C).to(torch.int64)
    u, p = torch.nonzero(seq, as_tuple=True)
    history_index = (u, seq[u, p])
    positive_u = torch.arange(B).unsqueeze(1).expand(B, Pe).contiguous()
    return dict(item_seq=seq, item_target=target, target_tags=target_tags,
                history_index=history_index, positive_u=positive_u)
