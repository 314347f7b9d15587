"""Human priors -> decode heads: adapter for the reference's prior dictionaries and GPU construction of new ones.

Part 1: the reference turns the hand-curated outputs of item-clustering.py / user-clustering.py
(`REC/data/<dataset>_tag_dict.py`, `<dataset>_cluster_dict.py`, `<dataset>_user_cluster_dict.py`: a module-level
`tag_to_general`) into the quantities that define the heads (data/dataload.py:347-371):

    category_by in {item, user}:  spec = tag_to_general[tag_version]
        category_counts  = spec['category_counts']           -> prior_loss_weight (hstu.py:503-510)
        tag_to_category  = spec['tag_to_category']           -> item -> category multi-hot (dataload.py:226-246)
        category_to_int  = {cat: i for i, cat in enumerate(sorted(category_counts))}
        int_to_category  = inverse                            -> config['int_to_category'], head order
    category_by == 'event':       category_counts / category_to_int are given directly (merrec_2000_tag_dict.py)

`PriorSpec.from_source` reads such a module (by file path or import name) or a plain dict; `item_tag_table` builds
the bool [N, C] item -> category table and the per-category item pools (dataload.py:312-327
`int_category_to_item_id`) that the negative sampler and the eval masks consume.

Part 2 lives in `build_prior_from_interactions` below: co-occurrence graph + community detection on
the GPU (item-clustering.py:152-162, 227-250; user-clustering.py:218-307), producing a dict in the same format.
"""
import importlib
import importlib.util
import os

import torch


class PriorSpec(object):
    """What the model / data layer read from a prior dictionary (dataload.py:347-371)."""

    def __init__(self, category_counts, category_to_int, tag_to_category=None, category_percent=None):
        self.category_counts = dict(category_counts)
        self.category_to_int = dict(category_to_int)
        self.int_to_category = {v: k for k, v in self.category_to_int.items()}
        self.tag_to_category = tag_to_category
        self.category_percent = category_percent
        n = len(self.category_to_int)
        assert set(self.int_to_category) == set(range(n)), \
            f"config[int_to_category] keys must be 0..{n - 1}"          # dataload.py:330-332

    @property
    def num_categories(self):
        return len(self.category_to_int)

    @staticmethod
    def _tag_to_general(source):
        if isinstance(source, dict):
            return source
        if isinstance(source, str) and os.path.isfile(source):
            spec = importlib.util.spec_from_file_location("_b200rec_prior_dict", source)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.tag_to_general
        return importlib.import_module(source).tag_to_general

    @classmethod
    def from_source(cls, source, category_by="item", tag_version=None):
        """source: a `tag_to_general` dict, the path of a `*_dict.py` file, or an importable module name
        (the reference resolves `REC.data.{dataset}_tag_dict` / `_cluster_dict` / `_user_cluster_dict`)."""
        ttg = cls._tag_to_general(source)
        if category_by in ("item", "user"):
            if tag_version is None or tag_version not in ttg:
                raise KeyError(f"tag_version {tag_version!r} not in {sorted(ttg)}")   # e.g. EB-NeRD v3/v16 are not shipped
            spec = ttg[tag_version]
            counts = spec["category_counts"]
            c2i = {cat: idx for idx, cat in enumerate(sorted(counts.keys()))}           # dataload.py:361-362
            return cls(counts, c2i, spec["tag_to_category"], spec.get("category_percent"))
        if category_by == "event":
            return cls(ttg["category_counts"], ttg["category_to_int"], None, ttg.get("category_percent"))   # :364-366
        raise ValueError(f"category_by = {category_by} is not defined.")

    def prior_loss_weight(self):
        """hstu.py:503-510 (weighted_prior_loss): count_c / sum of counts, in head order."""
        tot = float(sum(self.category_counts.values()))
        w = [0.0] * self.num_categories
        for name, cnt in self.category_counts.items():
            w[self.category_to_int[name]] = cnt / tot
        return w

    def apply_to_config(self, config):
        """Sets the config keys the reference derives from the dictionary (dataload.py:363,366; run.py)."""
        config["int_to_category"] = dict(self.int_to_category)
        config["eval_num_cats"] = self.num_categories
        return config

    def dataload(self, item_num):
        """The `dataload` argument of HSTU(config, dataload): item_num, category_counts, category_to_int."""
        from .synth import Dataload
        return Dataload(item_num, self.category_counts, self.category_to_int)


def item_tag_table(item_tags, spec, item_num=None):
    """item_tags: per item id (index 0 = padding placeholder) the item's raw tag / cluster id (None: unknown item).
    Returns (tags bool [N, C], pools: list of C int64 tensors of item ids) exactly as dataload.py:226-246 builds
    `item_to_info[i]['tag_category']` (multi-hot over int_to_category order; tags without a mapping -> all False)
    and dataload.py:312-327 builds `int_category_to_item_id` (item ids per category, ascending item order)."""
    N = len(item_tags) if item_num is None else item_num
    C = spec.num_categories
    tags = torch.zeros((N, C), dtype=torch.bool)
    t2c = spec.tag_to_category or {}
    c2i = spec.category_to_int
    for i in range(1, N):
        t = item_tags[i]
        if t is None:
            continue
        for cat in t2c.get(t, []):
            j = c2i.get(cat)
            if j is not None:
                tags[i, j] = True
    pools = [torch.nonzero(tags[:, c], as_tuple=False).flatten() for c in range(C)]
    return tags, pools


def user_cluster_one_hot(user_clusters, spec):
    """category_by == 'user' (trainset.py:44-47, evalset.py:21-23): one-hot user cluster rows [U, n_clusters]."""
    n = max(spec.category_to_int.values()) + 1
    return torch.nn.functional.one_hot(torch.as_tensor(user_clusters, dtype=torch.int64), n)
