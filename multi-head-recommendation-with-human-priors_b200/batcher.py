"""GPU-side batch construction: the interaction data lives in HBM in CSR form and one kernel
per batch replaces the reference's per-sample Python in DataLoader workers (data/dataset/trainset.py:70-177,
evalset.py:81-155, collate_fn.py:59-90).  Output tensors have exactly the reference's batch layout, so they
feed `HSTU.forward` / `predict_topk` / `Collector.eval_batch_collect` unchanged.
"""
import numpy as np
import torch

from . import _lib as L


class InteractionData(object):
    """What the kernels need of the reference's `Data` object (data/dataload.py:86-194), as device arrays.

    user_seq      list (per user, index 0 unused like the reference's 1-based uid) of item-id lists
    train_seq_len per-user length of the training prefix (dataload.py: train_seq_len)
    event_seq     optional per-user event-type lists (category_by == 'event')
    item_tags     optional bool [N, C] item -> category table (category_by == 'item')
    """

    def __init__(self, user_seq, train_seq_len, item_num, max_seq_length, item_tags=None, event_seq=None,
                 device="cuda", sample_last_only=False, pred_len=1, include_empty_context=False):
        self.device = torch.device(device)
        self.item_num, self.L = int(item_num), int(max_seq_length)
        lens = np.asarray([len(s) for s in user_seq], dtype=np.int64)
        off = np.zeros(len(user_seq) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.concatenate([np.asarray(s, dtype=np.int64) for s in user_seq]) if off[-1] else np.zeros(0, np.int64)
        self.h_off, self.h_len = off, lens
        self.h_train_len = np.asarray(train_seq_len, dtype=np.int32)
        self.user_seq = torch.from_numpy(flat).to(self.device)
        self.user_off = torch.from_numpy(off).to(self.device)
        self.train_len = torch.from_numpy(self.h_train_len).to(self.device)
        self.event_seq = None
        if event_seq is not None:
            ev = np.concatenate([np.asarray(s, dtype=np.int32) for s in event_seq])
            self.event_seq = torch.from_numpy(ev).to(self.device)
        self.item_tags = None
        self.cat_items = self.cat_off = None
        self.C = 0
        if item_tags is not None:
            t = torch.as_tensor(item_tags).to(torch.uint8)
            self.C = t.shape[1]
            self.item_tags = t.contiguous().to(self.device)
            # per-category item pools (dataload.int_category_to_item_id), id 0 excluded
            pools = [torch.nonzero(t[1:, c], as_tuple=False).flatten() + 1 for c in range(self.C)]
            co = np.zeros(self.C + 1, dtype=np.int64)
            np.cumsum([p.numel() for p in pools], out=co[1:])
            self.cat_items = torch.cat(pools).to(torch.int64).to(self.device)
            self.cat_off = torch.from_numpy(co).to(self.device)
        # valid_sample_locations (dataload.py:165-194; the reference's max_item_list_len is L + 1): one window for
        # a short training prefix, else non-overlapping windows of stride L + 1 anchored at the end of the prefix
        # (the first of them may have an empty context: context_end == 0 when (n - 1) % (L + 1) == 0)
        uid, end = [], []
        stride = self.L + 1
        for u in range(1, len(user_seq)):
            n = int(self.h_train_len[u])
            if n <= 1:
                continue
            if sample_last_only:                                  # dataload.py:173-177 (Amazon Books)
                locs = [n - 1] if n < pred_len + 3 else [n - pred_len]
            elif n <= stride:
                locs = [n - 1]
            else:
                locs = list(range((n - 1) % stride, n, stride))
            if not include_empty_context:
                # a window with context_end == 0 has no query position: it adds no loss term and no gradient, only an
                # idle batch row.  Dropped by default; include_empty_context=True reproduces the reference's list.
                locs = [e for e in locs if e > 0]
            uid += [u] * len(locs)
            end += locs
        self.h_sample_uid = np.asarray(uid, dtype=np.int64)
        self.h_sample_end = np.asarray(end, dtype=np.int32)
        self.sample_uid = torch.from_numpy(self.h_sample_uid).to(self.device)
        self.sample_end = torch.from_numpy(self.h_sample_end).to(self.device)

    def __len__(self):
        return len(self.h_sample_uid)


class GpuTrainBatcher(object):
    """batch(indices, step) -> (items, neg_items, mask, tags) on the device + the host-side token count."""

    def __init__(self, data, config, world_size=1, seed=None):
        self.data, self.cfg = data, config
        self.L, self.P = config["MAX_ITEM_LIST_LENGTH"], config["pred_len"]
        self.pad_random = bool(config.get("pad_random_sample", True))
        self.by_cat = bool(config["neg_sample_by_cat"]) and config["loss"] == "prior"
        self.category_by = config["category_by"]
        nn_ = config["num_negatives"]
        B = config["train_batch_size"]
        self.n_neg = int(np.ceil(nn_ / world_size / B)) if nn_ else self.L          # trainset.py:58-63
        self.n_pools = data.C if (self.by_cat and data.cat_items is not None) else 0
        self.n_sets = self.n_pools + 1
        self.mix = float(config.get("neg_sample_mix_ratio", 0.0) or 0.0)
        self.return_tags = config["loss"] == "prior"
        self.seed = int(config.get("seed", 2020) if seed is None else seed)

    def batch(self, indices, step):
        d, dev = self.data, self.data.device
        idx_h = np.asarray(indices, dtype=np.int64)
        B, LP = len(idx_h), self.L + self.P
        idx = torch.from_numpy(idx_h).to(dev)
        items = torch.empty((B, LP), dtype=torch.int64, device=dev)
        neg = torch.empty((B, self.n_sets, self.n_neg), dtype=torch.int64, device=dev)
        mask = torch.empty((B, LP), dtype=torch.int64, device=dev)
        C = d.C if self.category_by == "item" else int(self.cfg["eval_num_cats"])
        use_item_tags = self.category_by == "item" and d.item_tags is not None
        tags = torch.empty((B, LP, C if self.return_tags else 0), dtype=torch.int64, device=dev)
        L.call("b200rec_build_train_batch", d.user_seq.data_ptr(), d.user_off.data_ptr(), d.train_len.data_ptr(),
               L.ptr(d.event_seq), d.sample_uid.data_ptr(), d.sample_end.data_ptr(), idx.data_ptr(), B, self.L, self.P,
               1 if self.pad_random else 0, d.item_num, self.n_sets, self.n_neg, L.ptr(d.cat_items), L.ptr(d.cat_off),
               self.n_pools, self.mix, L.ptr(d.item_tags) if use_item_tags else None, C, self.seed, int(step),
               items.data_ptr(), neg.data_ptr(), mask.data_ptr(), tags.data_ptr() if self.return_tags and C else None,
               L.stream())
        # context tokens per row = min(context_end, L): host metadata for the CUDA-graph bucket (no device sync)
        n_tokens = int(np.minimum(d.h_sample_end[idx_h], self.L).sum())
        return (items, neg, mask, tags), n_tokens


class GpuEvalBatcher(object):
    """batch(uids, phase) -> dict in the layout of collate_fn.seq_eval_collate (collate_fn.py:59-90)."""

    def __init__(self, data, config):
        self.data, self.cfg = data, config
        self.L, self.Pe = config["MAX_ITEM_LIST_LENGTH"], config["eval_pred_len"]
        self.category_by = config["category_by"]

    def batch(self, uids, phase="valid"):
        d, dev = self.data, self.data.device
        ph = 0 if phase == "valid" else 1
        u_h = np.asarray(uids, dtype=np.int64)
        B = len(u_h)
        n_hist = d.h_train_len[u_h].astype(np.int64) if ph == 0 else d.h_len[u_h] - self.Pe
        hoff_h = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(n_hist, out=hoff_h[1:])
        u = torch.from_numpy(u_h).to(dev)
        hoff = torch.from_numpy(hoff_h).to(dev)
        C = d.C if self.category_by == "item" else int(self.cfg["eval_num_cats"])
        use_item_tags = self.category_by == "item" and d.item_tags is not None
        item_seq = torch.empty((B, self.L), dtype=torch.int64, device=dev)
        target = torch.empty((B, self.Pe), dtype=torch.int64, device=dev)
        ttags = torch.empty((B, self.Pe, C), dtype=torch.int64, device=dev)
        hu = torch.empty(int(hoff_h[-1]), dtype=torch.int64, device=dev)
        hi = torch.empty(int(hoff_h[-1]), dtype=torch.int64, device=dev)
        L.call("b200rec_build_eval_batch", d.user_seq.data_ptr(), d.user_off.data_ptr(), d.train_len.data_ptr(),
               L.ptr(d.event_seq), u.data_ptr(), B, self.L, self.Pe, ph, L.ptr(d.item_tags) if use_item_tags else None, C,
               hoff.data_ptr(), item_seq.data_ptr(), target.data_ptr(), ttags.data_ptr(), hu.data_ptr(), hi.data_ptr(),
               L.stream())
        positive_u = torch.arange(B).unsqueeze(-1).repeat(1, self.Pe)
        return dict(user_ids=u, item_seq=item_seq, item_target=target, history_index=(hu, hi), positive_u=positive_u,
                    target_tags=ttags)
