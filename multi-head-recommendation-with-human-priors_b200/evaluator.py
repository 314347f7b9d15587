"""Collector / Evaluator with the reference's contract (code/REC/evaluator/collector.py:58-395,
evaluator.py:10-40, metrics.py:145-238, base_metric.py:37-94) for the `rec.topk` metrics path
(Recall, NDCG, Entropy).  The top-K, cross-head merge and hit matrix run in CUDA
(b200rec_score_mask_topk / b200rec_hit_matrix); the final per-user metric arithmetic is the same
small float64 numpy computation the reference does on the host.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L


class DataStruct(object):
    """collector.py:13-55."""

    def __init__(self):
        self._tensor_lists = {}
        self._data_dict = {}

    def __getitem__(self, name):
        return self._data_dict[name]

    def __setitem__(self, name, value):
        self._data_dict[name] = value

    def __delitem__(self, name):
        self._data_dict.pop(name)

    def __contains__(self, key):
        return key in self._data_dict

    def get(self, name):
        if name not in self._data_dict:
            raise IndexError("Can not load the data without registration !")
        return self[name]

    def set(self, name, value):
        self._data_dict[name] = value

    def update_tensor(self, name, value):
        self._tensor_lists.setdefault(name, []).append(value.detach().cpu().clone())

    def finalize_tensors(self):
        for name, lst in self._tensor_lists.items():
            if lst:
                self._data_dict[name] = torch.cat(lst, dim=0)
        self._tensor_lists.clear()


def _num_heads(config):
    if config["head_interaction"] in ("multiplicative", "hierarchical"):
        return config["num_segment_head"] * config["num_prior_head"]
    if config["head_interaction"] == "additive":
        return config["num_segment_head"] + config["num_prior_head"]
    raise ValueError(f'Unknown head_interaction: {config["head_interaction"]}')


class Collector(object):
    def __init__(self, config):
        self.config = config
        self.metrics_pred_len_list = list(config["metrics_pred_len_list"])
        self.eval_pred_len = config["eval_pred_len"]
        self.data_struct = {p: DataStruct() for p in self.metrics_pred_len_list}
        self.data_struct[-1] = DataStruct()
        self.topk = config["topk"]
        self.medusa_num_heads = _num_heads(config)
        self.split_mode = config["split_mode"]
        self.all_tags = None
        self.last_topk = None

    def set_all_tags(self, item_tags):
        self.all_tags = item_tags

    def reset_all_tags(self):
        self.all_tags = None

    def topk_from_scores(self, scores_tensor):
        """[B, H, N] (already masked) -> (idx, val, head) via the fused fold + radix-select kernel."""
        if not scores_tensor.is_cuda:
            raise L.B200RecError("Collector needs CUDA scores (there is no CPU path)")
        s = scores_tensor.float()
        if s.dim() == 2:
            s = s.unsqueeze(1)
        B, H, N = s.shape
        s = s.contiguous().view(B * H, N)
        K = max(self.topk)
        dev = s.device
        idx = torch.empty((B, K), dtype=torch.int64, device=dev)
        val = torch.empty((B, K), dtype=torch.float32, device=dev)
        hsrc = torch.empty((B, K), dtype=torch.int32, device=dev)
        ws_bytes = L.lib().b200rec_topk_workspace_bytes(B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        mode = 1 if (self.split_mode == "average" and H > 1) else 0
        if self.split_mode not in ("combine", "average"):
            raise ValueError(f"Unknown split_mode: {self.split_mode}")
        # masks were applied by predict() / the caller; id 0 is re-masked inside (idempotent)
        L.call("b200rec_score_mask_topk", s.data_ptr(), N, B, H, N, K, None, None, None, None, None, mode, 0, 1,
               idx.data_ptr(), val.data_ptr(), hsrc.data_ptr(), ws.data_ptr(), ws_bytes, L.stream())
        return idx, val, hsrc

    def eval_batch_collect(self, scores_tensor, positive_u, positive_i, tag_category=None, outlier_users=None,
                           log_detailed_results=False, topk=None):
        """collector.py:153-325 (`rec.topk` path).  `topk` may carry a precomputed (idx, val, head) triple
        from HSTU.predict_topk, in which case scores_tensor may be None."""
        if tag_category is not None:
            for p in self.metrics_pred_len_list:
                self.data_struct[p].update_tensor("rec.tgt_tags", torch.any(tag_category[:, :p + 1].bool(), dim=1))
        if outlier_users is not None:
            self.data_struct[self.eval_pred_len - 1].update_tensor("rec.outlier_users", outlier_users)
        idx, val, hsrc = topk if topk is not None else self.topk_from_scores(scores_tensor)
        self.last_topk = (idx, val, hsrc)
        B, K = idx.shape
        if self.all_tags is not None:
            self.data_struct[-1].update_tensor("rec.rec_tags", self.all_tags.to(idx.device)[idx])
        positive_i = positive_i.to(idx.device).to(torch.int64).contiguous()
        Pe = positive_i.shape[1]
        plist = np.asarray(self.metrics_pred_len_list, dtype=np.int32)
        out = torch.empty((len(plist), B, K + 1), dtype=torch.int32, device=idx.device)
        L.call("b200rec_hit_matrix", idx.data_ptr(), positive_i.data_ptr(), B, K, Pe, plist.ctypes.data, len(plist),
               out.data_ptr(), L.stream())
        for q, p in enumerate(self.metrics_pred_len_list):
            self.data_struct[p].update_tensor("rec.topk", out[q])
        top = {}
        if log_detailed_results:
            top["values"] = val.float().cpu().numpy()
            top["head_source"] = hsrc.cpu().numpy()
            top["idx"] = idx.cpu().tolist()
        return top

    def get_data_struct(self, pred_idx=0):
        import copy
        self.data_struct[pred_idx].finalize_tensors()
        returned = copy.deepcopy(self.data_struct[pred_idx])
        for key in ["rec.rec_tags", "rec.tgt_tags", "rec.outlier_users", "rec.topk"]:
            if key in self.data_struct[pred_idx]:
                del self.data_struct[pred_idx][key]
        return returned


# ---- metrics (metrics.py:145-238, base_metric.py:37-94): SUMS over users, the trainer divides -------------
class _TopkMetric(object):
    def __init__(self, config):
        self.topk = config["topk"]
        self.num_prior_categories = config["eval_num_cats"]
        self.eval_by_cat = config.get("eval_by_cat", True)
        self.int_to_category = config["int_to_category"]

    @staticmethod
    def used_info(dataobject, kmax):
        rec = dataobject.get("rec.topk")
        hit, pos_len = torch.split(rec, [kmax, 1], dim=1)
        return hit.to(torch.bool).numpy(), pos_len.squeeze(-1).numpy()

    def topk_result(self, metric, value, num_samples=None, prefix=None):
        out = {}
        s = value.sum(axis=0)
        for k in self.topk:
            key = f"{metric}@{k}" if prefix is None else f"{prefix}-{metric}@{k}"
            out[key] = (s[k - 1], num_samples) if num_samples is not None else s[k - 1]
        return out

    def calculate_metric(self, dataobject, pred_len=1):
        hit, pos_len = self.used_info(dataobject, max(self.topk))
        res = self.topk_result(self.name, self.metric_info(hit, pos_len))
        if self.num_prior_categories > 1 and self.eval_by_cat and "rec.tgt_tags" in dataobject:
            tags = dataobject.get("rec.tgt_tags")
            for c in range(self.num_prior_categories):
                sel = tags[:, c].numpy().astype(bool)
                res.update(self.topk_result(self.name, self.metric_info(hit[sel], pos_len[sel]),
                                            num_samples=int(sel.sum()), prefix=self.int_to_category[c]))
        return res


class Recall(_TopkMetric):
    name = "recall"

    def metric_info(self, pos_index, pos_len):
        return np.cumsum(pos_index, axis=1) / pos_len.reshape(-1, 1)


class NDCG(_TopkMetric):
    name = "ndcg"

    def metric_info(self, pos_index, pos_len):
        K = pos_index.shape[1]
        idcg_len = np.minimum(pos_len, K)
        ranks = np.arange(1, K + 1, dtype=np.float64)
        disc = 1.0 / np.log2(ranks + 1)
        idcg = np.broadcast_to(np.cumsum(disc), pos_index.shape).copy()
        for row, n in enumerate(idcg_len):
            idcg[row, n:] = idcg[row, n - 1]
        dcg = np.cumsum(np.where(pos_index, disc, 0.0), axis=1)
        return dcg / idcg


class Entropy(object):
    """metrics.py:17-41."""

    def __init__(self, config):
        self.topk = config["topk"]

    def calculate_metric(self, dataobject, pred_len=1):
        rec_tags = dataobject.get("rec.rec_tags").numpy()
        counts = np.cumsum(rec_tags * 1.0, axis=1)
        out = {}
        for k in self.topk:
            p = counts[:, k - 1, :] / counts[:, k - 1, :].sum(axis=1, keepdims=True)
            with np.errstate(divide="ignore", invalid="ignore"):
                ent = -np.sum(np.where(p > 0, p * np.log2(p), 0.0), axis=1)
            out[f"Entropy@{k}"] = ent.sum(axis=0)
        return out


_METRICS = {"recall": Recall, "ndcg": NDCG, "entropy": Entropy}


class Evaluator(object):
    """evaluator.py:10-40."""

    def __init__(self, config):
        self.config = config
        self.metrics = [m.lower() for m in config["metrics"]]
        self.shared_metrics = [m.lower() for m in (config["shared_metrics"] or [])]
        self.metric_class = {}
        for m in self.metrics + self.shared_metrics:
            if m not in _METRICS:
                raise NotImplementedError(f"metric {m} is outside the built hot path (Recall, NDCG, Entropy)")
            self.metric_class[m] = _METRICS[m](config)

    def evaluate(self, dataobject, pred_len=1):
        result = OrderedDict()
        for m in (self.shared_metrics if pred_len == -1 else self.metrics):
            result.update(self.metric_class[m].calculate_metric(dataobject, pred_len=pred_len))
        return result
