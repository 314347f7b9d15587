"""B200-native HSTU with multi-head prior-guided decoding: drop-in for the reference model class.

Mirrors `REC.model.IDNet.hstu.HSTU` (reference code/REC/model/IDNet/hstu.py:331-1030): same
constructor `HSTU(config, dataload)`, same `forward(interaction) -> dict` with `'loss'` and the
logging keys, same `predict(...)`, `compute_item_all()`, parameter / buffer names and init
order (so a reference checkpoint loads with `load_state_dict`, and the same seed gives the same
initial weights).  All arithmetic runs in hand-written sm_100a kernels behind the C ABI in
include/b200rec.h; torch only owns memory and streams.  No CPU path: tensors must be CUDA.

Restructurings relative to the reference (each proven equal on valid rows):
  * jagged body: only valid tokens are computed (hstu.py:645 mask -> seq_off / tok_pos);
  * one query row per (head, token) shared by all prediction offsets, one fix-mask row per
    target position (hstu.py:600-619 recomputes both per (b, p, l));
  * deterministic sorted-segment embedding gradient instead of atomic index_add;
  * eval: top-K of the max over heads == the collector's per-head top-K + sort + dedupe.
Not implemented (raise at construction, never a silent fallback): see DESIGN.md section 6.
"""
from collections import defaultdict
from logging import getLogger

import numpy as np
import ctypes

import torch
import torch.nn as nn

from . import _lib as L
from . import parallel

LN_EPS = 1e-6  # hstu.py:177


def truncated_normal_(x, mean, std):
    """hstu.py:23-31 (resample-4 truncated normal); same RNG consumption as the reference."""
    with torch.no_grad():
        tmp = x.new_empty(x.shape + (4,)).normal_()
        valid = (tmp < 2) & (tmp > -2)
        ind = valid.max(-1, keepdim=True)[1]
        x.data.copy_(tmp.gather(-1, ind).squeeze(-1))
        x.data.mul_(std).add_(mean)
    return x


class _RelBias(nn.Module):
    """Parameters of RelativeBucketedTimeAndPositionBasedBias (hstu.py:74-97).  The reference builds
    them but never applies them in the block; kept for state-dict parity, no gradient."""

    def __init__(self, max_seq_len, num_buckets):
        super().__init__()
        self._ts_w = nn.Parameter(torch.empty(num_buckets + 1).normal_(mean=0, std=0.02))
        self._pos_w = nn.Parameter(torch.empty(2 * max_seq_len - 1).normal_(mean=0, std=0.02))


class _STU(nn.Module):
    """Parameter holder for one SequentialTransductionUnitJagged (hstu.py:163-211)."""

    def __init__(self, D, rel_bias):
        super().__init__()
        self._rel_attn_bias = rel_bias
        self._uvqk = nn.Parameter(torch.empty((D, 4 * D)).normal_(mean=0, std=0.02))
        self._o = nn.Linear(D, D)
        nn.init.xavier_uniform_(self._o.weight)


class _Body(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self._attention_layers = nn.ModuleList(blocks)


class _ResBlock(nn.Module):
    """Parameter holder for llm_heads.ResBlock (llm_heads.py:16-27): [LayerNorm(D) when use_norm,] linear [D, D] + bias,
    created in the reference's order."""

    def __init__(self, D, zero_init=True, use_norm=False):
        super().__init__()
        self.use_norm = use_norm
        if use_norm:
            self.norm = nn.LayerNorm(D)
        self.linear = nn.Linear(D, D)
        if zero_init:
            nn.init.zeros_(self.linear.weight)
        else:                                   # llm_heads.py:24-25 (same RNG consumption as the reference)
            nn.init.trunc_normal_(self.linear.weight, std=0.02)


class _Job:
    """One NCE contraction: query head `head` against negative set `nset`, serving the prediction
    offsets in `p_mask` with token-validity column `col` and loss weight `w`."""
    __slots__ = ("part", "cat", "head", "nset", "p_mask", "col", "w", "seg")

    def __init__(self, part, cat, head, nset, p_mask, col, w, seg):
        self.part, self.cat, self.head, self.nset = part, cat, head, nset
        self.p_mask, self.col, self.w, self.seg = p_mask, col, w, seg


class HSTU(nn.Module):
    def __init__(self, config, dataload, compute_dtype=torch.bfloat16):
        super().__init__()
        self.logger = getLogger()
        self.compute_dtype = compute_dtype
        self.item_num = dataload.item_num
        D_item = config["item_embedding_size"]
        D = config["hstu_embedding_size"]
        self._item_embedding_dim = D_item
        self._hstu_embedding_dim = D
        self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"]
        self.pred_len = config["pred_len"]
        self.medusa_lambda = config["medusa_lambda"]
        self.num_segment_head = config["num_segment_head"]
        self.num_prior_head = config["num_prior_head"]
        self.head_interaction = config["head_interaction"]
        if self.head_interaction == "multiplicative":
            self.medusa_num_heads = self.num_segment_head * self.num_prior_head
        elif self.head_interaction == "additive":
            self.medusa_num_heads = self.num_segment_head + self.num_prior_head
        elif self.head_interaction == "hierarchical":
            self.medusa_num_heads = self.num_segment_head * self.num_prior_head
        else:
            raise ValueError(f'Unknown head_interaction: {config["head_interaction"]}')
        self.medusa_num_layers = config["medusa_num_layers"]
        self.category_by = config["category_by"]
        self._num_blocks = config["n_layers"]
        self._num_heads = config["n_heads"]
        self._dqk = D // self._num_heads
        if (config["hidden_act"] or "silu") != "silu":
            raise NotImplementedError("only hidden_act='silu' is built")
        self._linear_dropout_rate = config["hidden_dropout_prob"] or 0.0
        self._enable_relative_attention_bias = bool(config["enable_relative_attention_bias"])
        # The reference builds the bias module and never applies it: parity = OFF.  This flag applies
        # the position part, silu(q k^T + pos_w[N - 1 - (i - j)] + ts_w[0]) / n (timestamps are not part of the batch, so
        # every pair falls into time bucket 0), through the SIMT attention kernels.
        self.apply_rel_bias = bool(config.get("apply_relative_attention_bias", False)) and \
            self._enable_relative_attention_bias
        # ---- parameters, created in the reference's order (hstu.py:380-425, 486-493) ----
        self.position_embedding = nn.Embedding(self.max_seq_length + 1, D)
        blocks = []
        for _ in range(self._num_blocks):
            rb = _RelBias(2 * self.max_seq_length, 128) if self._enable_relative_attention_bias else None
            blocks.append(_STU(D, rb))
        self._hstu = _Body(blocks)
        self.item_embedding = nn.Embedding(self.item_num, D_item, padding_idx=0)
        # hstu.py:414: bias-free Linear when the table is narrower / wider than the HSTU width
        self.item_id_proj_tower = nn.Identity() if D_item == D else nn.Linear(D_item, D, bias=False)
        self.loss = config["loss"]
        self.neg_sample_by_cat = bool(config["neg_sample_by_cat"]) and self.loss == "prior"
        if (config["pos_sample_mix_ratio"] or 0) > 0:
            raise NotImplementedError("pos_sample_mix_ratio > 0 draws torch RNG inside forward; not built")
        if self.loss not in ("nce", "prior"):
            raise NotImplementedError(f"loss={self.loss} is not supported")
        if config["fix_temp"]:
            self.register_buffer("logit_scale", torch.tensor(np.log(1 / 0.05)))
        else:
            self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.05))
        self.nce_thres = config["nce_thres"] if config["nce_thres"] else 0.99
        self.seg_len = self.pred_len
        if self.medusa_num_layers > 0:
            assert self.pred_len % self.num_segment_head == 0, "pred_len must be divisible by the number of segments"
            self.seg_len = self.pred_len // self.num_segment_head
        hd = torch.tensor([self.medusa_lambda ** i for i in range(self.pred_len)])
        self.register_buffer("horizon_discount", hd / sum(hd))
        if self.medusa_num_layers == 0:
            self.medusa_head = nn.ModuleList([nn.Identity() for _ in range(self.medusa_num_heads)])
            self.prior_loss_weight = [1 / self.num_prior_head] * self.num_prior_head
        else:
            if self.head_interaction == "hierarchical":
                # hstu.py:443-483: per-category block, then per-(category, segment) block; distinct ResBlocks per layer
                # options (hstu.py:444-451): LayerNorm inside every ResBlock, a LN-Linear-SiLU-Linear bottleneck in front of
                # the category block, ONE segment block shared by every (category, segment), a learned segment offset
                nl = self.medusa_num_layers
                self.head_norm = bool(config.get("head_norm", False))
                self.cat_bottleneck = bool(config.get("cat_bottleneck", False))
                self.cat_bottleneck_dim = int(config.get("cat_bottleneck_dim", D // 2))
                if self.cat_bottleneck and compute_dtype != torch.float32 and self.cat_bottleneck_dim % 8 != 0:
                    raise NotImplementedError("cat_bottleneck_dim must be a multiple of 8 in bf16 mode (16-byte TMA row pitch "
                                              f"of the tcgen05 GEMM operands), got {self.cat_bottleneck_dim}")
                self.share_seg_weights = bool(config.get("share_seg_weights", False))
                self.use_seg_embed = bool(config.get("segment_embed", False))
                if self.use_seg_embed:
                    self.segment_emb = nn.Embedding(self.num_segment_head, D)

                def _blocks():
                    return [_ResBlock(D, zero_init=False, use_norm=self.head_norm) for _ in range(nl)]

                def _cat_block():
                    layers = []
                    if self.cat_bottleneck:
                        layers += [nn.LayerNorm(D), nn.Linear(D, self.cat_bottleneck_dim), nn.SiLU(),
                                   nn.Linear(self.cat_bottleneck_dim, D)]
                    return nn.Sequential(*(layers + _blocks()))

                self.medusa_cat_head = nn.ModuleList([_cat_block() for _ in range(self.num_prior_head)])
                if self.share_seg_weights:
                    shared_seg = nn.Sequential(*_blocks())
                    self.medusa_seg_head = nn.ModuleList(
                        [nn.ModuleList([shared_seg for _ in range(self.num_segment_head)])
                         for _ in range(self.num_prior_head)])
                else:
                    self.medusa_seg_head = nn.ModuleList(
                        [nn.ModuleList([nn.Sequential(*_blocks()) for _ in range(self.num_segment_head)])
                         for _ in range(self.num_prior_head)])
            else:
                self.medusa_head = nn.ModuleList(
                    [nn.Sequential(*([_ResBlock(D)] * self.medusa_num_layers)) for _ in range(self.medusa_num_heads)])
            self.weighted_prior_loss = config["weighted_prior_loss"]
            if self.loss != "prior":
                assert self.num_prior_head == 1, "Only prior loss is allowed for num_prior_head > 1"
            if self.loss == "prior" and self.weighted_prior_loss:
                tot = sum(dataload.category_counts.values())
                self.prior_loss_weight = [0 for _ in range(self.num_prior_head)]
                for name, cnt in dataload.category_counts.items():
                    self.prior_loss_weight[dataload.category_to_int[name]] = cnt / tot
            else:
                self.prior_loss_weight = [1 / self.num_prior_head] * self.num_prior_head
            self.aux_cat_head = None
            if self.loss == "prior" and config["prior_switch"] is not None:
                # hstu.py:512-544: one Linear(D -> 1) per prior category on the body output ('in')
                if config["prior_switch"] == "in_out":
                    raise NotImplementedError("prior_switch='in_out' (aux input = [body output, head output]) is not built; "
                                              "'in' is")
                if config["prior_switch"] != "in":
                    config["prior_switch"] = None          # hstu.py:538-540: unknown value -> switched off
                else:
                    assert config["split_mode"] == "combine"
                    self.use_asym_switch_loss = bool(config.get("asym_switch_loss", False))
                    self.switch_last_only = bool(config.get("switch_last_only", False))
                    self.gamma_pos = float(config.get("gamma_pos", 4.0))
                    self.gamma_neg = float(config.get("gamma_neg", 0.0))
                    self.master_switch = bool(config.get("master_switch", False))
                    self.aux_cat_head = nn.ModuleList([nn.Linear(D, 1) for _ in range(self.num_prior_head)])
                    if self.master_switch:
                        for i in range(1, self.num_prior_head):
                            for p_ in self.aux_cat_head[i].parameters():
                                p_.requires_grad_(False)
                    self.prior_switch_loss_weight = float(config["prior_switch_loss_weight"])
        self.prior_switch = config["prior_switch"] if (self.medusa_num_layers > 0 and self.loss == "prior") else None
        self.use_prior_switch_test = bool(config.get("use_prior_switch_test", False))
        self.detach_aux_in = bool(config.get("detach_aux_in", False))
        self.eval_pred_len = config["eval_pred_len"]
        self.prior_given_at_test = config.get("prior_given_at_test", False)
        self.given_prior_len = config.get("given_prior_len", self.eval_pred_len) if self.prior_given_at_test \
            else self.eval_pred_len
        self.int_to_category = config["int_to_category"]
        self.register_buffer("_attn_mask", torch.triu(
            torch.ones((self.max_seq_length, self.max_seq_length), dtype=torch.bool), diagonal=1))
        self.sparse_embedding_grad = bool(config.get("sparse_embedding_grad", False))
        self.use_tc_attention = bool(config.get("tc_attention", True))
        self.use_fused_eval = bool(config.get("fused_eval", True))
        self.use_fused_nce = bool(config.get("fused_nce", True))            # bf16 mode: softmax numerators from the GEMM epilogue
        self.use_streamed_eval = bool(config.get("streamed_eval", True))    # top-K candidates filtered in the scoring GEMM
        self.use_pruned_filter = bool(config.get("pruned_filter", True))    # false-negative filter via an upper bound
        self.share_negatives = bool(config.get("share_negatives", True))   # all-gather negatives across ranks
        self.dropout_seed = int(config.get("seed", 2020)) & 0xffffffff
        self._rng_step = None      # device counter feeding the Philox dropout stream
        self._drop_p_last = 0.0
        self.sharded_table = None  # parallel.ShardedTable once shard_item_table() was called
        self._shadow_buf, self._shadow_state, self._shadow_view = {}, {}, {}
        self._heads_upper = []     # (input, pre-activation) of the weight-tied decode-head layers above the first
        # set by FusedAdamW(lazy_table=True): bring table rows up to date before they are read
        self._table_sync = self._table_flush = None
        self._simt_len = 0         # extra sequence-length bound for the SIMT attention launch (static-token mode)
        self.emb_grad = None       # (uniq_ids, uniq_rows, n_uniq) of the last backward
        self._table_cache = None   # normalised compute-dtype item table for predict
        self._verbose = False
        self._switch_logs = {}
        self._debug = None         # dict: when set, forward / backward stash clones of key intermediates (diagnostics)
        self._readout = getattr(self, "_readout", None)   # comirec.ComiRec: multi-interest readout instead of decode heads
        self.reset_params()
        self._jobs = self._build_jobs()

    # ------------------------------------------------------------------ init (hstu.py:574-588)
    def reset_params(self):
        for name, p in self.named_parameters():
            if ("_hstu" in name) or ("_embedding_module" in name) or ("logit_scale" in name):
                continue
            truncated_normal_(p.data, mean=0.0, std=0.02)

    def _build_jobs(self):
        P, S, C = self.pred_len, self.num_segment_head, self.num_prior_head
        n_sets_global = C if self.neg_sample_by_cat else 0   # index of the global set in neg_items[:, -1]
        jobs = []

        def pmask(seg, seg_len):
            return sum(1 << p for p in range(P) if p // seg_len == seg)

        if self.loss == "nce" or (self.loss == "prior" and self.head_interaction == "additive"):
            n_seg = P // self.seg_len
            for s in range(n_seg):
                jobs.append(_Job("nce", -1, s, n_sets_global, pmask(s, self.seg_len), 0, 1.0, s))
        if self.loss == "prior":
            seg_len = P if self.head_interaction == "additive" else self.seg_len
            for c in range(C):
                nset = c if self.neg_sample_by_cat else n_sets_global
                for s in range(P // seg_len):
                    head = S + c if self.head_interaction == "additive" else s * C + c
                    jobs.append(_Job("prior", c, head, nset, pmask(s, seg_len), 1 + c, float(self.prior_loss_weight[c]), s))
        return jobs

    # ------------------------------------------------------------------ helpers
    def _act(self):
        return self.compute_dtype

    def _tc_attention(self):
        """tcgen05 attention: bf16 compute dtype and a head dim the tensor-core kernel tiles (32 / 64)."""
        return self.compute_dtype == torch.bfloat16 and self._dqk in (32, 64) and self.use_tc_attention

    def _phys_heads(self):
        return self.medusa_num_heads if self.medusa_num_layers > 0 else 1

    def _cast_weights(self):
        """Compute-dtype copies of the dense weights (fp32 masters stay in the Parameters).  The bf16 copies are
        persistent "shadows": FusedAdamW rewrites them in its update pass, so a weight is only re-cast here when
        its Parameter changed by other means (load_state_dict, another optimizer, .to()) — detected through the
        tensor version counter and data pointer."""
        act = self._act()
        D = self._hstu_embedding_dim
        w = {}
        if act == torch.float32:
            for i, blk in enumerate(self._hstu._attention_layers):
                w[f"uvqk{i}"], w[f"o{i}"] = blk._uvqk.data, blk._o.weight.data
        else:
            for i, blk in enumerate(self._hstu._attention_layers):
                w[f"uvqk{i}"] = self._shadow_get(blk._uvqk, f"uvqk{i}", (D, 4 * D))
                w[f"o{i}"] = self._shadow_get(blk._o.weight, f"o{i}", (D, D))
        if self.medusa_num_layers > 0 and self.head_interaction != "hierarchical":
            H = self.medusa_num_heads
            dev = self.item_embedding.weight.device
            if act == torch.float32:
                wc = torch.empty((H * D, D), dtype=act, device=dev)
                for h in range(H):
                    wc[h * D:(h + 1) * D].copy_(self.medusa_head[h][0].linear.weight.data)
            else:
                for h in range(H):
                    self._shadow_get(self.medusa_head[h][0].linear.weight, "heads_w", (H * D, D), row0=h * D)
                wc = self._shadow_buf["heads_w"]
            bc = torch.empty((H * D,), dtype=torch.float32, device=dev)
            for h in range(H):
                bc[h * D:(h + 1) * D].copy_(self.medusa_head[h][0].linear.bias.data)
            w["heads_w"], w["heads_b"] = wc, bc
        return w

    def _shadow_get(self, p, name, shape, row0=0):
        """bf16 shadow of Parameter p = rows [row0, row0 + p.shape[0]) of the persistent buffer `name`."""
        buf = self._shadow_buf.get(name)
        if buf is None or buf.device != p.device or buf.dtype != self._act():
            buf = torch.empty(shape, dtype=self._act(), device=p.device)
            self._shadow_buf[name] = buf
            self._shadow_state = {k: v for k, v in self._shadow_state.items() if v[2] != name}
        view = buf[row0:row0 + p.shape[0]] if p.dim() == 2 else buf
        state = (p._version, p.data_ptr(), name)
        if self._shadow_state.get(p) != state:
            L.call("b200rec_cast", p.data_ptr(), p.numel(), view.data_ptr(), L.dt(view), L.stream())
            self._shadow_state[p] = state
            self._shadow_view[p] = view
        return view

    def shadow_of(self, p):
        """The bf16 shadow the optimizer must refresh when it updates p in place (None: p has none)."""
        st = self._shadow_state.get(p)
        if st is None or st[:2] != (p._version, p.data_ptr()):
            self._shadow_state.pop(p, None)      # stale: the next forward re-casts
            return None
        return self._shadow_view[p]

    def state_dict(self, *args, **kwargs):
        if self._table_flush is not None:      # deferred table updates must land before the weights are exported
            self._table_flush()
        return super().state_dict(*args, **kwargs)

    def shadows_stale(self):
        """True if a shadowed Parameter changed since its shadow was written (a captured graph contains no cast)."""
        return any(st[:2] != (p._version, p.data_ptr()) for p, st in self._shadow_state.items())

    def invalidate_shadows(self):
        """Force a re-cast of every weight at the next forward (call after changing parameter memory through
        raw pointers, e.g. restoring a snapshot with copy_ on .data does this automatically)."""
        self._shadow_state.clear()

    def shard_item_table(self, group=None):
        """Keep only the rows `id % world == rank` of the item table on this rank.  Call after
        .to(device).  Lookups / gradient rows then travel by all-to-all (parallel.ShardedTable)."""
        W = parallel.world(group)
        rank = torch.distributed.get_rank(group) if W > 1 else 0
        full = self.item_embedding.weight.data
        self.item_embedding.weight.data = parallel.ShardedTable.shard_of(full, W, rank)
        self.sharded_table = parallel.ShardedTable(self.item_embedding.weight.data, self.item_num, group)
        self.sharded_table.pre_gather = lambda idx: self._table_sync(idx) if self._table_sync is not None else None
        self._table_cache = None
        return self

    def _has_tower(self):
        return isinstance(self.item_id_proj_tower, nn.Linear)

    def _project_rows(self, raw, fp32=False):
        """item_id_proj_tower (hstu.py:414,637,670): fp32 rows [R, D_item] -> fp32 [R, D]; also returns the
        compute-dtype copy of the input rows (the dW operand of the backward).  fp32=True keeps the contraction
        in fp32 (compute_item_all: once per evaluation, feeds the normalised catalogue)."""
        R, Di = raw.shape
        D, act = self._hstu_embedding_dim, (torch.float32 if fp32 else self._act())
        Wt = self.item_id_proj_tower.weight.data                      # [D, D_item] = [N, K]
        if act == torch.float32:
            raw_a, W_a = raw, Wt
        else:
            raw_a = torch.empty((R, Di), dtype=act, device=raw.device)
            L.call("b200rec_cast", raw.data_ptr(), raw.numel(), raw_a.data_ptr(), L.dt(act), L.stream())
            W_a = self._shadow_get(self.item_id_proj_tower.weight, "tower", (D, Di))
        out = torch.empty((R, D), dtype=torch.float32, device=raw.device)
        L.gemm(raw_a, W_a, out, R, D, Di, lda=Di, ldb=Di, ldc=D)
        return out, raw_a, W_a

    def _table_rows(self, items, neg_ids, cache_out=None):
        """Returns (table, item_index [B, LP], neg_index [sets, n_neg], cache info or None): the tensor the
        gather kernels read and the row indices into it.
          replicated table, no tower: the table itself and the ids (info None);
          sharded table: the unique requested rows fetched by all-to-all, positions in that cache;
          projection tower: a row cache holding proj(W[id]) — one row per requested position (static shapes,
          no sync) with a replicated table, one row per unique id on top of the sharded fetch."""
        tower = self._has_tower()
        if self._table_sync is not None and self.sharded_table is None:
            self._table_sync(items)
            self._table_sync(neg_ids)
        if self.sharded_table is None and not tower:
            return self.item_embedding.weight.data, items, neg_ids, None
        B, LP = items.shape
        all_ids = torch.cat([items.reshape(-1), neg_ids.reshape(-1)])
        info = dict(raw_a=None, W_a=None)
        if self.sharded_table is not None:
            uniq, inv = torch.unique(all_ids, return_inverse=True)
            cache = self.sharded_table.fetch(uniq, out=None if tower else cache_out)
            info["ids"] = uniq
        else:
            cache = parallel.cuda_row_gather(self.item_embedding.weight.data, all_ids)
            inv = torch.arange(all_ids.numel(), dtype=torch.int64, device=items.device)
            info["ids"] = all_ids
        if tower:
            cache, info["raw_a"], info["W_a"] = self._project_rows(cache)
        return cache, inv[:B * LP].view(B, LP).contiguous(), inv[B * LP:].view(neg_ids.shape).contiguous(), info

    def prepare_rows(self, items, neg_items, static=False, cache_out=None):
        """Everything of a training step that needs collectives or data-dependent shapes, so it can run
        eagerly in front of a captured graph: the dummy row of static-shape mode, the cross-rank negative
        id all-gather (hstu.py:673,755) and, for a row-sharded table, the all-to-all fetch of the unique
        requested rows.  Returns the table the kernels read and the row indices into it."""
        if static:
            items = torch.cat([items, torch.zeros_like(items[:1])], dim=0)
        items = items.contiguous()
        if self.share_negatives:
            neg_items = parallel.gather_negative_ids(neg_items)      # [W*B, sets, n]
        n_sets = neg_items.shape[1]
        n_neg = neg_items.shape[0] * neg_items.shape[2]
        neg_ids = neg_items.permute(1, 0, 2).contiguous().view(n_sets, n_neg)   # set-major id lists
        W, items_idx, neg_idx, info = self._table_rows(items, neg_ids, cache_out)
        return dict(W=W, items_idx=items_idx, neg_idx=neg_idx, gl_items=items, gl_neg=neg_ids,
                    cached=info is not None, n_rows=(W.shape[0] if info is not None else None), info=info)

    @staticmethod
    def _tokens(valid, force_last=False):
        """Jagged index of a [B, L] validity mask.  force_last adds position L-1 of every sequence as a
        (key-masked) query token so predict() can read it like the reference reads output[:, -1]."""
        B, Lx = valid.shape
        tok_mask = valid.clone()
        if force_last:
            tok_mask[:, -1] = True
        idx = tok_mask.nonzero(as_tuple=False)            # sorted by (b, pos); host sync for T
        tok_b = idx[:, 0].to(torch.int32).contiguous()
        tok_pos = idx[:, 1].to(torch.int32).contiguous()
        seq_off = torch.zeros(B + 1, dtype=torch.int32, device=valid.device)
        seq_off[1:] = tok_mask.sum(1).cumsum(0).to(torch.int32)
        key_valid = valid[idx[:, 0], idx[:, 1]].to(torch.uint8).contiguous()
        return tok_b, tok_pos, seq_off, key_valid, int(idx.shape[0])

    @staticmethod
    def _tokens_static(valid, T):
        """Sync-free jagged index with exactly T token slots.  `valid` is [B+1, L] whose LAST row is the
        all-False dummy row; slots beyond the real tokens become dummy tokens (b = B, pos = 0, key masked)
        that form one trailing dummy sequence.  T must be >= the number of valid positions."""
        Bx, Lx = valid.shape
        assert T <= (Bx - 1) * Lx + Lx, "n_tokens exceeds the number of context positions"
        flat = valid.reshape(-1)
        order = torch.argsort((~flat).to(torch.int8), stable=True)     # valid positions first, in (b, pos) order
        sel = order[:T]
        real = flat[sel]
        tok_b = torch.where(real, torch.div(sel, Lx, rounding_mode="floor"), torch.full_like(sel, Bx - 1))
        tok_pos = torch.where(real, sel % Lx, torch.zeros_like(sel))
        counts = valid.sum(1)
        counts[Bx - 1] = T - counts.sum()
        seq_off = torch.zeros(Bx + 1, dtype=torch.int32, device=valid.device)
        seq_off[1:] = counts.cumsum(0).to(torch.int32)
        return (tok_b.to(torch.int32).contiguous(), tok_pos.to(torch.int32).contiguous(), seq_off,
                real.to(torch.uint8).contiguous(), T)

    # ------------------------------------------------------------------ body (hstu.py:221-328)
    def _body_forward(self, x, w, seq_off, key_valid, B, T, n_pad, max_len, save):
        D, nh, dh = self._hstu_embedding_dim, self._num_heads, self._dqk
        act, dev = self._act(), x.device
        a_dt = L.dt(act)
        saved = []
        st = L.stream()
        drop_p = float(self._linear_dropout_rate) if (self.training and save) else 0.0
        if drop_p > 0:
            if self._rng_step is None or self._rng_step.device != dev:
                self._rng_step = torch.zeros(1, dtype=torch.int64, device=dev)
            L.call("b200rec_counter_add", self._rng_step.data_ptr(), 1, st)   # new keep-masks every forward
        self._drop_p_last = drop_p
        for i in range(self._num_blocks):
            blk = self._hstu._attention_layers[i]
            n = torch.empty((T, D), dtype=act, device=dev)
            mean1 = torch.empty(T, dtype=torch.float32, device=dev)
            rstd1 = torch.empty(T, dtype=torch.float32, device=dev)
            L.call("b200rec_layernorm_fwd", x.data_ptr(), T, D, LN_EPS, n.data_ptr(), a_dt, mean1.data_ptr(),
                   rstd1.data_ptr(), st)
            actv = torch.empty((T, 4 * D), dtype=act, device=dev)
            pre = torch.empty((T, 4 * D), dtype=act, device=dev)
            # uvqk = silu(n @ W): W is [D, 4D] = [K, N] -> MN-major B operand
            L.gemm(n, w[f"uvqk{i}"], actv, T, 4 * D, D, lda=D, ldb=4 * D, b_major=1, ldc=4 * D,
                   epilogue=L.EPI_SILU_DUAL, C2=pre, ldc2=4 * D)
            a = torch.empty((T, D), dtype=torch.float32, device=dev)
            u, v, q, k = actv[:, 0:D], actv[:, D:2 * D], actv[:, 2 * D:3 * D], actv[:, 3 * D:4 * D]
            bias_d = None
            if self.apply_rel_bias:
                rb = blk._rel_attn_bias
                n_b = max(max_len, self._simt_len)
                Nrb = (rb._pos_w.numel() + 1) // 2                      # module length N: pos_w has 2N - 1 entries
                bias_d = torch.zeros(n_b, dtype=torch.float32, device=dev)
                dd = torch.arange(min(n_b, Nrb), device=dev)
                bias_d[:dd.numel()] = rb._pos_w.data[Nrb - 1 - dd] + rb._ts_w.data[0]
                L.call("b200rec_hstu_attn_bias_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), 4 * D, a_dt,
                       seq_off.data_ptr(), key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad, n_b, bias_d.data_ptr(),
                       a.data_ptr(), st)
            elif self._tc_attention() and max_len <= 64:
                L.call("b200rec_hstu_attn_seq_fwd", actv.data_ptr(), 4 * D, seq_off.data_ptr(), key_valid.data_ptr(), B,
                       T, nh, dh, 1.0 / n_pad, max_len, a.data_ptr(), st)
            elif self._tc_attention():
                L.call("b200rec_hstu_attn_tc_fwd", actv.data_ptr(), 4 * D, seq_off.data_ptr(), key_valid.data_ptr(), B,
                       T, nh, dh, 1.0 / n_pad, a.data_ptr(), st)
            else:
                # the launch geometry must cover the LONGEST sequence: in static-token mode the trailing dummy
                # sequence can be longer than the model's max_seq_length (self._simt_len bounds it by T)
                L.call("b200rec_hstu_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), 4 * D, a_dt,
                       seq_off.data_ptr(), key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad,
                       max(max_len, self._simt_len), a.data_ptr(), st)
            oin = torch.empty((T, D), dtype=act, device=dev)
            mean2 = torch.empty(T, dtype=torch.float32, device=dev)
            rstd2 = torch.empty(T, dtype=torch.float32, device=dev)
            L.call("b200rec_gate_ln_fwd", u.data_ptr(), 4 * D, a.data_ptr(), T, D, LN_EPS, oin.data_ptr(), a_dt,
                   mean2.data_ptr(), rstd2.data_ptr(), drop_p, self.dropout_seed, i, L.ptr(self._rng_step), st)
            x_next = torch.empty((T, D), dtype=torch.float32, device=dev)
            L.gemm(oin, w[f"o{i}"], x_next, T, D, D, lda=D, ldb=D, ldc=D, epilogue=L.EPI_BIAS_RESID,
                   bias=blk._o.bias.data, resid=x, ldr=D)
            if save:
                saved.append((x, mean1, rstd1, n, actv, pre, a, mean2, rstd2, oin, bias_d))
            x = x_next
        return x, saved

    def _body_backward(self, dx, saved, w, seq_off, key_valid, B, T, n_pad, max_len, grads):
        D, nh, dh = self._hstu_embedding_dim, self._num_heads, self._dqk
        act, dev = self._act(), dx.device
        a_dt = L.dt(act)
        st = L.stream()
        drop_p = self._drop_p_last
        # weight gradients are off the critical path: they are collected and run as two grouped launches
        # (all layers' dW_o, all layers' dW_uvqk) after the chain, where 16 small problems fill the machine
        dwo_jobs, dwu_jobs = [], []
        dxb = None   # act-dtype copy of dx: written by the previous block's LayerNorm backward
        for i in reversed(range(self._num_blocks)):
            blk = self._hstu._attention_layers[i]
            x, mean1, rstd1, n, actv, pre, a, mean2, rstd2, oin, bias_d = saved[i]
            if act == torch.float32:
                dxb = dx
            elif dxb is None:
                dxb = torch.empty((T, D), dtype=act, device=dev)
                L.call("b200rec_cast", dx.data_ptr(), dx.numel(), dxb.data_ptr(), a_dt, st)
            # d_oin = dx @ W_o   (W_o [Dout, Din] = [K, N] -> MN-major B)
            d_oin = torch.empty((T, D), dtype=act, device=dev)
            L.gemm(dxb, w[f"o{i}"], d_oin, T, D, D, lda=D, ldb=D, b_major=1, ldc=D)
            # dW_o[Dout, Din] = dx^T @ oin  (both operands MN-major, K = T)
            dWo = torch.empty((D, D), dtype=torch.float32, device=dev)
            dwo_jobs.append((dxb, oin, dWo))
            dbo = torch.empty(D, dtype=torch.float32, device=dev)
            L.colsum(dx, T, D, D, dbo)
            d_pre = torch.empty((T, 4 * D), dtype=act, device=dev)
            da = torch.empty((T, D), dtype=act, device=dev)
            L.call("b200rec_gate_ln_bwd", d_oin.data_ptr(), actv.data_ptr(), pre.data_ptr(), 4 * D, a.data_ptr(),
                   mean2.data_ptr(), rstd2.data_ptr(), T, D, d_pre.data_ptr(), da.data_ptr(), a_dt, drop_p,
                   self.dropout_seed, i, L.ptr(self._rng_step), st)
            sl = lambda t, j: t[:, j * D:(j + 1) * D]
            if bias_d is not None:
                n_b = bias_d.numel()
                nws = L.lib().b200rec_hstu_attn_bias_ws_floats(B, nh, n_b)
                part = torch.zeros((nws // n_b, n_b), dtype=torch.float32, device=dev)
                L.call("b200rec_hstu_attn_bias_bwd", sl(actv, 2).data_ptr(), sl(actv, 3).data_ptr(), sl(actv, 1).data_ptr(),
                       sl(pre, 2).data_ptr(), sl(pre, 3).data_ptr(), sl(pre, 1).data_ptr(), 4 * D, a_dt,
                       seq_off.data_ptr(), key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad, n_b, bias_d.data_ptr(),
                       da.data_ptr(), sl(d_pre, 2).data_ptr(), sl(d_pre, 3).data_ptr(), sl(d_pre, 1).data_ptr(),
                       part.data_ptr(), st)
                dbias = torch.empty(n_b, dtype=torch.float32, device=dev)
                L.colsum(part, part.shape[0], n_b, n_b, dbias)
                rb = blk._rel_attn_bias
                Nrb = (rb._pos_w.numel() + 1) // 2
                dpw = torch.zeros_like(rb._pos_w.data)
                dd = torch.arange(min(n_b, Nrb), device=dev)
                dpw[Nrb - 1 - dd] = dbias[:dd.numel()]
                dts = torch.zeros_like(rb._ts_w.data)
                dts[0] = dbias[:dd.numel()].sum()
                grads[rb._pos_w], grads[rb._ts_w] = dpw, dts
            elif self._tc_attention() and max_len <= 64:
                L.call("b200rec_hstu_attn_seq_bwd", actv.data_ptr(), pre.data_ptr(), 4 * D, seq_off.data_ptr(),
                       key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad, max_len, da.data_ptr(), d_pre.data_ptr(), st)
            elif self._tc_attention():
                L.call("b200rec_hstu_attn_tc_bwd", actv.data_ptr(), pre.data_ptr(), 4 * D, seq_off.data_ptr(),
                       key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad, da.data_ptr(), d_pre.data_ptr(), st)
            else:
              L.call("b200rec_hstu_attn_bwd", sl(actv, 2).data_ptr(), sl(actv, 3).data_ptr(), sl(actv, 1).data_ptr(),
                   sl(pre, 2).data_ptr(), sl(pre, 3).data_ptr(), sl(pre, 1).data_ptr(), 4 * D, a_dt,
                   seq_off.data_ptr(), key_valid.data_ptr(), B, T, nh, dh, 1.0 / n_pad,
                   max(max_len, self._simt_len), da.data_ptr(),
                   sl(d_pre, 2).data_ptr(), sl(d_pre, 3).data_ptr(), sl(d_pre, 1).data_ptr(), st)
            # dW_uvqk[D, 4D] = n^T @ d_pre  (both MN-major, K = T)
            dWu = torch.empty((D, 4 * D), dtype=torch.float32, device=dev)
            dwu_jobs.append((n, d_pre, dWu))
            # dn = d_pre @ W_uvqk^T  (W [D, 4D] = [N, K] -> K-major B)
            dn = torch.empty((T, D), dtype=act, device=dev)
            L.gemm(d_pre, w[f"uvqk{i}"], dn, T, D, 4 * D, lda=4 * D, ldb=4 * D, ldc=D)
            dx_prev = torch.empty((T, D), dtype=torch.float32, device=dev)
            dxb = torch.empty((T, D), dtype=act, device=dev) if (act != torch.float32 and i > 0) else None
            L.call("b200rec_layernorm_bwd", dn.data_ptr(), a_dt, D, x.data_ptr(), mean1.data_ptr(), rstd1.data_ptr(),
                   T, D, dx.data_ptr(), dx_prev.data_ptr(), L.ptr(dxb), st)
            grads[blk._uvqk] = dWu
            grads[blk._o.weight] = dWo
            grads[blk._o.bias] = dbo
            dx = dx_prev
            saved[i] = None
        L.gemm_grouped(dwo_jobs, D, D, T, lda=D, a_major=1, ldb=D, b_major=1, ldc=D)
        L.gemm_grouped(dwu_jobs, D, 4 * D, T, lda=D, a_major=1, ldb=4 * D, b_major=1, ldc=4 * D)
        return dx

    # ------------------------------------------------------------------ heads (hstu.py:652-667)
    def _heads_forward(self, y, w, rows):
        """y fp32 [rows, D] -> hd fp32 [rows, Hx, D], z (pre-activation, act) or None, yb (act copy)."""
        D, Hx = self._hstu_embedding_dim, self._phys_heads()
        act, dev = self._act(), y.device
        self._heads_upper = []
        if self.medusa_num_layers == 0:
            return y.view(rows, 1, D), None, None
        if self.head_interaction == "hierarchical":
            return self._hier_forward(y, rows), None, None
        if act == torch.float32:
            yb = y
        else:
            yb = torch.empty((rows, D), dtype=act, device=dev)
            L.call("b200rec_cast", y.data_ptr(), y.numel(), yb.data_ptr(), L.dt(act), L.stream())
        hd = torch.empty((rows, Hx, D), dtype=torch.float32, device=dev)
        z = torch.empty((rows, Hx, D), dtype=act, device=dev)
        L.gemm(yb, w["heads_w"], hd, rows, Hx * D, D, lda=D, ldb=D, ldc=Hx * D, epilogue=L.EPI_RESBLOCK,
               bias=w["heads_b"], resid=y, ldr=D, C2=z, ldc2=Hx * D, n_split=D)
        # medusa_num_layers > 1: the SAME ResBlock is applied again (hstu.py:486-493 builds `[ResBlock] * n`, i.e.
        # weight-tied layers): in_l[h] = hd_{l-1}[h], one GEMM per head with the head's own input
        self._heads_upper = []
        for _ in range(1, self.medusa_num_layers):
            if act == torch.float32:
                hin = hd
            else:
                hin = torch.empty((rows, Hx, D), dtype=act, device=dev)
                L.call("b200rec_cast", hd.data_ptr(), hd.numel(), hin.data_ptr(), L.dt(act), L.stream())
            hd2 = torch.empty((rows, Hx, D), dtype=torch.float32, device=dev)
            z2 = torch.empty((rows, Hx, D), dtype=act, device=dev)
            hin2, hd_2, hdo2, z_2 = hin.view(rows, Hx * D), hd.view(rows, Hx * D), hd2.view(rows, Hx * D), z2.view(rows, Hx * D)
            for h in range(Hx):
                sl = slice(h * D, (h + 1) * D)
                L.gemm(hin2[:, sl], w["heads_w"][sl], hdo2[:, sl], rows, D, D, lda=Hx * D, ldb=D, ldc=Hx * D,
                       epilogue=L.EPI_RESBLOCK, bias=w["heads_b"][sl], resid=hd_2[:, sl], ldr=Hx * D, C2=z_2[:, sl],
                       ldc2=Hx * D, n_split=D)
            self._heads_upper.append((hin, z2))
            hd = hd2
        return hd, z, yb

    # ---- hierarchical heads (hstu.py:652-663): head[s*C + c] = seg[c][s](cat[c](y) [+ segment_emb[s]]), every block a
    # chain of ResBlocks [with a LayerNorm in front of each], the category block optionally behind a LN-Linear-SiLU-Linear
    # bottleneck.  A non-shipped variant of the HSTU scripts: one GEMM (ResBlock / bias epilogue) per layer application,
    # recorded on a tape that the backward walks in reverse; gradients of tensors / weights used more than once add up.
    def _hier_operand(self, lin, x):
        """(activation-dtype copy of x, weight operand) of one Linear on the tape."""
        act = self._act()
        if act == torch.float32:
            return x, lin.weight.data
        xa = torch.empty(x.shape, dtype=act, device=x.device)
        L.call("b200rec_cast", x.data_ptr(), x.numel(), xa.data_ptr(), L.dt(act), L.stream())
        return xa, self._shadow_get(lin.weight, f"hier{id(lin)}", tuple(lin.weight.shape))

    def _hier_ln(self, norm, x):
        """nn.LayerNorm with affine parameters: out = LN(x) * weight + bias (statistics by the LayerNorm kernel)."""
        rows, D = x.shape
        dev = x.device
        xhat = torch.empty((rows, D), dtype=torch.float32, device=dev)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        L.call("b200rec_layernorm_fwd", x.data_ptr(), rows, D, float(norm.eps), xhat.data_ptr(), L.F32, mean.data_ptr(),
               rstd.data_ptr(), L.stream())
        out = torch.addcmul(norm.bias.data, xhat, norm.weight.data)
        self._hier_tape.append(("ln", norm, x, xhat, mean, rstd, out))
        return out

    def _hier_linear(self, lin, x, silu):
        """out = [silu](x @ W^T + b) without a residual (the bottleneck layers, hstu.py:456-460)."""
        rows, Din = x.shape
        Dout = lin.weight.shape[0]
        act, dev = self._act(), x.device
        xa, Wa = self._hier_operand(lin, x)
        out = torch.empty((rows, Dout), dtype=torch.float32, device=dev)
        z = None
        if silu:
            z = torch.empty((rows, Dout), dtype=act, device=dev)
            zero = torch.zeros((rows, Dout), dtype=torch.float32, device=dev)       # RESBLOCK epilogue with a zero residual
            L.gemm(xa, Wa, out, rows, Dout, Din, lda=Din, ldb=Din, ldc=Dout, epilogue=L.EPI_RESBLOCK, bias=lin.bias.data,
                   resid=zero, ldr=Dout, C2=z, ldc2=Dout)      # n_split = 0: no column wrap (Dout need not be a multiple of 32)
        else:
            L.gemm(xa, Wa, out, rows, Dout, Din, lda=Din, ldb=Din, ldc=Dout, epilogue=L.EPI_BIAS_RESID, bias=lin.bias.data)
        self._hier_tape.append(("lin", lin, x, xa, Wa, z, out))
        return out

    def _hier_apply(self, blk, x, out=None, ld_out=None, head=None):
        """One llm_heads.ResBlock: x = norm(x) when use_norm, then x + silu(linear(x)) (llm_heads.py:37-40)."""
        if blk.use_norm:
            x = self._hier_ln(blk.norm, x)
        lin = blk.linear
        rows, D = x.shape
        act, dev = self._act(), x.device
        xa, Wa = self._hier_operand(lin, x)
        if out is None:
            out, ld_out = torch.empty((rows, D), dtype=torch.float32, device=dev), D
        z = torch.empty((rows, D), dtype=act, device=dev)
        L.gemm(xa, Wa, out, rows, D, D, lda=D, ldb=D, ldc=ld_out, epilogue=L.EPI_RESBLOCK, bias=lin.bias.data, resid=x,
               ldr=D, C2=z, ldc2=D)
        self._hier_tape.append(("res", lin, x, xa, Wa, z, out, head))
        return out

    def _hier_forward(self, y, rows):
        D, S, C = self._hstu_embedding_dim, self.num_segment_head, self.num_prior_head
        H = S * C
        self._hier_tape = []
        hd = torch.empty((rows, H, D), dtype=torch.float32, device=y.device)
        hd2 = hd.view(rows, H * D)
        for c in range(C):
            x = y
            layers = list(self.medusa_cat_head[c])
            if self.cat_bottleneck:
                x = self._hier_ln(layers[0], x)
                x = self._hier_linear(layers[1], x, silu=True)
                x = self._hier_linear(layers[3], x, silu=False)
                layers = layers[4:]
            for blk in layers:
                x = self._hier_apply(blk, x)
            for s_ in range(S):
                xs = x
                if self.use_seg_embed:                                    # hstu.py:657-661: seg_in = cat_emb + segment_emb[s]
                    xs = x + self.segment_emb.weight.data[s_]
                    self._hier_tape.append(("segadd", s_, x, xs))
                blocks = list(self.medusa_seg_head[c][s_])
                for blk in blocks[:-1]:
                    xs = self._hier_apply(blk, xs)
                h = s_ * C + c
                self._hier_apply(blocks[-1], xs, out=hd2[:, h * D:(h + 1) * D], ld_out=H * D, head=h)
        return hd

    def _hier_backward(self, d_hd, tape, y, grads):
        """d_hd fp32 [rows, H, D] -> dy fp32 [rows, D]; fills grads of every parameter on the tape."""
        rows, D = y.shape
        H = self.num_segment_head * self.num_prior_head
        act, a_dt, st, dev = self._act(), L.dt(self._act()), L.stream(), y.device
        d_hd2 = d_hd.view(rows, H * D)
        gmap = {}      # data_ptr of a tape tensor -> accumulated fp32 gradient

        def g_in(x, g):                      # a tensor feeding several ops (y, a category output) collects every branch
            key = x.data_ptr()
            if key in gmap:
                gmap[key].add_(g)
            else:
                gmap[key] = g

        def g_par(p, g):                     # share_seg_weights: one block serves every (category, segment)
            grads[p] = grads[p] + g if p in grads else g

        def linear_bwd(lin, xa, Wa, dz, Din, Dout, acc=None):
            """dz (act dtype) [rows, Dout] -> d_in fp32 [rows, Din] = [acc +] dz @ W; dW = dz^T @ x, db = column sums of dz."""
            d_in = acc if acc is not None else torch.empty((rows, Din), dtype=torch.float32, device=dev)
            L.gemm(dz, Wa, d_in, rows, Din, Dout, lda=Dout, ldb=Din, b_major=1, ldc=Din,    # W [Dout, Din] = [K, N]
                   epilogue=L.EPI_ACCUM if acc is not None else L.EPI_STORE)
            dW = torch.empty((Dout, Din), dtype=torch.float32, device=dev)
            L.gemm(dz, xa, dW, Dout, Din, rows, lda=Dout, a_major=1, ldb=Din, b_major=1, ldc=Din)
            db = torch.empty(Dout, dtype=torch.float32, device=dev)
            L.colsum(dz, rows, Dout, Dout, db)
            g_par(lin.weight, dW)
            g_par(lin.bias, db)
            return d_in

        for op in reversed(tape):
            kind = op[0]
            if kind == "res":
                _, lin, x, xa, Wa, z, out, head = op
                if head is None:
                    d_out = gmap.pop(out.data_ptr())
                else:      # the last block of head `head`: its slice of d_hd
                    d_out = d_hd2[:, head * D:(head + 1) * D].contiguous()
                dz = torch.empty((rows, D), dtype=act, device=dev)
                d_x = torch.empty((rows, D), dtype=torch.float32, device=dev)
                L.call("b200rec_resblock_bwd", d_out.data_ptr(), z.data_ptr(), a_dt, rows, 1, D, dz.data_ptr(),
                       d_x.data_ptr(), st)                                # dz = d_out * silu'(z) ; d_x = d_out (residual)
                g_in(x, linear_bwd(lin, xa, Wa, dz, D, D, acc=d_x))          # d_x += dz @ W
            elif kind == "lin":
                _, lin, x, xa, Wa, z, out = op
                d_out = gmap.pop(out.data_ptr())
                Dout, Din = lin.weight.shape
                if z is not None:                                          # through the SiLU
                    dz = torch.empty((rows, Dout), dtype=act, device=dev)
                    scratch = torch.empty((rows, Dout), dtype=torch.float32, device=dev)
                    L.call("b200rec_resblock_bwd", d_out.data_ptr(), z.data_ptr(), a_dt, rows, 1, Dout, dz.data_ptr(),
                           scratch.data_ptr(), st)
                elif act == torch.float32:
                    dz = d_out
                else:
                    dz = torch.empty((rows, Dout), dtype=act, device=dev)
                    L.call("b200rec_cast", d_out.data_ptr(), d_out.numel(), dz.data_ptr(), a_dt, st)
                g_in(x, linear_bwd(lin, xa, Wa, dz, Din, Dout))
            elif kind == "ln":
                _, norm, x, xhat, mean, rstd, out = op
                d_out = gmap.pop(out.data_ptr())
                Dn = x.shape[1]
                dg = torch.empty(Dn, dtype=torch.float32, device=dev)
                L.colsum(d_out * xhat, rows, Dn, Dn, dg)
                db = torch.empty(Dn, dtype=torch.float32, device=dev)
                L.colsum(d_out, rows, Dn, Dn, db)
                g_par(norm.weight, dg)
                g_par(norm.bias, db)
                d_xhat = d_out * norm.weight.data
                d_x = torch.empty((rows, Dn), dtype=torch.float32, device=dev)
                L.call("b200rec_layernorm_bwd", d_xhat.data_ptr(), L.F32, Dn, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                       rows, Dn, None, d_x.data_ptr(), None, st)
                g_in(x, d_x)
            else:                                                          # "segadd": xs = x + segment_emb[s]
                _, s_, x, xs = op
                d_out = gmap.pop(xs.data_ptr())
                w = self.segment_emb.weight
                if w not in grads:
                    grads[w] = torch.zeros_like(w.data)
                L.colsum(d_out, rows, D, D, grads[w][s_], accumulate=True)
                g_in(x, d_out)
        return gmap.pop(y.data_ptr())

    # ------------------------------------------------------------------ prior switch (hstu.py:512-544, 731-805)
    def _switch_heads(self):
        """(active categories, stacked weight [n_act, D], bias [n_act]) — master_switch uses head 0 only."""
        act_c = [0] if self.master_switch else list(range(self.num_prior_head))
        Wa = torch.cat([self.aux_cat_head[c].weight.data for c in act_c], dim=0).contiguous()
        ba = torch.cat([self.aux_cat_head[c].bias.data for c in act_c], dim=0).contiguous()
        return act_c, Wa, ba

    def _switch_forward(self, y, tok_index, items_idx, table, tags, Breal, LP, need_grad):
        """Aux loss over ALL context positions of the real batch rows (padded positions rebuilt in closed form, see
        csrc/switch.cu).  Returns the weighted loss tensor, the logging scalars and what the backward needs."""
        D, P, Lc = self._hstu_embedding_dim, self.pred_len, self.max_seq_length
        dev, st = y.device, L.stream()
        l0, Ls = (Lc - 1, 1) if self.switch_last_only else (0, Lc)
        R = Breal * Ls
        act_c, Wa, ba = self._switch_heads()
        n_act = len(act_c)
        bo_sum = torch.stack([blk._o.bias.data for blk in self._hstu._attention_layers]).sum(0).contiguous()
        rows = torch.empty((R, D), dtype=torch.float32, device=dev)
        L.call("b200rec_switch_rows", y.data_ptr(), tok_index.data_ptr(), items_idx.data_ptr(), table.data_ptr(),
               self.position_embedding.weight.data_ptr(), bo_sum.data_ptr(), Breal, LP, l0, Ls, D, rows.data_ptr(), st)
        head_cat = torch.tensor(act_c, dtype=torch.int32, device=dev)
        pw = []
        for c in act_c:                                           # hstu.py:787-790
            p_ = max(min(float(self.prior_loss_weight[c]), 1.0 - 1e-6), 1e-6)
            pw.append((1.0 - p_) / p_)
        pos_w = torch.tensor(pw, dtype=torch.float32, device=dev)
        asl = self.use_asym_switch_loss
        # BCE: mean over all B x Ls elements; ASL: sum over positions, mean over the batch (layers.py:80-82)
        norm = 1.0 / Breal if asl else 1.0 / R
        logits = torch.empty((R, n_act), dtype=torch.float32, device=dev)
        loss_el = torch.empty((R, n_act), dtype=torch.float32, device=dev)
        correct = torch.empty((R, n_act), dtype=torch.float32, device=dev)
        dlogit = torch.empty((R, n_act), dtype=torch.float32, device=dev)
        tags = tags.contiguous()
        L.call("b200rec_switch_loss", rows.data_ptr(), R, Ls, l0, D, Wa.data_ptr(), ba.data_ptr(), n_act,
               head_cat.data_ptr(), pos_w.data_ptr(), tags.data_ptr(), LP, tags.shape[-1], P, 1 if asl else 0,
               self.gamma_pos, self.gamma_neg, 0.05, 1e-8, self.prior_switch_loss_weight * norm, logits.data_ptr(),
               loss_el.data_ptr(), correct.data_ptr(), dlogit.data_ptr(), st)
        lsum = torch.empty(n_act, dtype=torch.float32, device=dev)
        csum = torch.empty(n_act, dtype=torch.float32, device=dev)
        L.colsum(loss_el, R, n_act, n_act, lsum)
        L.colsum(correct, R, n_act, n_act, csum)
        per_head = lsum * (self.prior_switch_loss_weight * norm)
        logs = {}
        for i, c in enumerate(act_c):
            name = self.int_to_category[c]
            logs[f"head_cat_{name}_acc"] = (csum[i] / R).detach()
            logs[f"head_cat_{name}_loss"] = per_head[i].detach()
        return dict(loss=per_head.sum(), logs=logs, rows=rows, dlogit=dlogit, Wa=Wa, act_c=act_c, Breal=Breal, l0=l0,
                    Ls=Ls, logits=logits)

    def _switch_backward(self, ctx, dy, gscale, grads):
        sw = ctx["switch"]
        D = self._hstu_embedding_dim
        dev, st = dy.device, L.stream()
        Breal, l0, Ls = sw["Breal"], sw["l0"], sw["Ls"]
        R, n_act = Breal * Ls, len(sw["act_c"])
        dW = torch.empty((n_act, D), dtype=torch.float32, device=dev)
        db = torch.empty(n_act, dtype=torch.float32, device=dev)
        pad_rows = None if self.detach_aux_in else torch.empty((R, D), dtype=torch.float32, device=dev)
        L.call("b200rec_switch_bwd", sw["dlogit"].data_ptr(), sw["rows"].data_ptr(), sw["Wa"].data_ptr(), R, Ls, l0,
               ctx["LP"], n_act, D, ctx["tok_index"].data_ptr(), gscale.data_ptr(), dW.data_ptr(), db.data_ptr(),
               None if self.detach_aux_in else dy.data_ptr(), L.ptr(pad_rows), st)
        for i, c in enumerate(sw["act_c"]):
            grads[self.aux_cat_head[c].weight] = dW[i:i + 1]
            grads[self.aux_cat_head[c].bias] = db[i:i + 1]
        return None if self.detach_aux_in else (pad_rows, Breal, l0, Ls)

    # ------------------------------------------------------------------ training (hstu.py:631-872)
    def forward(self, interaction, n_tokens=None, prepared=None):
        """`n_tokens` (optional host int >= number of valid context tokens, e.g. from the collate fn): builds
        the jagged index with static shapes and no host sync, padding with dummy tokens up to n_tokens — the
        form a CUDA graph can capture (see graphed.GraphedTrainStep)."""
        items, neg_items, mask, tags = interaction
        if not items.is_cuda:
            raise L.B200RecError("b200rec.HSTU.forward needs CUDA tensors (there is no CPU path)")
        params = [p for p in self.parameters()]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        loss = _TrainStep.apply(self, (need_grad, n_tokens, prepared), items, neg_items, mask, tags, *params)
        out = defaultdict(float)
        out.update(self._last_logs)
        out["loss"] = loss
        return out

    def _train_forward(self, items, neg_items, mask, tags, need_grad, n_tokens=None, prepared=None):
        dev = items.device
        D, P, Lc = self._hstu_embedding_dim, self.pred_len, self.max_seq_length
        static = n_tokens is not None
        if prepared is None:
            prepared = self.prepare_rows(items, neg_items, static)
        if static:
            # static-shape mode: one all-padding dummy row (index B) owns the dummy tokens; it has no valid
            # position, so it contributes no loss, no gradient and no attention keys.
            mask = torch.cat([mask, torch.zeros_like(mask[:1])], dim=0)
            if tags.numel() > 0:
                tags = torch.cat([tags, torch.zeros_like(tags[:1])], dim=0)
        W, items, neg_ids = prepared["W"], prepared["items_idx"], prepared["neg_idx"]
        gl_items, gl_neg_ids = prepared["gl_items"], prepared["gl_neg"]
        uniq_rows_ids = prepared["cached"] or None
        B, LP = items.shape
        assert LP == Lc + P, f"items must be [B, L+P] = [B, {Lc + P}], got {tuple(items.shape)}"
        act, a_dt, st = self._act(), L.dt(self._act()), L.stream()
        m = mask.bool()
        n_sets, n_neg = neg_ids.shape
        ctx = {}
        # ---- jagged token index (valid context positions only)
        if n_tokens is None:
            tok_b, tok_pos, seq_off, key_valid, T = self._tokens(m[:, :Lc])
        else:
            tok_b, tok_pos, seq_off, key_valid, T = self._tokens_static(m[:, :Lc], int(n_tokens))
        tok_index = torch.full((B * LP,), -1, dtype=torch.int32, device=dev)
        tok_index[tok_b.long() * LP + tok_pos.long()] = torch.arange(T, dtype=torch.int32, device=dev)
        C = self.num_prior_head
        n_col = 1 + (C if self.loss == "prior" else 0)
        tok_ok = torch.zeros((B * LP, n_col), dtype=torch.uint8, device=dev)
        tok_ok[:, 0] = m.reshape(-1)
        if self.loss == "prior":
            tok_ok[:, 1:] = (m.unsqueeze(-1) & tags[:, :, :C].bool()).reshape(B * LP, C)
        w = self._cast_weights()
        # ---- embedding + body
        x = torch.empty((T, D), dtype=torch.float32, device=dev)
        L.call("b200rec_embed_tokens", W.data_ptr(), self.position_embedding.weight.data_ptr(), items.data_ptr(),
               tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, D, x.data_ptr(), st)
        self._simt_len = T if static else 0      # dummy sequence of static-token mode: up to T tokens long
        y, saved = self._body_forward(x, w, seq_off, key_valid, B, T, Lc, Lc, need_grad)
        comi = None
        if self._readout is not None:
            # ComiRec (comirec.py): the queries are interests pooled over the causal context, one per offset
            hd, comi = self._readout.train_forward(self, y, W, items, tok_b, tok_pos, seq_off, B, LP, T,
                                                   B_real=B - (1 if static else 0), need_grad=need_grad)
            z = yb = None
        else:
            hd, z, yb = self._heads_forward(y, w, T)
        Hx = hd.shape[1]
        qhat = torch.empty((T * Hx, D), dtype=act, device=dev)
        qinv = torch.empty(T * Hx, dtype=torch.float32, device=dev)
        L.call("b200rec_gather_l2norm", None, hd.data_ptr(), D, None, T * Hx, qhat.data_ptr(), a_dt, qinv.data_ptr(), st)
        # ---- targets and negatives: gather + L2 normalise (hstu.py:605-606, 670-672)
        flat_items = items.reshape(-1)
        that = torch.empty((B * LP, D), dtype=act, device=dev)
        tinv = torch.empty(B * LP, dtype=torch.float32, device=dev)
        L.call("b200rec_gather_l2norm", W.data_ptr(), None, D, flat_items.data_ptr(), B * LP, that.data_ptr(), a_dt,
               tinv.data_ptr(), st)
        n_words = (n_neg + 31) // 32
        ld_neg = n_words * 32
        used_sets = sorted({j.nset for j in self._jobs})
        nhat, ninv, bits, row_any = {}, {}, {}, {}
        prune_k0 = 64 if (act == torch.bfloat16 and self.use_pruned_filter and D >= 256 and D <= 2048) else 0
        if prune_k0:
            t_aug = torch.empty((B * LP, prune_k0 + 16), dtype=act, device=dev)
            L.call("b200rec_prefix_aug", that.data_ptr(), B * LP, D, prune_k0, t_aug.data_ptr(), st)
        ra_all = torch.zeros((len(used_sets), B * LP + P + 1), dtype=torch.uint8, device=dev)   # one memset for all sets
        # every used negative set in ONE gather + normalise launch (and one prefix launch): the sets are row blocks
        nU = len(used_sets)
        if used_sets == list(range(n_sets)):
            ids_used = neg_ids
        else:
            idx = getattr(self, "_used_sets_idx", None)      # device index, built once (never during a graph capture)
            if idx is None or idx.device != dev or idx.numel() != nU:
                idx = self._used_sets_idx = torch.tensor(used_sets, dtype=torch.int64, device=dev)
            ids_used = neg_ids.index_select(0, idx)
        nh_all = torch.empty((nU * n_neg, D), dtype=act, device=dev)
        ni_all = torch.empty(nU * n_neg, dtype=torch.float32, device=dev)
        L.call("b200rec_gather_l2norm", W.data_ptr(), None, D, ids_used.data_ptr(), nU * n_neg, nh_all.data_ptr(), a_dt,
               ni_all.data_ptr(), st)
        if prune_k0:
            n_aug_all = torch.empty((nU * n_neg, prune_k0 + 16), dtype=act, device=dev)
            L.call("b200rec_prefix_aug", nh_all.data_ptr(), nU * n_neg, D, prune_k0, n_aug_all.data_ptr(), st)
        bits_all = torch.empty((nU, B * LP, n_words), dtype=torch.int32, device=dev)
        for si, s in enumerate(used_sets):
            nh_ = nh_all[si * n_neg:(si + 1) * n_neg]
            ni_ = ni_all[si * n_neg:(si + 1) * n_neg]
            bt = bits_all[si]
            # false-negative filter bits: that @ nhat^T > nce_thres   (hstu.py:613-614)
            ra = ra_all[si]
            if prune_k0:
                # exact, ~16x fewer FLOPs: cos <= <prefix of k0 dims> + |tail_t| |tail_n| (Cauchy-Schwarz) marks the pairs
                # that CAN pass; the full dot product is recomputed only for those (duplicates of a target)
                pass                                   # candidate bits of every set: one grouped launch below
            else:
                L.gemm(that, nh_, bt, B * LP, n_neg, D, lda=D, ldb=D, ldc=n_words, epilogue=L.EPI_GT_BITS,
                       alpha=float(self.nce_thres), C2=ra)
            nhat[s], ninv[s], bits[s], row_any[s] = nh_, ni_, bt, ra
        ctx_neg_all = (nh_all, ni_all, ids_used)
        if prune_k0:
            ka = prune_k0 + 16
            L.gemm_grouped([(t_aug, n_aug_all[si * n_neg:(si + 1) * n_neg], bits[s]) for si, s in enumerate(used_sets)],
                           B * LP, n_neg, ka, lda=ka, ldb=ka, ldc=n_words, epilogue=L.EPI_GT_BITS,
                           alpha=float(self.nce_thres) - 1e-5)
            L.call("b200rec_gt_bits_verify_sets", bits_all.data_ptr(), B * LP, n_words, n_neg, that.data_ptr(),
                   nh_all.data_ptr(), D, float(self.nce_thres), ra_all.data_ptr(), nU, B * LP * n_words, n_neg * D,
                   ra_all.shape[1], st)
        # ---- per-offset token counts -> loss coefficients (hstu.py:704-712, 846-852)
        lam = self.horizon_discount.to(torch.float32)
        keys = []
        for j in self._jobs:
            if (j.col, j.w) not in keys:
                keys.append((j.col, j.w))
        cnt_all = torch.empty(n_col * P, dtype=torch.int32, device=dev)
        coef_all = torch.empty((len(keys), P), dtype=torch.float32, device=dev)
        key_col = (ctypes.c_int32 * len(keys))(*[k[0] for k in keys])
        key_w = (ctypes.c_float * len(keys))(*[float(k[1]) for k in keys])
        L.call("b200rec_nce_coefs", tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P, tok_ok.data_ptr(), n_col,
               ctypes.cast(key_col, ctypes.c_void_p), ctypes.cast(key_w, ctypes.c_void_p), len(keys), lam.data_ptr(),
               cnt_all.data_ptr(),
               coef_all.data_ptr(), st)
        coefs = {k: coef_all[i] for i, k in enumerate(keys)}
        # ---- NCE jobs
        scale = self.logit_scale.data.to(torch.float32)
        job_out = []
        qv = qhat.view(T, Hx * D)
        hqs = [j.head if (self.medusa_num_layers > 0 or self._readout is not None) else 0 for j in self._jobs]
        ihn_beta = float(getattr(self, "_ihn_beta", 0.0))      # comirec.REMI: interest-aware hard negatives (remi.py:198-277)
        fused = act == torch.bfloat16 and self.use_fused_nce and D % 4 == 0 and D <= 2048 and ihn_beta <= 0
        if fused:
            # fused path: no fp32 [T, Nneg] logits in HBM.  positives + row reference -> ONE grouped
            # GEMM whose epilogue writes bf16 softmax numerators E and per-row partial sums -> combine (loss, scalars,
            # row_scale with dL/dlogit = row_scale * E, scaled query copy for the dn GEMM)
            n_parts = L.lib().b200rec_gemm_nce_parts(n_neg)
            J = len(self._jobs)
            # per-job row kernels run as ONE grouped launch each (grid.y = job): buffers are slices of job-major tensors
            pos_cos_all = torch.empty((J, T, P), dtype=torch.float32, device=dev)
            mref_all = torch.empty((J, T), dtype=torch.float32, device=dev)
            thr_all = torch.empty((J, T), dtype=torch.float32, device=dev)
            q_hs = [qv[:, hq * D:(hq + 1) * D] for hq in hqs]
            L.call_grouped("b200rec_nce_pos_ref_grouped", L.job_array(L.NcePosRefJob, [
                dict(q_hat=q_hs[i], p_mask=j.p_mask, tok_ok_col=j.col, pos_cos=pos_cos_all[i], mref=mref_all[i],
                     thr=thr_all[i]) for i, j in enumerate(self._jobs)]),
                Hx * D, that.data_ptr(), D, tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P, tok_ok.data_ptr(), n_col,
                scale.data_ptr(), st)
            Es = [torch.empty((T, ld_neg), dtype=act, device=dev) for _ in range(J)]
            stats_all = torch.empty((J, T, n_parts, 4), dtype=torch.float32, device=dev)
            L.gemm_grouped([(q_hs[i], nhat[j.nset], Es[i]) for i, j in enumerate(self._jobs)],
                           T, n_neg, D, lda=Hx * D, ldb=D, ldc=ld_neg, epilogue=L.EPI_NCE_EXP,
                           nce=[(mref_all[i], thr_all[i], stats_all[i]) for i in range(J)], nce_logit_scale=scale)
            lossv_all = torch.empty((T, J * P), dtype=torch.float32, device=dev)     # job i = columns [i*P, (i+1)*P)
            g0_all = torch.empty((J, T, P), dtype=torch.float32, device=dev)
            dsc_all = torch.empty((J, T, P), dtype=torch.float32, device=dev)
            rank0_all = torch.empty((J, T, P), dtype=torch.int32, device=dev)
            nval_all = torch.empty((J, T, P), dtype=torch.int32, device=dev)
            rscale_all = torch.empty((J, T), dtype=torch.float32, device=dev)
            qss = [torch.empty((T, D), dtype=act, device=dev) if need_grad else None for _ in range(J)]
            L.call_grouped("b200rec_nce_combine_grouped", L.job_array(L.NceCombineJob, [
                dict(stats=stats_all[i], E=Es[i], same_bits=bits[j.nset], row_any=row_any[j.nset], pos_cos=pos_cos_all[i],
                     mref=mref_all[i], q_hat=q_hs[i], coef=coefs[(j.col, j.w)], loss=lossv_all[:, i * P:],
                     g0=g0_all[i], dscale=dsc_all[i], rank0=rank0_all[i], nvalid=nval_all[i], row_scale=rscale_all[i],
                     qs=qss[i]) for i, j in enumerate(self._jobs)]),
                n_parts, ld_neg, n_neg, Hx * D, D, tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P, J * P,
                scale.data_ptr(), D, st)
            per_p_all = torch.empty(J * P, dtype=torch.float32, device=dev)
            L.colsum(lossv_all, T, J * P, J * P, per_p_all)                       # every job's per-offset loss sums
            for i, (j, hq) in enumerate(zip(self._jobs, hqs)):
                job_out.append(dict(job=j, per_p=per_p_all[i * P:(i + 1) * P], g0=g0_all[i], dsc=dsc_all[i],
                                    rank0=rank0_all[i], nval=nval_all[i], hq=hq, G=Es[i] if need_grad else None,
                                    rscale=rscale_all[i], qs=qss[i]))
            job_out[0]["dsc_all"] = dsc_all
            job_sums = per_p_all.view(J, P).sum(dim=1)                             # one launch for all jobs
        else:
            pos_ws = torch.empty(T * P, dtype=torch.float32, device=dev)
            # all (head, negative set) logit GEMMs in one persistent launch (hstu.py:697: one matmul per head)
            all_logits = [torch.empty((T, ld_neg), dtype=torch.float32, device=dev) for _ in self._jobs]
            L.gemm_grouped([(qv[:, hq * D:(hq + 1) * D], nhat[j.nset], lg) for j, hq, lg in zip(self._jobs, hqs, all_logits)],
                           T, n_neg, D, lda=Hx * D, ldb=D, ldc=ld_neg)
            for j, hq, logits in zip(self._jobs, hqs, all_logits):
                q_h = qv[:, hq * D:(hq + 1) * D]
                lossv = torch.empty((T, P), dtype=torch.float32, device=dev)
                g0 = torch.empty((T, P), dtype=torch.float32, device=dev)
                dsc = torch.empty((T, P), dtype=torch.float32, device=dev)
                rank0 = torch.empty((T, P), dtype=torch.int32, device=dev)
                nval = torch.empty((T, P), dtype=torch.int32, device=dev)
                G = torch.empty((T, ld_neg), dtype=act, device=dev) if need_grad else None
                if ihn_beta > 0:
                    L.call("b200rec_nce_ihn_loss_fwd", logits.data_ptr(), ld_neg, n_neg, bits[j.nset].data_ptr(),
                           q_h.data_ptr(), Hx * D, that.data_ptr(), a_dt, D, tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P,
                           j.p_mask, tok_ok.data_ptr(), n_col, j.col, coefs[(j.col, j.w)].data_ptr(), scale.data_ptr(),
                           ihn_beta, lossv.data_ptr(), g0.data_ptr(), dsc.data_ptr(), rank0.data_ptr(), nval.data_ptr(),
                           L.ptr(G), ld_neg, st)
                else:
                    L.call("b200rec_nce_loss_fwd", logits.data_ptr(), ld_neg, n_neg, bits[j.nset].data_ptr(),
                           row_any[j.nset].data_ptr(), pos_ws.data_ptr(), q_h.data_ptr(),
                           Hx * D, that.data_ptr(), a_dt, D, tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P, j.p_mask,
                           tok_ok.data_ptr(), n_col, j.col, coefs[(j.col, j.w)].data_ptr(), scale.data_ptr(),
                           lossv.data_ptr(), g0.data_ptr(), dsc.data_ptr(), rank0.data_ptr(), nval.data_ptr(), L.ptr(G),
                           ld_neg, st)
                per_p = torch.empty(P, dtype=torch.float32, device=dev)
                L.colsum(lossv, T, P, P, per_p)
                job_out.append(dict(job=j, per_p=per_p, g0=g0, dsc=dsc, rank0=rank0, nval=nval, G=G, hq=hq,
                                    rscale=None, qs=None))
            del all_logits
        # ---- total loss + logging scalars (tiny [P] vectors)
        half = 0.5 if (self.loss == "prior" and self.head_interaction == "additive") else 1.0   # hstu.py:870
        total = torch.zeros((), dtype=torch.float32, device=dev)
        logs = {}
        S = self.num_segment_head
        seg_acc = {}
        if not fused:
            job_sums = torch.stack([o["per_p"].sum() for o in job_out])
        total = total + job_sums.sum()
        job_sums = job_sums.detach()
        for i, o in enumerate(job_out):
            j = o["job"]
            if j.part == "nce":
                logs[f"seg_{j.seg}_loss"] = job_sums[i]
            else:
                name = f"head_nce_{self.int_to_category[j.cat]}_loss"
                logs[name] = logs.get(name, 0) + job_sums[i]
                if self.head_interaction != "additive":
                    seg_acc[j.seg] = seg_acc.get(j.seg, 0) + job_sums[i]
        if self.loss == "prior" and self.head_interaction != "additive":
            for s in range(S):
                logs[f"seg_{s}_loss"] = logs.get(f"seg_{s}_loss", 0) + seg_acc.get(s, 0)
        # top-k logging from the rank of the positive (hstu.py:621-629, 720-723, 860-863): the nce part
        # logs first; the first prior head overwrites it when it has offset-0 tokens (device-side select).
        cur = None
        for o in job_out:
            j = o["job"]
            if not (j.p_mask & 1) or (j.part == "prior" and j.cat != 0):
                continue
            tk = torch.empty(7, dtype=torch.float32, device=dev)
            L.call("b200rec_nce_topk_logs", o["rank0"].data_ptr(), o["nval"].data_ptr(), T, P,
                   torch.empty(8, dtype=torch.int64, device=dev).data_ptr(), tk.data_ptr(), st)
            names = ["nce_samples"] + [f"nce_top{k}_acc" for k in (1, 5, 10, 50, 100) if k <= n_neg + 1]
            cur = tk if cur is None else torch.where(tk[0] > 0, tk, cur)
        if cur is not None:
            cur = {n: cur[1 + i] for i, n in enumerate(names)}
        if cur is not None:
            logs.update(cur)
        sw = None
        if self.prior_switch is not None:
            sw = self._switch_forward(y, tok_index, items, W, tags, B - (1 if static else 0), LP, need_grad)
            total = total + sw["loss"]
            logs.update(sw["logs"])
        if comi is not None and comi.get("rr") is not None:          # REMI routing regulariser (remi.py:356-372)
            total = total + comi["rr"] * float(self.lambda_rr)
            logs["rr_loss"] = comi["rr"].detach()
        loss = total * half
        if self._debug is not None:
            self._debug.update(x0=x.clone(), y=y.clone(), hd=hd.clone(), qhat=qhat.clone(), that=that.clone(),
                               per_p=[o["per_p"].clone() for o in job_out])
        if need_grad:
            ctx = dict(B=B, LP=LP, T=T, tok_b=tok_b, tok_pos=tok_pos, seq_off=seq_off, key_valid=key_valid,
                       tok_index=tok_index, w=w, saved=saved, hd=hd, z=z, yb=yb, heads_upper=list(self._heads_upper), hier_tape=getattr(self, "_hier_tape", None), y=y,
                       simt_len=self._simt_len, qhat=qhat, qinv=qinv, that=that,
                       tinv=tinv, nhat=nhat, ninv=ninv, neg_ids=neg_ids, job_out=job_out, scale=scale, half=half,
                       items=items, mask=m, n_neg=n_neg, ld_neg=ld_neg, Hx=Hx, used_sets=used_sets,
                       gl_items=gl_items, gl_neg_ids=gl_neg_ids, uniq_rows_ids=uniq_rows_ids,
                       n_cache_rows=W.shape[0], push=prepared.get("push", True), cache_info=prepared.get("info"),
                       grad_out=prepared.get("grad_out"), switch=sw, table=W, comi=comi, neg_all=ctx_neg_all)
        return loss, logs, ctx

    def _train_backward(self, ctx, gscale):
        """gscale: device fp32 scalar = upstream grad * (0.5 for additive).  Returns {param: grad}."""
        D, P, Lc = self._hstu_embedding_dim, self.pred_len, self.max_seq_length
        B, LP, T, Hx = ctx["B"], ctx["LP"], ctx["T"], ctx["Hx"]
        dev = ctx["that"].device
        act, a_dt, st = self._act(), L.dt(self._act()), L.stream()
        w = ctx["w"]
        n_neg, ld_neg = ctx["n_neg"], ctx["ld_neg"]
        grads = {}
        qhat2 = ctx["qhat"].view(T, Hx * D)
        dqhat = torch.zeros((T, Hx * D), dtype=torch.float32, device=dev)
        dthat = torch.zeros((B * LP, D), dtype=torch.float32, device=dev)
        dnhat = {}
        dscale_sum = torch.zeros((), dtype=torch.float32, device=dev)
        outs = ctx["job_out"]

        def rounds(keys):
            """Jobs with the same output accumulate in job order; jobs of one round have distinct outputs."""
            rs = []
            for idx, k in enumerate(keys):
                for r in rs:
                    if k not in r:
                        r[k] = idx
                        break
                else:
                    rs.append({k: idx})
            return [list(r.values()) for r in rs]

        # dq_hat = G @ nhat   (nhat [n_neg, D] = [K, N] -> MN-major B): one grouped launch per round.  Fused path:
        # G = row_scale * E, the row factor is applied in the GEMM epilogue (dq) / folded into the query copy qs (dn)
        fused = outs[0]["rscale"] is not None
        for r, idxs in enumerate(rounds([o["hq"] for o in outs])):
            L.gemm_grouped([(outs[i]["G"], ctx["nhat"][outs[i]["job"].nset], dqhat[:, outs[i]["hq"] * D:(outs[i]["hq"] + 1) * D])
                            for i in idxs], T, D, n_neg, lda=ld_neg, ldb=D, b_major=1, ldc=Hx * D,
                           epilogue=L.EPI_STORE if r == 0 else L.EPI_ACCUM, alpha_dev=gscale,
                           row_scales=[outs[i]["rscale"] for i in idxs] if fused else None)
        # dn_hat += G^T @ q_hat   (both MN-major, K = T)
        sets_used = ctx["used_sets"]
        dnhat_all = torch.empty((len(sets_used) * n_neg, D), dtype=torch.float32, device=dev)
        for si, s_ in enumerate(sets_used):
            dnhat[s_] = dnhat_all[si * n_neg:(si + 1) * n_neg]
        for r, idxs in enumerate(rounds([o["job"].nset for o in outs])):
            if fused:
                probs = [(outs[i]["G"], outs[i]["qs"], dnhat[outs[i]["job"].nset]) for i in idxs]
                ldq_ = D
            else:
                probs = [(outs[i]["G"], qhat2[:, outs[i]["hq"] * D:(outs[i]["hq"] + 1) * D], dnhat[outs[i]["job"].nset])
                         for i in idxs]
                ldq_ = Hx * D
            L.gemm_grouped(probs, n_neg, D, T, lda=ld_neg, a_major=1, ldb=ldq_, b_major=1, ldc=D,
                           epilogue=L.EPI_STORE if r == 0 else L.EPI_ACCUM, alpha_dev=gscale)
        # positive-logit backward: query side grouped over jobs with distinct head slices (rounds), target side in one launch
        for idxs in rounds([o["hq"] for o in outs]):
            L.call_grouped("b200rec_nce_pos_bwd_q_grouped", L.job_array(L.NcePosBwdJob, [
                dict(g0=outs[i]["g0"], q_hat=None, d_qhat=dqhat[:, outs[i]["hq"] * D:]) for i in idxs]),
                ctx["that"].data_ptr(), a_dt, D, ctx["tok_b"].data_ptr(), ctx["tok_pos"].data_ptr(), T, LP, P,
                ctx["scale"].data_ptr(), gscale.data_ptr(), Hx * D, st)
        L.call_grouped("b200rec_nce_pos_bwd_t_grouped", L.job_array(L.NcePosBwdJob, [
            dict(g0=o["g0"], q_hat=qhat2[:, o["hq"] * D:], d_qhat=None) for o in outs]),
            Hx * D, a_dt, D, ctx["tok_index"].data_ptr(), B, LP, P, ctx["scale"].data_ptr(), gscale.data_ptr(),
            dthat.data_ptr(), st)
        if outs[0].get("dsc_all") is not None:             # fused path: the jobs' [T, P] blocks are one contiguous tensor
            dscale_sum = dscale_sum + outs[0]["dsc_all"].sum()                     # (torch's sum is a fixed-order tree)
        else:
            for o in outs:
                L.call("b200rec_reduce_sum", o["dsc"].data_ptr(), T * P, 1.0, dscale_sum.data_ptr(), 1, st)
        for o in outs:
            o["G"] = o["qs"] = None
        if isinstance(self.logit_scale, nn.Parameter):
            grads[self.logit_scale] = (dscale_sum * gscale).reshape(self.logit_scale.shape)
        # ---- through the L2 normalisation of the heads, the ResBlocks, into dy
        if self._debug is not None:
            self._debug.update(dqhat=dqhat.clone(), dthat=dthat.clone())
        d_hd = torch.empty((T * Hx, D), dtype=torch.float32, device=dev)
        L.call("b200rec_l2norm_bwd", ctx["qhat"].data_ptr(), a_dt, ctx["qinv"].data_ptr(), dqhat.data_ptr(), T * Hx, D,
               d_hd.data_ptr(), 0, st)
        dy = torch.empty((T, D), dtype=torch.float32, device=dev)
        if self._readout is not None:
            ctx["gscale"] = gscale
            dy = self._readout.train_backward(self, d_hd, ctx, grads)
        elif self.medusa_num_layers > 0 and self.head_interaction == "hierarchical":
            dy = self._hier_backward(d_hd, ctx["hier_tape"], ctx["y"], grads)
        elif self.medusa_num_layers > 0:
            dWc = torch.empty((Hx * D, D), dtype=torch.float32, device=dev)
            dbc = torch.empty(Hx * D, dtype=torch.float32, device=dev)
            upper = ctx["heads_upper"]
            for li, (hin, z_l) in enumerate(reversed(upper)):
                # weight-tied upper layers: d_in[h] = d_out[h] + (d_out[h] * silu'(z_l[h])) @ W_h ; dW_h, db_h accumulate
                dz_l = torch.empty((T, Hx * D), dtype=act, device=dev)
                d_in = torch.empty((T * Hx, D), dtype=torch.float32, device=dev)
                L.call("b200rec_resblock_bwd", d_hd.data_ptr(), z_l.data_ptr(), a_dt, T * Hx, 1, D, dz_l.data_ptr(),
                       d_in.data_ptr(), st)
                d_in2, hin2 = d_in.view(T, Hx * D), hin.view(T, Hx * D)
                sls = [slice(h * D, (h + 1) * D) for h in range(Hx)]
                L.gemm_grouped([(dz_l[:, sl], w["heads_w"][sl], d_in2[:, sl]) for sl in sls], T, D, D, lda=Hx * D,
                               ldb=D, b_major=1, ldc=Hx * D, epilogue=L.EPI_ACCUM)
                L.gemm_grouped([(dz_l[:, sl], hin2[:, sl], dWc[sl]) for sl in sls], D, D, T, lda=Hx * D, a_major=1,
                               ldb=Hx * D, b_major=1, ldc=D, epilogue=L.EPI_STORE if li == 0 else L.EPI_ACCUM)
                L.colsum(dz_l, T, Hx * D, Hx * D, dbc, accumulate=li > 0)
                d_hd = d_in
            dz = torch.empty((T, Hx * D), dtype=act, device=dev)
            L.call("b200rec_resblock_bwd", d_hd.data_ptr(), ctx["z"].data_ptr(), a_dt, T, Hx, D, dz.data_ptr(),
                   dy.data_ptr(), st)
            # dy += dz @ Wcat   (Wcat [H*D, D] = [K, N] -> MN-major B)
            L.gemm(dz, w["heads_w"], dy, T, D, Hx * D, lda=Hx * D, ldb=D, b_major=1, ldc=D, epilogue=L.EPI_ACCUM)
            # dWcat[H*D, D] (+)= dz^T @ yb  (both MN-major, K = T)
            L.gemm(dz, ctx["yb"], dWc, Hx * D, D, T, lda=Hx * D, a_major=1, ldb=D, b_major=1, ldc=D,
                   epilogue=L.EPI_ACCUM if upper else L.EPI_STORE)
            L.colsum(dz, T, Hx * D, Hx * D, dbc, accumulate=bool(upper))
            for h in range(Hx):
                lin = self.medusa_head[h][0].linear
                grads[lin.weight] = dWc[h * D:(h + 1) * D]
                grads[lin.bias] = dbc[h * D:(h + 1) * D]
        else:
            L.call("b200rec_resblock_bwd", d_hd.data_ptr(), None, a_dt, T, Hx, D, None, dy.data_ptr(), st)
        # ---- prior-switch aux heads: into dy (valid positions) / pad rows, aux weights
        sw_pad = None
        if ctx.get("switch") is not None:
            sw_pad = self._switch_backward(ctx, dy, gscale, grads)
        # ---- body
        if self._debug is not None:
            self._debug.update(d_hd=d_hd.clone(), dy=dy.clone())
        self._simt_len = ctx["simt_len"]
        dx0 = self._body_backward(dy, ctx["saved"], w, ctx["seq_off"], ctx["key_valid"], B, T, Lc, Lc, grads)
        if self._debug is not None:
            self._debug.update(dx0=dx0.clone())
        # ---- position embedding (rows 0..L-1 used; row L never, hstu.py:380,640-643)
        dpos = torch.zeros_like(self.position_embedding.weight.data)
        L.call("b200rec_pos_emb_grad", dx0.data_ptr(), ctx["tok_index"].data_ptr(), B, LP, Lc, D, dpos.data_ptr(), st)
        grads[self.position_embedding.weight] = dpos
        if sw_pad is not None:
            # padded positions of the aux loss: y_pad = E[item] + P[pos] + sum of the blocks' output biases
            pad_rows, Breal, l0, Ls = sw_pad
            L.call("b200rec_switch_pos_grad", pad_rows.data_ptr(), Breal, Ls, l0, D, dpos.data_ptr(), st)
            dbo_x = torch.empty(D, dtype=torch.float32, device=dev)
            L.colsum(pad_rows, Breal * Ls, D, D, dbo_x)
            for blk in self._hstu._attention_layers:
                grads[blk._o.bias] = grads[blk._o.bias] + dbo_x
        # ---- item embedding: concatenate gradient rows + ids, one sorted-segment reduction
        # `items` / `neg_ids` index the table the kernels read (global ids, or positions in the fetched-row
        # cache of a sharded table); keys <= 0 carry no gradient (padding_idx 0, masked positions).
        items, m = ctx["items"], ctx["mask"]
        sharded = ctx["uniq_rows_ids"] is not None
        shift = 1 if sharded else 0          # cache position 0 is a real row: shift keys by one
        sets = ctx["used_sets"]
        n_sw = 0 if sw_pad is None else sw_pad[1] * sw_pad[3]
        n_rows = T + B * LP + len(sets) * n_neg + n_sw
        rows = torch.empty((n_rows, D), dtype=torch.float32, device=dev)
        ids = torch.empty(n_rows, dtype=torch.int64, device=dev)
        rows[:T].copy_(dx0)
        tok_flat = ctx["tok_b"].long() * LP + ctx["tok_pos"].long()
        ids[:T] = items.reshape(-1)[tok_flat] + shift
        L.call("b200rec_l2norm_bwd", ctx["that"].data_ptr(), a_dt, ctx["tinv"].data_ptr(), dthat.data_ptr(), B * LP, D,
               rows[T:].data_ptr(), 0, st)
        tgt_ids = torch.where(m, items + shift, torch.zeros_like(items))
        tgt_ids[:, 0] = 0                        # position 0 is never a target
        ids[T:T + B * LP] = tgt_ids.reshape(-1)
        off = T + B * LP
        nh_all, ni_all, ids_used = ctx["neg_all"]
        L.call("b200rec_l2norm_bwd", nh_all.data_ptr(), a_dt, ni_all.data_ptr(), dnhat_all.data_ptr(), len(sets) * n_neg, D,
               rows[off:].data_ptr(), 0, st)                       # every set in one launch (row blocks in set order)
        ids[off:off + len(sets) * n_neg] = ids_used.reshape(-1) + shift
        off += len(sets) * n_neg
        sw_gl = []
        if sw_pad is not None:
            # gradient rows of the aux loss at padded context positions -> the pad items' table rows (zero rows, and id 0
            # keys, at valid positions: they were routed into dy)
            pad_rows, Breal, l0, Ls = sw_pad
            rows[off:off + n_sw].copy_(pad_rows)
            pos_valid = ctx["tok_index"].view(B, LP)[:Breal, l0:l0 + Ls] >= 0
            sw_items = items[:Breal, l0:l0 + Ls]
            ids[off:off + n_sw] = torch.where(pos_valid, torch.zeros_like(sw_items), sw_items + shift).reshape(-1)
            sw_gl = [torch.where(pos_valid, torch.zeros_like(sw_items), ctx["gl_items"][:Breal, l0:l0 + Ls]).reshape(-1)]
            off += n_sw
        if sharded:                              # padding id 0 never gets a gradient
            gl = torch.cat([ctx["gl_items"].reshape(-1)[tok_flat], ctx["gl_items"].reshape(-1)] +
                           [ctx["gl_neg_ids"][s] for s in sets] + sw_gl)
            ids = torch.where(gl == 0, torch.zeros_like(ids), ids)
        uniq_ids, uniq_rows, n_uniq = parallel.cuda_segment_reduce(ids, rows)
        if sharded:
            # one gradient row per cache row (fetched / projected rows)
            U = ctx["n_cache_rows"]
            g = ctx.get("grad_out")              # caller-owned buffer (peer-mapped by the sharded step), else a fresh one
            if g is None:
                g = torch.zeros((U, D), dtype=torch.float32, device=dev)
            else:
                assert tuple(g.shape) == (U, D) and g.dtype == torch.float32
                g.zero_()
            L.call("b200rec_rows_to_dense", (uniq_ids - 1).contiguous().data_ptr(), uniq_rows.data_ptr(),
                   n_uniq.data_ptr(), n_rows, D, g.data_ptr(), 0, st)
            info = ctx["cache_info"] or {}
            if info.get("raw_a") is not None:
                # projection tower backward: d_raw = g @ W (W [D, D_item] = [K, N] -> MN-major B),
                # dW[D, D_item] = g^T @ raw (both MN-major, K = cache rows)
                Di = self._item_embedding_dim
                if act == torch.float32:
                    g_a = g
                else:
                    g_a = torch.empty((U, D), dtype=act, device=dev)
                    L.call("b200rec_cast", g.data_ptr(), g.numel(), g_a.data_ptr(), a_dt, st)
                d_raw = torch.empty((U, Di), dtype=torch.float32, device=dev)
                L.gemm(g_a, info["W_a"], d_raw, U, Di, D, lda=D, ldb=Di, b_major=1, ldc=Di)
                dWt = torch.empty((D, Di), dtype=torch.float32, device=dev)
                L.gemm(g_a, info["raw_a"], dWt, D, Di, U, lda=D, a_major=1, ldb=Di, b_major=1, ldc=Di)
                grads[self.item_id_proj_tower.weight] = dWt
                g = d_raw
            self.cache_grad = g
            if self.sharded_table is not None:
                if ctx["push"]:
                    self.emb_grad = self.sharded_table.push_grads(g, scale=1.0 / self.sharded_table.W)
                return grads
            # replicated table behind a tower: cache row -> item id, untouched cache rows carry no gradient
            touched = torch.zeros(U + 1, dtype=torch.bool, device=dev)
            k = n_uniq.to(torch.int64)
            sel = torch.arange(uniq_ids.numel(), device=dev) < k
            touched[torch.where(sel, uniq_ids, torch.zeros_like(uniq_ids))] = True      # keys are cache row + 1
            ids_g = torch.where(touched[1:], info["ids"], torch.zeros_like(info["ids"]))
            uniq_ids, uniq_rows, n_uniq = parallel.cuda_segment_reduce(ids_g, g)
            n_rows = U
        self.emb_grad = (uniq_ids, uniq_rows, n_uniq)
        if not self.sparse_embedding_grad:
            dense = torch.zeros_like(self.item_embedding.weight.data)
            L.call("b200rec_rows_to_dense", uniq_ids.data_ptr(), uniq_rows.data_ptr(), n_uniq.data_ptr(), n_rows,
                   uniq_rows.shape[1], dense.data_ptr(), 0, st)
            grads[self.item_embedding.weight] = dense
        return grads

    # ------------------------------------------------------------------ eval (hstu.py:874-1021)
    @torch.no_grad()
    def compute_item_all(self):
        """hstu.py:1018-1021: L2-normalised (projected) item table, fp32 [N, D]."""
        if self._table_flush is not None:
            self._table_flush()
        self._table_cache = None        # the cached compute-dtype catalogue of predict() belongs to the previous table
        W = self.item_embedding.weight.data
        N = W.shape[0]
        D = self._hstu_embedding_dim
        out = torch.empty((N, D), dtype=torch.float32, device=W.device)
        inv = torch.empty(N, dtype=torch.float32, device=W.device)
        if not self._has_tower():
            L.call("b200rec_gather_l2norm", None, W.data_ptr(), D, None, N, out.data_ptr(), L.F32, inv.data_ptr(), L.stream())
            return out
        step = 1 << 16                                   # project the table in slabs, normalise in place
        for r0 in range(0, N, step):
            r1 = min(N, r0 + step)
            proj, _, _ = self._project_rows(W[r0:r1], fp32=True)
            L.call("b200rec_gather_l2norm", None, proj.data_ptr(), D, None, r1 - r0, out[r0:r1].data_ptr(), L.F32,
                   inv[r0:r1].data_ptr(), L.stream())
        return out

    @torch.no_grad()
    def user_heads(self, item_seq):
        """L2-normalised decode-head embeddings of the last position: act dtype [B, H, D]."""
        if not item_seq.is_cuda:
            raise L.B200RecError("b200rec.HSTU.predict needs CUDA tensors (there is no CPU path)")
        dev = item_seq.device
        D = self._hstu_embedding_dim
        B, Ls = item_seq.shape
        item_seq = item_seq.contiguous()
        act, a_dt, st = self._act(), L.dt(self._act()), L.stream()
        tok_b, tok_pos, seq_off, key_valid, T = self._tokens(item_seq != 0, force_last=True)
        w = self._cast_weights()
        x = torch.empty((T, D), dtype=torch.float32, device=dev)
        if self._table_sync is not None and self.sharded_table is None:
            self._table_sync(item_seq)
        table, seq_idx = self.item_embedding.weight.data, item_seq
        if self.sharded_table is not None:
            uniq, inv = torch.unique(item_seq.reshape(-1), return_inverse=True)
            table, seq_idx = self.sharded_table.fetch(uniq), inv.view(B, Ls).contiguous()
            if self._has_tower():
                table = self._project_rows(table)[0]
        elif self._has_tower():
            table = self._project_rows(parallel.cuda_row_gather(table, item_seq.reshape(-1)))[0]
            seq_idx = torch.arange(B * Ls, dtype=torch.int64, device=dev).view(B, Ls)
        L.call("b200rec_embed_tokens", table.data_ptr(), self.position_embedding.weight.data_ptr(),
               seq_idx.data_ptr(), tok_b.data_ptr(), tok_pos.data_ptr(), T, Ls, D, x.data_ptr(), st)
        self._simt_len = 0
        y, _ = self._body_forward(x, w, seq_off, key_valid, B, T, Ls, Ls, False)
        last = (seq_off[1:] - 1).long()
        y_last = torch.empty((B, D), dtype=torch.float32, device=dev)
        L.call("b200rec_gather_rows", y.data_ptr(), D, last.data_ptr(), B, y_last.data_ptr(), L.F32, st)
        self._last_user_y = y_last                 # body output of the last position (prior-switch heads read it)
        if self._readout is not None:
            hd = self._readout.predict_heads(self, y, seq_off, B, T)       # [B, K interests, D] of the whole sequence
        else:
            hd, _, _ = self._heads_forward(y_last, w, B)
        H = self.medusa_num_heads
        if hd.shape[1] != H:                       # identity heads: every head is the body output
            hd = hd.expand(B, H, D).contiguous()
        U = torch.empty((B * H, D), dtype=act, device=dev)
        uinv = torch.empty(B * H, dtype=torch.float32, device=dev)
        L.call("b200rec_gather_l2norm", None, hd.data_ptr(), D, None, B * H, U.data_ptr(), a_dt, uinv.data_ptr(), st)
        return U.view(B, H, D)

    def _table_hat(self, all_item_feature):
        """predict() re-normalises the table (hstu.py:974-975); cache the compute-dtype copy."""
        key = (all_item_feature.data_ptr(), all_item_feature._version, tuple(all_item_feature.shape), self._act())
        if self._table_cache is None or self._table_cache[0] != key:
            N, D = all_item_feature.shape
            feat = all_item_feature.float().contiguous()
            th = torch.empty((N, D), dtype=self._act(), device=feat.device)
            inv = torch.empty(N, dtype=torch.float32, device=feat.device)
            L.call("b200rec_gather_l2norm", None, feat.data_ptr(), D, None, N, th.data_ptr(), L.dt(self._act()),
                   inv.data_ptr(), L.stream())
            self._table_cache = (key, th)
        return self._table_cache[1]

    def eval_masks(self, all_item_tags, target_tags, B, device):
        """head_cat[H] (category whose item tags gate head h, -1 none), item_tag_bits[N] (bit c = tag c),
        head_on[B, H] (prior_given_at_test), per hstu.py:982-999."""
        H, S, C = self.medusa_num_heads, self.num_segment_head, self.num_prior_head
        self._switch_logs = {}
        if self.loss != "prior":
            return None, None, None
        if self.head_interaction == "additive":
            cats = [-1] * S + list(range(C))
        else:
            cats = [h % C for h in range(H)]
        head_cat = torch.tensor(cats, dtype=torch.int32, device=device)
        tagsb = all_item_tags.bool()                                   # [C, N]
        weights = (1 << torch.arange(tagsb.shape[0], device=device, dtype=torch.int64)).unsqueeze(1)
        item_tag_bits = (tagsb.to(torch.int64) * weights).sum(0).to(torch.int32).contiguous()
        head_on = None
        if self.prior_given_at_test:
            on_c = target_tags[:, :self.given_prior_len].bool().any(dim=1)[:, :C]          # [B, C]
            if self.head_interaction == "additive":
                head_on = torch.cat([torch.ones(B, S, dtype=torch.bool, device=device), on_c], dim=1)
            else:
                head_on = on_c.repeat(1, S)
        self._switch_logs = {}
        if self.prior_switch is not None:
            # hstu.py:935-956, 1001-1015: aux logits of the last position; optionally switch prior heads off
            act_c, Wa, ba = self._switch_heads()
            y_last = self._last_user_y
            logit = torch.empty((B, len(act_c)), dtype=torch.float32, device=device)
            L.gemm(y_last, Wa, logit, B, len(act_c), self._hstu_embedding_dim, lda=self._hstu_embedding_dim,
                   ldb=self._hstu_embedding_dim, ldc=len(act_c), epilogue=L.EPI_BIAS_RESID, bias=ba)
            pred = logit >= 0                                                              # [B, n_act]
            for i, c in enumerate(act_c):
                lab = target_tags[:, :, c].sum(dim=-1) > 0
                self._switch_logs[f"head_cat_{self.int_to_category[c]}_num_correct"] = (lab == pred[:, i]).float().sum()
            if self.use_prior_switch_test:
                if self.master_switch:
                    on_c = torch.cat([pred[:, :1], (~pred[:, :1]).expand(B, C - 1)], dim=1)
                else:
                    on_c = pred
                if self.head_interaction == "additive":
                    sw_on = torch.cat([torch.ones(B, S, dtype=torch.bool, device=device), on_c], dim=1)
                else:
                    sw_on = on_c.repeat(1, S)
                head_on = sw_on if head_on is None else (head_on & sw_on)
        if head_on is not None:
            head_on = head_on.to(torch.uint8).contiguous()
        return head_cat, item_tag_bits, head_on

    @torch.no_grad()
    def predict(self, item_seq, time_seq, all_item_feature, all_item_tags, target_tags, save_for_eval=False):
        """Reference-compatible: returns (scores fp32 [B, H, N] with prior masks, logs, user_embs, head_embs)."""
        U = self.user_heads(item_seq)
        B, H, D = U.shape
        table = self._table_hat(all_item_feature)
        N = table.shape[0]
        scores = torch.empty((B * H, N), dtype=torch.float32, device=U.device)
        L.gemm(U.view(B * H, D), table, scores, B * H, N, D, lda=D, ldb=D, ldc=N)
        head_cat, bits, head_on = self.eval_masks(all_item_tags, target_tags, B, U.device)
        if head_cat is not None:
            L.call("b200rec_apply_score_masks", scores.data_ptr(), N, B, H, N, head_cat.data_ptr(), bits.data_ptr(),
                   L.ptr(head_on), L.stream())
        logs = {"num_samples": self.eval_pred_len * B}
        logs.update(self._switch_logs)
        head_embs = U.float().cpu().numpy() if save_for_eval else None
        return scores.view(B, H, N), logs, None, head_embs

    @torch.no_grad()
    def predict_topk(self, item_seq, all_item_feature, all_item_tags, target_tags, history_index=None, K=200,
                     split_mode="combine", user_chunk=None):
        """Fused eval entry: masks (prior / id 0 / history) + cross-head merge + top-K without returning
        the [B, H, N] tensor.  Returns (topk_idx i64 [B,K], topk_val f32 [B,K], topk_head i32 [B,K]).
        With a row-sharded table (`shard_item_table`) `all_item_feature` / `all_item_tags` are this rank's
        rows (`id % W == rank`); every rank scores all ranks' users against its rows and the per-shard
        lists are exchanged and merged (value desc, id asc)."""
        U = self.user_heads(item_seq)
        B, H, D = U.shape
        dev = U.device
        head_cat, bits, head_on = self.eval_masks(all_item_tags, target_tags, B, dev)
        hu = hi = None
        if history_index is not None:
            hu, hi = history_index[0].to(torch.int64), history_index[1].to(torch.int64)
        st_ = self.sharded_table
        Wd, rank = (st_.W, st_.rank) if st_ is not None else (1, 0)
        if Wd > 1:
            import torch.distributed as dist
            g = st_.group

            def gather_cat(t):
                out = [torch.empty_like(t) for _ in range(Wd)]
                dist.all_gather(out, t.contiguous(), group=g)
                return torch.cat(out, dim=0)

            # ranks may hold different numbers of users (last batch of a strided sampler): pad to the largest
            bt = torch.tensor([B], device=dev)
            bts = [torch.zeros_like(bt) for _ in range(Wd)]
            dist.all_gather(bts, bt, group=g)
            Bmax = max(int(b.item()) for b in bts)

            def pad_rows(t, fill=0):
                if t.shape[0] == Bmax:
                    return t
                out = torch.full((Bmax,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)
                out[:t.shape[0]] = t
                return out

            U = gather_cat(pad_rows(U))
            if head_on is not None:
                head_on = gather_cat(pad_rows(head_on))
            # ragged (user, item) pairs: pad to the max count, shift users by rank * Bmax; every rank takes part in
            # the exchange even when it passes no history (collectives must match across ranks)
            n_h = 0 if hu is None else hu.numel()
            cnt = torch.tensor([n_h, 0 if hu is None else 1], device=dev)
            cnts = [torch.zeros_like(cnt) for _ in range(Wd)]
            dist.all_gather(cnts, cnt, group=g)
            if any(int(c[1].item()) for c in cnts):
                mx = max(1, max(int(c[0].item()) for c in cnts))
                pu = torch.full((mx,), -1, dtype=torch.int64, device=dev)
                pi = torch.zeros((mx,), dtype=torch.int64, device=dev)
                if n_h:
                    pu[:n_h] = hu + rank * Bmax
                    pi[:n_h] = hi
                pu, pi = gather_cat(pu), gather_cat(pi)
                hu, hi = pu[pu >= 0], pi[pu >= 0]
        Ball = U.shape[0]
        table = self._table_hat(all_item_feature)
        N = table.shape[0]
        hist_off = hist_items = None
        if hu is not None:
            order = torch.argsort(hu, stable=True)
            hist_items = hi[order].contiguous()
            hist_off = torch.zeros(Ball + 1, dtype=torch.int32, device=dev)
            hist_off[1:] = torch.bincount(hu, minlength=Ball).cumsum(0).to(torch.int32)
        idx = torch.empty((Ball, K), dtype=torch.int64, device=dev)
        val = torch.empty((Ball, K), dtype=torch.float32, device=dev)
        hsrc = torch.empty((Ball, K), dtype=torch.int32, device=dev)
        mode = 1 if (split_mode == "average" and H > 1) else 0
        fused = self._act() == torch.bfloat16 and mode == 0 and H <= 32 and self.use_fused_eval
        ldn = (N + 3) // 4 * 4                     # padded row pitch: 16-byte stores / loads for any shard size
        if fused:
            # scoring GEMM with the fold-heads epilogue: [Ball*hp, D] x [N, D]^T -> (max, argmax head) per item;
            # the [B, H, N] score tensor is never written.
            # H <= 16: item-row formulation (FOLD_ITEMS: rows = items, columns = (user, head); heads padded to a multiple
            # of 4 only, the fold is register-local).  17..32 heads: (user, head) rows padded to 32 (FOLD_HEADS).
            by_items = H <= 16
            if by_items:
                hp = H if H <= 2 else (H + 3) // 4 * 4
            else:
                hp = 32
            Up = torch.zeros((Ball, hp, D), dtype=U.dtype, device=dev)
            Up[:, :H] = U
            on = torch.zeros((Ball, hp), dtype=torch.uint8, device=dev)
            on[:, :H] = head_on if head_on is not None else 1
            cat = None
            if head_cat is not None:
                cat = torch.full((hp,), -1, dtype=torch.int32, device=dev)
                cat[:H] = head_cat
            on_bits = None
            if by_items:
                on_bits = (on.to(torch.int64) << torch.arange(hp, device=dev)).sum(dim=1).to(torch.int32)

            def fold_gemm(b0, b1, n_rows, fval, fhead, ld, stream_args=()):
                """scores of users [b0, b1) against the first n_rows table rows, folded over heads in the epilogue"""
                nb = b1 - b0
                if by_items:
                    L.gemm(table, Up[b0:b1].reshape(nb * hp, D), fval, n_rows, nb * hp, D, lda=D, ldb=D, ldc=ld,
                           epilogue=L.EPI_FOLD_ITEMS, C2=fhead, ldc2=ld,
                           fold_items=(hp, on_bits[b0:b1].contiguous(), cat, bits, rank, Wd) + tuple(stream_args))
                else:
                    L.gemm(Up[b0:b1].reshape(nb * hp, D), table, fval, nb * hp, n_rows, D, lda=D, ldb=D, ldc=ld,
                           epilogue=L.EPI_FOLD_HEADS, C2=fhead, ldc2=ld,
                           fold=(hp, on[b0:b1].reshape(-1).contiguous(), cat, bits, rank, Wd) + tuple(stream_args))

            def materialised(b0, b1):
                """fold epilogue -> (max, arg-max head) per (user, item) in HBM -> two-read select."""
                nb = b1 - b0
                fval = torch.empty((nb, ldn), dtype=torch.float32, device=dev)
                fhead = torch.empty((nb, ldn), dtype=torch.uint8, device=dev)
                fold_gemm(b0, b1, N, fval, fhead, ldn)
                ho = hist_off[b0:b1 + 1].contiguous() if hist_off is not None else None
                L.call("b200rec_topk_select", fval.data_ptr(), fhead.data_ptr(), nb, N, ldn, K, L.ptr(ho), L.ptr(hist_items),
                       rank, Wd, idx[b0:b1].data_ptr(), val[b0:b1].data_ptr(), hsrc[b0:b1].data_ptr(), L.stream())

            # STREAMED path: the table is read once per user batch and nothing of size users x items
            # reaches HBM.  (1) the first N0 rows go through the materialising path: the K-th best masked score there is
            # a lower bound thr[u] of the K-th best overall; (2) one scoring GEMM over ALL rows whose epilogue keeps only
            # folded scores >= thr[u] (~K * N / N0 per user) in per-user candidate lists; (3) a per-user sort of the
            # candidates (history / id 0 dropped) gives the list.  A rare overflow (counts beyond the capacity: a
            # catalogue whose first rows are unrepresentative) re-runs the batch through the materialising path.
            N0 = min(N, max(32768, ((N + 15) // 16 + 255) // 256 * 256))
            cap = 8192
            streamed = self.use_streamed_eval and N >= 4 * N0 and K <= N0 and N < (1 << 27) and user_chunk is None
            if streamed:
                ldn0 = (N0 + 3) // 4 * 4
                fval = torch.empty((Ball, ldn0), dtype=torch.float32, device=dev)
                fhead = torch.empty((Ball, ldn0), dtype=torch.uint8, device=dev)
                fold_gemm(0, Ball, N0, fval, fhead, ldn0)
                L.call("b200rec_topk_select", fval.data_ptr(), fhead.data_ptr(), Ball, N0, ldn0, K, L.ptr(hist_off),
                       L.ptr(hist_items), rank, Wd, idx.data_ptr(), val.data_ptr(), hsrc.data_ptr(), L.stream())
                thr = val[:, K - 1].contiguous()
                cnt = torch.zeros(Ball, dtype=torch.int32, device=dev)
                keys = torch.empty((Ball, cap), dtype=torch.int64, device=dev)
                ovf = torch.zeros(1, dtype=torch.int32, device=dev)
                fold_gemm(0, Ball, N, fval, None, ldn0, (thr, cnt, keys, cap))
                L.call("b200rec_topk_from_candidates", keys.data_ptr(), cnt.data_ptr(), cap, Ball, K, 0,
                       L.ptr(hist_off), L.ptr(hist_items), rank, Wd, idx.data_ptr(), val.data_ptr(), hsrc.data_ptr(),
                       ovf.data_ptr(), L.stream())
                if int(ovf.item()) != 0:                # also the point where eval hands results to the host anyway
                    streamed = False
            if not streamed:
                if user_chunk is None:
                    user_chunk = max(1, min(Ball, int((8 << 30) // max(1, N * 5))))
                for b0 in range(0, Ball, user_chunk):
                    materialised(b0, min(Ball, b0 + user_chunk))
        else:
            if user_chunk is None:
                user_chunk = max(1, min(Ball, int((8 << 30) // max(1, H * N * 4))))
            for b0 in range(0, Ball, user_chunk):
                b1 = min(Ball, b0 + user_chunk)
                nb = b1 - b0
                scores = torch.empty((nb * H, N), dtype=torch.float32, device=dev)
                L.gemm(U[b0:b1].reshape(nb * H, D), table, scores, nb * H, N, D, lda=D, ldb=D, ldc=N)
                ws_bytes = L.lib().b200rec_topk_workspace_bytes(nb, N)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                ho = hist_off[b0:b1 + 1].contiguous() if hist_off is not None else None
                L.call("b200rec_score_mask_topk", scores.data_ptr(), N, nb, H, N, K, L.ptr(head_cat), L.ptr(bits),
                       L.ptr(head_on[b0:b1].contiguous() if head_on is not None else None), L.ptr(ho),
                       L.ptr(hist_items), mode, rank, Wd, idx[b0:b1].data_ptr(), val[b0:b1].data_ptr(),
                       hsrc[b0:b1].data_ptr(), ws.data_ptr(), ws_bytes, L.stream())
        if Wd == 1:
            return idx, val, hsrc
        # exchange: block w of my lists (users of rank w) goes to rank w; I receive every shard's list of my users
        import torch.distributed as dist
        outs = []
        for t in (val, idx, hsrc):
            r = torch.empty_like(t)
            dist.all_to_all_single(r, t.contiguous(), group=st_.group)
            outs.append(r.view(Wd, Bmax, K)[:, :B])
        v, i, h = outs
        return parallel.merge_topk(list(v), list(i), list(h), K)

    def get_attention_mask(self, item_seq, bidirectional=False):
        """hstu.py:1023-1030 (API parity; the kernels take seq_off / key_valid instead)."""
        keep = (item_seq != 0).unsqueeze(1).unsqueeze(2)
        if not bidirectional:
            keep = torch.tril(keep.expand((-1, -1, item_seq.size(-1), -1)))
        return keep


class _TrainStep(torch.autograd.Function):
    """Whole-model autograd node: forward runs the CUDA forward and keeps activations, backward runs
    the hand-written backward and hands one gradient per parameter to autograd."""

    @staticmethod
    def forward(ctx, model, flags, items, neg_items, mask, tags, *params):
        need_grad, n_tokens, prepared = flags
        loss, logs, saved = model._train_forward(items, neg_items, mask, tags, need_grad, n_tokens, prepared)
        ctx.model, ctx.saved, ctx.params = model, saved, params
        ctx.set_materialize_grads(False)
        model._last_logs = logs
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        model, saved = ctx.model, ctx.saved
        if g_loss is None or not saved:
            return (None,) * (6 + len(ctx.params))
        gscale = (g_loss.reshape(()).to(torch.float32) * saved["half"]).contiguous()
        grads = model._train_backward(saved, gscale)
        ctx.saved = None
        out = []
        for p in ctx.params:
            g = grads.get(p)
            out.append(g if (g is not None and p.requires_grad) else None)
        return (None, None, None, None, None, None) + tuple(out)
