"""CPU oracle for the HSTU multi-head train / eval hot path  (TEST INFRASTRUCTURE).

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm,
written from the behaviour of /root/reference/code/REC (cited per function as
file:line relative to /root/reference/code/REC).  It is the checker for the CUDA
path: only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import it.  Nothing in the
product package imports this module.

Parity pin: tests/test_oracle_vs_reference.py runs this oracle against the
UNMODIFIED reference modules (oracle/ref_harness.py) when /root/reference is
present, and tests/test_oracle_golden.py checks it against fixtures generated
from the live reference by tests/golden/make_golden.py (committed).  The
reference itself ships no tests or golden vectors.

Like the reference this oracle is dense and padded ([B, L, D] + explicit mask):
it is also the "port" CPU baseline timed by bench.py, so it keeps the
reference's op sequence and cost, not the CUDA path's jagged/dedup'ed one.
"""
import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # model/IDNet/hstu.py:177


def cfg_get(cfg, k, d=None):
    v = cfg.get(k) if hasattr(cfg, "get") else None
    return d if v is None else v


def num_heads_total(cfg):
    """model/IDNet/hstu.py:361-366."""
    S, C = cfg["num_segment_head"], cfg["num_prior_head"]
    if cfg["head_interaction"] in ("multiplicative", "hierarchical"):
        return S * C
    if cfg["head_interaction"] == "additive":
        return S + C
    raise ValueError(f"Unknown head_interaction: {cfg['head_interaction']}")


def ln(x):
    """LayerNorm without affine, eps 1e-6 (hstu.py:213-219)."""
    return F.layer_norm(x, [x.shape[-1]], eps=LN_EPS)


def rel_pos_bias(pos_w, ts_w, L):
    """The position part of RelativeBucketedTimeAndPositionBasedBias.forward (hstu.py:99-134) for all-equal timestamps:
    bias[i, j] = pos_w[N - 1 + (j - i)] + ts_w[bucket 0], N = (len(pos_w) + 1) / 2, cropped to [L, L].  The reference
    never applies it; `apply_relative_attention_bias` turns it on in both oracle and CUDA path and is
    pinned against the LIVE module's forward in tests/test_oracle_vs_reference.py."""
    N = (pos_w.numel() + 1) // 2
    i = torch.arange(L)
    return pos_w[N - 1 + (i[None, :] - i[:, None])] + ts_w[0]                 # [L, L]


def hstu_block(x, keep, w_uvqk, w_o, b_o, n_heads, bias=None):
    """One HSTU block (hstu.py:221-290 + 137-160), dropout off.
    x [B,L,D]; keep bool [B,1,L,L] (hstu.py:1023-1028).  bias: optional [L, L] added to q k^T before SiLU."""
    B, L, D = x.shape
    dh = D // n_heads
    z = F.silu(ln(x) @ w_uvqk)                       # :241-245 (no bias)
    u, v, q, k = torch.split(z, [D, D, D, D], dim=-1)  # :248-257 (dv == dqk == D/heads)
    qh = q.reshape(B, L, n_heads, dh).permute(0, 2, 1, 3)
    kh = k.reshape(B, L, n_heads, dh).permute(0, 2, 1, 3)
    vh = v.reshape(B, L, n_heads, dh).permute(0, 2, 1, 3)
    sc = qh @ kh.transpose(-1, -2)
    if bias is not None:
        sc = sc + bias
    a = F.silu(sc) / L                                # :148-153, divides by PADDED length
    a = a * keep.to(a.dtype)                          # :154
    o = (a @ vh).permute(0, 2, 1, 3).reshape(B, L, D)  # :155-159
    return (u * ln(o)) @ w_o.t() + b_o + x            # :277-288


def causal_keep(valid):
    """keep[b,0,i,j] = valid[b,j] and j <= i  (hstu.py:1023-1030)."""
    L = valid.shape[1]
    tri = torch.ones(L, L, dtype=torch.bool, device=valid.device).tril()
    return (valid[:, None, None, :] & tri[None, None]).contiguous()


def l2n(x):
    return x / x.norm(dim=-1, keepdim=True)


def nce_token_loss(q, t, negs, tau, thres):
    """Per-token sampled-softmax loss (hstu.py:600-619 + cross_entropy with label 0).
    q,t [R,D]; negs [Nn,D] already L2-normalised; returns (loss[R], logits[R,1+Nn])."""
    qh, th = l2n(q), l2n(t)
    pos = (qh * th).sum(-1, keepdim=True)
    neg = (qh @ negs.t()) * tau
    same = (th @ negs.t()) > thres                    # :613-614 false-negative filter
    neg = neg.masked_fill(same, float("-inf"))        # reference: finfo.min * scale == -inf in fp32
    logits = torch.cat([pos * tau, neg], dim=-1)      # :616
    loss = torch.logsumexp(logits, dim=-1) - logits[:, 0]
    return loss, logits


def train_topk_logs(logits):
    """hstu.py:621-629.  top-k accuracy == (rank of logit 0) < k on tie-free rows."""
    out = {"nce_samples": torch.isfinite(logits).sum(dim=1).float().mean().detach()}
    rank0 = (logits[:, 1:] > logits[:, :1]).sum(dim=1)
    for k in (1, 5, 10, 50, 100):
        if k > logits.shape[-1]:
            break
        out[f"nce_top{k}_acc"] = (rank0 < k).float().mean().detach()
    return out


class OracleHSTU:
    """Functional restatement of REC.model.IDNet.hstu.HSTU over a reference-keyed
    state dict.  `params` maps reference parameter names to tensors (leaf tensors
    with requires_grad for gradient checks)."""

    def __init__(self, cfg, params, category_counts=None, category_to_int=None):
        self.cfg = cfg
        self.p = params
        self.L = cfg["MAX_ITEM_LIST_LENGTH"]
        self.P = cfg["pred_len"]
        self.D = cfg["hstu_embedding_size"]
        self.n_layers = cfg["n_layers"]
        self.n_heads = cfg["n_heads"]
        self.S = cfg["num_segment_head"]
        self.C = cfg["num_prior_head"]
        self.H = num_heads_total(cfg)
        self.inter = cfg["head_interaction"]
        self.loss = cfg["loss"]
        self.mlayers = cfg["medusa_num_layers"]
        self.by_cat = bool(cfg["neg_sample_by_cat"]) and self.loss == "prior"   # hstu.py:416-418
        self.thres = cfg_get(cfg, "nce_thres", 0.99)                             # :426
        self.seg_len = self.P // self.S if self.mlayers > 0 else self.P           # :428-432
        lam = torch.tensor([cfg["medusa_lambda"] ** i for i in range(self.P)])
        self.lam = lam / lam.sum()                                                # :436-438
        if self.loss == "prior" and cfg["weighted_prior_loss"] and self.mlayers > 0:
            tot = sum(category_counts.values())                                   # :503-510
            w = [0.0] * self.C
            for name, cnt in category_counts.items():
                w[category_to_int[name]] = cnt / tot
            self.prior_w = w
        else:
            self.prior_w = [1.0 / self.C] * self.C
        self.int_to_category = cfg["int_to_category"]
        # hierarchical head options (hstu.py:444-483): LayerNorm inside every ResBlock, a bottleneck MLP in front of the
        # category block, one segment block shared by every (category, segment), a learned per-segment input offset
        self.head_norm = self.inter == "hierarchical" and bool(cfg_get(cfg, "head_norm", False))
        self.cat_bottleneck = self.inter == "hierarchical" and bool(cfg_get(cfg, "cat_bottleneck", False))
        self.seg_embed = self.inter == "hierarchical" and bool(cfg_get(cfg, "segment_embed", False))
        if self.inter == "hierarchical" and bool(cfg_get(cfg, "share_seg_weights", False)) and self.mlayers > 0:
            # hstu.py:473-478 registers ONE segment block under every medusa_seg_head.{c}.{s} name; a state dict saved to
            # disk no longer aliases them, so tie the entries again (the gradient lands on the first name, like
            # named_parameters() reports it)
            for k in list(self.p):
                if k.startswith("medusa_seg_head."):
                    parts = k.split(".")
                    parts[1], parts[2] = "0", "0"
                    self.p[k] = self.p[".".join(parts)]

    # ---- pieces -----------------------------------------------------------------
    def embed(self, ids):
        e = self.p["item_embedding.weight"][ids]                                  # :637
        if "item_id_proj_tower.weight" in self.p:                                 # :414
            e = e @ self.p["item_id_proj_tower.weight"].t()
        return e

    def body(self, x, valid):
        keep = causal_keep(valid)
        for i in range(self.n_layers):                                            # :322-326
            pre = f"_hstu._attention_layers.{i}."
            bias = None
            if cfg_get(self.cfg, "apply_relative_attention_bias", False) and (pre + "_rel_attn_bias._pos_w") in self.p:
                bias = rel_pos_bias(self.p[pre + "_rel_attn_bias._pos_w"], self.p[pre + "_rel_attn_bias._ts_w"], x.shape[1])
            x = hstu_block(x, keep, self.p[pre + "_uvqk"], self.p[pre + "_o.weight"],
                           self.p[pre + "_o.bias"], self.n_heads, bias)
        return x

    def heads(self, y):
        """y [..., D] -> [H, ..., D]; head h = x + silu(W_h x + b_h), weight-tied when
        medusa_num_layers > 1 (hstu.py:486-493, llm_heads.py:26-40)."""
        if self.inter == "hierarchical":                                          # hstu.py:652-663, 915-925
            D = y.shape[-1]

            def chain(x, prefix, first=0):
                for l in range(first, first + self.mlayers):                      # distinct ResBlocks per layer (:460,468)
                    if self.head_norm:                                            # llm_heads.py:37-39: x = norm(x) FIRST,
                        x = F.layer_norm(x, (D,), self.p[f"{prefix}.{l}.norm.weight"],   # the residual is the normed x
                                         self.p[f"{prefix}.{l}.norm.bias"], 1e-5)
                    W, b = self.p[f"{prefix}.{l}.linear.weight"], self.p[f"{prefix}.{l}.linear.bias"]
                    x = x + F.silu(x @ W.t() + b)
                return x

            def cat_block(c):
                x, pre, first = y, f"medusa_cat_head.{c}", 0
                if self.cat_bottleneck:                                           # :455-461 LN, Linear, SiLU, Linear (no residual)
                    x = F.layer_norm(x, (D,), self.p[f"{pre}.0.weight"], self.p[f"{pre}.0.bias"], 1e-5)
                    x = F.silu(x @ self.p[f"{pre}.1.weight"].t() + self.p[f"{pre}.1.bias"])
                    x = x @ self.p[f"{pre}.3.weight"].t() + self.p[f"{pre}.3.bias"]
                    first = 4
                return chain(x, pre, first)

            cat = [cat_block(c) for c in range(self.C)]
            outs = []
            for s in range(self.S):                                               # :656-663 (share_seg_weights: the state
                for c in range(self.C):                                           # dict aliases one block under every key)
                    x = cat[c] + self.p["segment_emb.weight"][s] if self.seg_embed else cat[c]
                    outs.append(chain(x, f"medusa_seg_head.{c}.{s}"))
            return torch.stack(outs, dim=0)                                       # h = s*C + c
        outs = []
        for h in range(self.H):
            z = y
            if self.mlayers > 0:
                W = self.p[f"medusa_head.{h}.0.linear.weight"]
                b = self.p[f"medusa_head.{h}.0.linear.bias"]
                for _ in range(self.mlayers):
                    z = z + F.silu(z @ W.t() + b)
            outs.append(z)
        return torch.stack(outs, dim=0)

    def tau(self):
        return self.p["logit_scale"].clamp(0, math.log(100)).exp()                # :601-603

    def negatives(self, neg_ids):
        return l2n(self.embed(neg_ids)).reshape(-1, self.D)                        # :670-673 (W == 1)

    # ---- training forward (hstu.py:631-872) ---------------------------------------
    def forward(self, interaction):
        items, neg_items, mask, tags = interaction
        L, P, D = self.L, self.P, self.D
        m = mask.bool()
        B = items.shape[0]
        E = self.embed(items)                                                     # [B, L+P, D]
        x = E[:, :L] + self.p["position_embedding.weight"][:L]                    # :640-643
        y = self.body(x, m[:, :L])                                                # :645-646
        hd = self.heads(y)                                                        # [H, B, L, D]
        tau = self.tau()
        out = defaultdict(float)
        # target windows: tgt[b,p,l] = E[b,l+1+p]; token valid iff m[b,l] & m[b,l+1+p]  (:682-688)
        widx = torch.arange(L)[None, :] + 1 + torch.arange(P)[:, None]            # [P, L]
        tgt = E[:, widx]                                                          # [B, P, L, D]
        tok = m[:, None, :L] & m[:, widx]                                         # [B, P, L]
        p_of = torch.arange(P)[None, :, None].expand(B, P, L)
        total = 0.0

        def offset_means(loss_tok, p_tok):
            s = torch.zeros(P, dtype=loss_tok.dtype).index_add_(0, p_tok, loss_tok)
            c = torch.zeros(P, dtype=loss_tok.dtype).index_add_(0, p_tok, torch.ones_like(loss_tok))
            return s / c.clamp_min(1.0)                                           # :704-708

        use_global = (not self.by_cat) or (self.loss == "prior" and self.inter == "additive")
        if use_global:
            neg_global = self.negatives(neg_items[:, -1])                         # :669-673
        if self.loss == "nce" or (self.loss == "prior" and self.inter == "additive"):
            head_for_p = torch.arange(P) // self.seg_len                          # :677
            q = hd[head_for_p].permute(1, 0, 2, 3)                                # [B, P, L, D]
            lt, logits = nce_token_loss(q[tok], tgt[tok], neg_global, tau, self.thres)
            ptok = p_of[tok]
            per_p = self.lam.to(lt.dtype) * offset_means(lt, ptok)                # :711-713
            total = total + per_p.sum()
            seg = per_p.detach().view(self.S, self.seg_len).sum(1)                # :716-718
            for s in range(self.S):
                out[f"seg_{s}_loss"] = seg[s]
            if (ptok == 0).any():                                                 # :721-723
                out.update(train_topk_logs(logits[ptok == 0].detach()))
        if self.loss == "prior" and cfg_get(self.cfg, "prior_switch") == "in" and self.mlayers > 0:
            total = total + self.prior_switch_loss(y, tags, out)                  # :731-805
        if self.loss == "prior":
            seg_len = P if self.inter == "additive" else self.seg_len             # :726-729
            seg_for_p = torch.arange(P) // seg_len
            acc = torch.zeros(P)
            for c in range(self.C):                                               # :748
                name = self.int_to_category[c]
                out[f"head_nce_{name}_loss"] = 0
                negs = self.negatives(neg_items[:, c]) if self.by_cat else neg_global  # :751-755
                tagwin = tags[:, :, c].bool()[:, widx]                            # :808-809
                tk = tok & tagwin                                                 # :813
                if tk.sum() == 0:                                                 # :815-839 guard: contributes 0
                    continue
                if self.inter == "additive":
                    head_for_p = torch.full((P,), self.S + c)                     # :823
                else:
                    head_for_p = seg_for_p * self.C + c                           # :825
                q = hd[head_for_p].permute(1, 0, 2, 3)
                lt, logits = nce_token_loss(q[tk], tgt[tk], negs, tau, self.thres)
                ptok = p_of[tk]
                per_p = self.lam.to(lt.dtype) * self.prior_w[c] * offset_means(lt, ptok)  # :852
                total = total + per_p.sum()
                acc = acc + per_p.detach().float()
                out[f"head_nce_{name}_loss"] = per_p.sum().detach()
                if c == 0 and (ptok == 0).any():                                  # :861-863
                    out.update(train_topk_logs(logits[ptok == 0].detach()))
            if self.inter != "additive":
                seg = acc.view(self.S, self.seg_len).sum(1)                       # :865-868
                for s in range(self.S):
                    out[f"seg_{s}_loss"] += seg[s]
            else:
                total = total / 2                                                 # :870
        out["loss"] = total
        return out

    def prior_switch_loss(self, y, tags, out):
        """Aux category heads on the dense body output (hstu.py:731-805, prior_switch == 'in'): weighted BCE
        (pos_weight = (1 - p_c) / p_c) or AsymmetricLoss (layers.py:16-84) over ALL [B, L] positions."""
        cfg, L, P = self.cfg, self.L, self.P
        w = float(cfg["prior_switch_loss_weight"])
        master = bool(cfg_get(cfg, "master_switch", False))
        last_only = bool(cfg_get(cfg, "switch_last_only", False))
        asl = bool(cfg_get(cfg, "asym_switch_loss", False))
        widx = torch.arange(L)[:, None] + 1 + torch.arange(P)[None, :]            # [L, P] target window of position l
        total = 0.0
        for c in range(self.C):
            if master and c > 0:
                continue
            tgt = tags[:, :, c].bool()[:, widx].any(dim=-1).float()               # [B, L]   (:733-736, 763-766)
            aux_in = y.detach() if cfg_get(cfg, "detach_aux_in", False) else y
            if last_only:
                tgt, aux_in = tgt[:, -1:], aux_in[:, -1:]
            logit = (aux_in @ self.p[f"aux_cat_head.{c}.weight"].t() + self.p[f"aux_cat_head.{c}.bias"]).squeeze(-1)
            if asl:
                gp, gn = float(cfg_get(cfg, "gamma_pos", 4.0)), float(cfg_get(cfg, "gamma_neg", 0.0))
                sg = torch.sigmoid(logit)
                xs_pos, xs_neg = sg, (1 - sg + 0.05).clamp(max=1)
                ls = tgt * torch.log(xs_pos.clamp(min=1e-8)) + (1 - tgt) * torch.log(xs_neg.clamp(min=1e-8))
                if gn > 0 or gp > 0:
                    pt = xs_pos * tgt + xs_neg * (1 - tgt)
                    ls = ls * torch.pow(1 - pt, gp * tgt + gn * (1 - tgt))
                lc = (-ls.sum(dim=-1)).mean()
            else:
                p_ = max(min(float(self.prior_w[c]), 1.0 - 1e-6), 1e-6)
                lc = F.binary_cross_entropy_with_logits(logit, tgt, pos_weight=torch.tensor((1.0 - p_) / p_))
            name = self.int_to_category[c]
            out[f"head_cat_{name}_acc"] = ((logit >= 0).int() == tgt.int()).float().mean().detach()
            out[f"head_cat_{name}_loss"] = (w * lc).detach()
            total = total + w * lc
        return total

    # ---- eval (hstu.py:874-1021) ------------------------------------------------------
    @torch.no_grad()
    def compute_item_all(self):
        return l2n(self.embed(torch.arange(self.p["item_embedding.weight"].shape[0])))

    @torch.no_grad()
    def user_heads(self, item_seq):
        """L2-normalised head embeddings of the last position, [B, H, D]  (:879-966)."""
        L = item_seq.shape[1]
        x = self.embed(item_seq) + self.p["position_embedding.weight"][:L]
        y = self.body(x, item_seq != 0)[:, -1]                                    # :908-913
        self._last_y = y
        return l2n(self.heads(y).permute(1, 0, 2).float())

    @torch.no_grad()
    def predict(self, item_seq, time_seq, all_item_feature, all_item_tags, target_tags, save_for_eval=False):
        U = self.user_heads(item_seq)                                             # [B,H,D]
        T = l2n(all_item_feature.float())                                         # :974-975 (re-normalised)
        scores = U @ T.t()                                                        # :979  [B,H,N]
        S, C = self.S, self.C
        if self.loss == "prior":
            sl = slice(S, None) if self.inter == "additive" else slice(None)
            rep = 1 if self.inter == "additive" else S
            if cfg_get(self.cfg, "prior_given_at_test", False):                   # :983-990
                g = cfg_get(self.cfg, "given_prior_len", self.cfg["eval_pred_len"])
                on = target_tags[:, :g].bool().any(dim=1).repeat(1, rep)          # [B, C*rep]
                scores[:, sl].masked_fill_(~on.unsqueeze(-1), float("-inf"))
            it = all_item_tags.bool().repeat(rep, 1)                              # :994-999  [C*rep, N]
            scores[:, sl].masked_fill_(~it.unsqueeze(0), float("-inf"))
        logs = {"num_samples": self.cfg["eval_pred_len"] * item_seq.shape[0]}     # :933
        if self.loss == "prior" and cfg_get(self.cfg, "prior_switch") == "in" and self.mlayers > 0:
            y_last = self._last_y                                                 # :935-956, 1001-1015
            master = bool(cfg_get(self.cfg, "master_switch", False))
            preds = []
            for c in range(1 if master else C):
                lg = (y_last @ self.p[f"aux_cat_head.{c}.weight"].t() + self.p[f"aux_cat_head.{c}.bias"]).squeeze(-1)
                preds.append(lg >= 0)
                lab = target_tags[:, :, c].sum(dim=-1) > 0
                logs[f"head_cat_{self.int_to_category[c]}_num_correct"] = (lab == preds[c]).float().sum()
            if cfg_get(self.cfg, "use_prior_switch_test", False):
                if master:
                    off = torch.cat([~preds[0][:, None], preds[0][:, None].expand(-1, C - 1)], dim=1)
                else:
                    off = ~torch.stack(preds, dim=1)
                sl = slice(S, None) if self.inter == "additive" else slice(None)
                rep = 1 if self.inter == "additive" else S
                scores[:, sl].masked_fill_(off.repeat(1, rep).unsqueeze(-1), float("-inf"))
        return scores, logs, None, None


def post_mask_scores(scores, history_index=None):
    """trainer/trainer.py:724-726."""
    scores[:, :, 0] = float("-inf")
    if history_index is not None:
        scores[history_index[0], :, history_index[1]] = float("-inf")
    return scores


# ---- Collector restatement (evaluator/collector.py:153-325, rec.topk path) --------------
def _topk_rows(vals, K):
    """Row-wise top-K with the stated tie rule: value desc, then id asc."""
    v = vals.numpy() if isinstance(vals, torch.Tensor) else vals
    N = v.shape[-1]
    flat = v.reshape(-1, N)
    ids = np.empty((flat.shape[0], K), dtype=np.int64)
    for r in range(flat.shape[0]):
        order = np.lexsort((np.arange(N), -flat[r].astype(np.float64)))
        ids[r] = order[:K]
    top = np.take_along_axis(flat, ids, axis=1)
    return top.reshape(v.shape[:-1] + (K,)), ids.reshape(v.shape[:-1] + (K,))


def collect_topk(scores, K, split_mode="combine"):
    """Returns (topk_idx[B,K] i64, values[B,K] f32, head_source[B,K] i64).
    H==1: plain top-K (collector.py:203-205).  'combine': per-head top-K, merged by value
    desc, first occurrence of each id kept, first K (collector.py:241-275).  'average':
    mean over finite heads then top-K (collector.py:227-230).
    Tie rule (the reference's is unspecified): value desc, item id asc, head asc."""
    s = scores.float().numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores, dtype=np.float32)
    B, H, N = s.shape
    if H == 1:
        v, i = _topk_rows(s[:, 0], K)
        return i, v, np.zeros_like(i)
    if split_mode == "average":
        fin = np.isfinite(s)
        avg = np.where(fin, s, 0).sum(1) / (fin.sum(1) + 1e-8)
        v, i = _topk_rows(avg.astype(np.float32), K)
        return i, v, np.zeros_like(i)
    if split_mode != "combine":
        raise ValueError(f"Unknown split_mode: {split_mode}")
    hv, hi = _topk_rows(s, K)                                   # [B,H,K]
    out_i = np.empty((B, K), np.int64)
    out_v = np.empty((B, K), np.float32)
    out_h = np.empty((B, K), np.int64)
    hsrc = np.broadcast_to(np.arange(H)[:, None], (H, K)).reshape(-1)
    for b in range(B):
        fv, fi = hv[b].reshape(-1), hi[b].reshape(-1)
        order = np.lexsort((hsrc, fi, -fv.astype(np.float64)))
        seen, k = set(), 0
        for j in order:
            it = int(fi[j])
            if it in seen:
                continue
            seen.add(it)
            out_i[b, k], out_v[b, k], out_h[b, k] = it, fv[j], hsrc[j]
            k += 1
            if k == K:
                break
        if k < K:
            raise AssertionError("Duplicated elements found in some batch samples")  # collector.py:290-293
    return out_i, out_v, out_h


def hit_matrices(topk_idx, positive_i, metrics_pred_len_list):
    """collector.py:300-316.  Returns {p: i32[B, K+1]} (hit flags | pos_len column).
    pos_len[b,p] = #distinct ids among the p+1 smallest-id targets of the FULL target row
    (reference quirk); hit slices are cumulative [:, 0:p+1] (prev_idx never moves)."""
    topk_idx = np.asarray(topk_idx)
    pos = np.asarray(positive_i)
    srt = np.sort(pos, axis=1)
    first = np.ones_like(srt, dtype=bool)
    first[:, 1:] = srt[:, 1:] != srt[:, :-1]
    pos_len_full = np.cumsum(first, axis=1).astype(np.int32)
    out = {}
    hit = np.zeros(topk_idx.shape, dtype=bool)
    for p in metrics_pred_len_list:
        sl = pos[:, 0:p + 1]
        hit |= (topk_idx[:, :, None] == sl[:, None, :]).any(-1)
        out[p] = np.concatenate([hit.astype(np.int32), pos_len_full[:, p:p + 1]], axis=1)
    return out


def recall_ndcg_sums(rec_topk, topk_list):
    """evaluator/metrics.py:179-238 + base_metric.py:51-81: SUMS over users (the trainer divides)."""
    rec = np.asarray(rec_topk)
    K = rec.shape[1] - 1
    hit = rec[:, :K].astype(bool)
    pos_len = rec[:, K].astype(np.int64)
    recall = np.cumsum(hit, axis=1) / pos_len.reshape(-1, 1)
    ranks = np.arange(1, K + 1, dtype=np.float64)
    disc = 1.0 / np.log2(ranks + 1)
    idcg_len = np.minimum(pos_len, K)
    idcg_all = np.cumsum(disc)
    idcg = np.broadcast_to(idcg_all, hit.shape).copy()
    for r, n in enumerate(idcg_len):
        idcg[r, n:] = idcg[r, n - 1]
    dcg = np.cumsum(np.where(hit, disc, 0.0), axis=1)
    ndcg = dcg / idcg
    res = {}
    for k in topk_list:
        res[f"recall@{k}"] = recall.sum(0)[k - 1]
    for k in topk_list:
        res[f"ndcg@{k}"] = ndcg.sum(0)[k - 1]
    return res


def state_dict_from_module(module, dtype=torch.float32, requires_grad=False):
    """Leaf copies of a module's state dict.  Entries that alias one tensor in the module (share_seg_weights registers one
    block under several names, hstu.py:473-478) stay ONE leaf, so its gradient is the sum over all uses, like autograd's."""
    sd, seen = {}, {}
    for k, v in module.state_dict().items():
        key = (v.data_ptr(), tuple(v.shape), v.dtype) if v.numel() > 0 else None
        if key is not None and key in seen:
            sd[k] = seen[key]
            continue
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            t.requires_grad_(requires_grad)
        sd[k] = t
        if key is not None:
            seen[key] = t
    return sd
