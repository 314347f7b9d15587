"""CPU oracle of ComiRec on the HSTU body.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch restatement of /root/reference/code/REC/model/IDNet/comirec.py (forward :203-349, predict :351-414) over a
reference-keyed state dict; the body, embedding, NCE token loss and logging are the pinned pieces of
oracle/hstu_oracle.py.  Pinned against tests/golden/comirec_*.pt, produced from the LIVE reference class by
tests/golden/make_golden_comirec.py (tests/test_oracle_golden.py::test_comirec_*).

The reference materialises, for every position l, the window of the L positions ending at l (comirec.py:236-247) and
runs the self-attentive pooling over each window; a window holds exactly the positions <= l, so the pooled interests are
causal prefix softmax averages -- written that way here."""
import torch
import torch.nn.functional as F

from oracle.hstu_oracle import OracleHSTU, nce_token_loss, train_topk_logs, l2n


class OracleComiRec(OracleHSTU):
    def __init__(self, cfg, params):
        cfg = dict(cfg)
        cfg.update(loss="nce", num_segment_head=1, num_prior_head=1, head_interaction="multiplicative",
                   medusa_num_layers=0, neg_sample_by_cat=False)
        super().__init__(cfg, params)
        self.K = params["attention_net.3.weight"].shape[0]

    def attention_logits(self, y):
        z = y @ self.p["attention_net.0.weight"].t()                                                  # comirec.py:92-96
        if "attention_net.0.bias" in self.p:                                                          # remi.py:91 (optional)
            z = z + self.p["attention_net.0.bias"]
        return torch.tanh(z) @ self.p["attention_net.3.weight"].t()                                   # [..., K]

    def routing(self, y, valid):
        """The causal routing matrix w [B, l, K, l'] (softmax over the valid positions l' <= l; 0 elsewhere) and its mask."""
        a = self.attention_logits(y)                                                                    # [B, L, K]
        Lq = y.shape[1]
        causal = torch.tril(torch.ones(Lq, Lq, dtype=torch.bool))                                       # [l, l']
        keep = causal[None, :, :] & valid[:, None, :]                                                   # [B, l, l']
        w = a.permute(0, 2, 1)[:, None, :, :].expand(-1, Lq, -1, -1)                                    # [B, l, K, l']
        w = torch.where(keep[:, :, None, :], w, torch.finfo(w.dtype).min)
        w = torch.nan_to_num(F.softmax(w, dim=-1), nan=0.0)
        w = torch.where(keep[:, :, None, :].any(-1, keepdim=True), w, torch.zeros_like(w))              # all-masked window
        return w, keep

    def interests(self, y, valid):
        """y [B, L, D], valid [B, L] -> u [B, L, K, D]: u[b, l, k] = sum_{l' <= l, valid} softmax_l'(a[b, l', k]) y[b, l']
        (comirec.py:236-268; a fully masked window gives 0 through nan_to_num)."""
        w, _ = self.routing(y, valid)
        return torch.einsum("blkm,bmd->blkd", w, y)

    # hooks of the REMI variant (remi.py): an extra loss on the routing matrix, another per-token loss
    def routing_loss(self, y, valid):
        return None

    def token_loss(self, cur, tgt, negs):
        return nce_token_loss(cur, tgt, negs, self.tau(), self.thres)

    def forward(self, interaction):
        items, neg_items, mask, _ = interaction
        L, P = self.L, self.P
        m = mask.bool()
        B = items.shape[0]
        E = self.embed(items)
        x = E[:, :L] + self.p["position_embedding.weight"][:L]                                          # :212-215
        y = self.body(x, m[:, :L])
        u = self.interests(y, m[:, :L])                                                                 # [B, L, K, D]
        widx = torch.arange(L)[None, :] + 1 + torch.arange(P)[:, None]                                  # [P, L]
        tgt = E[:, widx]                                                                                # [B, P, L, D]  :272
        tok = m[:, None, :L] & m[:, widx]                                                               # :273-275
        sim = torch.einsum("blkd,bpld->bplk", u, tgt)                                                   # :283
        best = torch.argmax(F.softmax(sim, dim=-1), dim=-1)                                             # :287-291 hard readout
        bi = torch.arange(B)[:, None, None].expand(-1, P, L)
        li = torch.arange(L)[None, None, :].expand(B, P, -1)
        cur = u[bi, li, best]                                                                           # [B, P, L, D]  :300
        negs = self.negatives(neg_items[:, -1])                                                         # :228-230
        lt, logits = self.token_loss(cur[tok], tgt[tok], negs)                                          # :309-314
        p_of = torch.arange(P)[None, :, None].expand(B, P, L)[tok]
        s = torch.zeros(P, dtype=lt.dtype).index_add_(0, p_of, lt)
        c = torch.zeros(P, dtype=lt.dtype).index_add_(0, p_of, torch.ones_like(lt))
        out = {"loss": (self.lam.to(lt.dtype) * (s / c.clamp_min(1.0))).sum()}                          # :321-333
        rr = self.routing_loss(y, m[:, :L])
        if rr is not None:                                                                              # remi.py:356-372
            out["rr_loss"] = rr[1].detach()
            out["loss"] = out["loss"] + rr[0] * rr[1]
        if (p_of == 0).any():                                                                           # :336-338
            out.update(train_topk_logs(logits[p_of == 0].detach()))
        return out

    def predict(self, item_seq, all_item_feature):
        """comirec.py:351-414: the K interests of the full sequence, cosine against the catalogue -> [B, K, N]."""
        valid = item_seq != 0
        Lq = item_seq.shape[1]
        x = self.embed(item_seq) + self.p["position_embedding.weight"][:Lq]
        y = self.body(x, valid)
        a = self.attention_logits(y).permute(0, 2, 1)                                                   # [B, K, L]
        a = torch.where(valid[:, None, :], a, torch.finfo(a.dtype).min)
        w = torch.nan_to_num(F.softmax(a, dim=-1), nan=0.0)
        heads = l2n(torch.matmul(w, y).float())
        return torch.matmul(heads, l2n(all_item_feature.float()).t())


def ihn_token_loss(q, t, negs, tau, thres, beta):
    """remi.py:198-277 (beta > 0): interest-aware hard negatives by importance sampling in log space,
    loss = logaddexp(s_pos, LSE((beta+1) s_neg) - (LSE(beta s_neg) - log N)) - s_pos with N = ALL negatives of the row
    (filtered ones included, remi.py:244), filtered negatives out of both sums.  Returns (loss [R], logits [R, 1+N])."""
    qh, th = l2n(q), l2n(t)
    pos = (qh * th).sum(-1, keepdim=True) * tau
    neg = (qh @ negs.t()) * tau
    same = (th @ negs.t()) > thres                                                                  # :226-230
    neg = neg.masked_fill(same, float("-inf"))          # reference: finfo.min, (beta+1) * finfo.min == -inf in fp32
    logits = torch.cat([pos, neg], dim=-1)
    n_all = neg.shape[1]
    log_num = torch.logsumexp((beta + 1.0) * neg, dim=1, keepdim=True)                              # :251-252
    log_z = torch.logsumexp(beta * neg, dim=1, keepdim=True) - torch.log(torch.tensor(float(n_all), dtype=neg.dtype))  # :247-257
    log_neg = torch.where(torch.isfinite(log_z), log_num - log_z, torch.full_like(log_num, float("-inf")))   # :261-264
    loss = (torch.logaddexp(pos, log_neg) - pos).squeeze(-1)                                        # :268-271
    return loss, logits


class OracleREMI(OracleComiRec):
    """REMI on the same body (remi.py:14-437): ComiRec + routing regularisation (RR) + interest-aware hard negatives."""

    def __init__(self, cfg, params):
        super().__init__(cfg, params)
        self.lambda_rr = float(cfg.get("lambda_rr", 100.0) if cfg.get("lambda_rr", None) is not None else 100.0)   # :38
        self.beta = float(cfg.get("beta_ihn", 1.0) if cfg.get("beta_ihn", None) is not None else 1.0)              # :40

    def routing_loss(self, y, valid):
        """remi.py:156-196 + 356-372: per (b, l) the squared norm over interests of the variance of the routing weights
        over the valid window positions, divided by D; mean over the valid (b, l)."""
        if self.lambda_rr <= 0:
            return None
        w, keep = self.routing(y, valid)                                                                # [B, l, K, l']
        mk = keep[:, :, None, :].to(w.dtype)
        n = mk.sum(-1, keepdim=True).clamp(min=1.0)
        wm = w * mk
        dev = (wm - wm.sum(-1, keepdim=True) / n) * mk
        var = (dev * dev).sum(-1) / y.shape[-1]                                                         # [B, l, K]  diag(C)
        per_step = (var ** 2).sum(-1)                                                                   # ||diag||^2
        vf = valid.to(w.dtype)
        return self.lambda_rr, (per_step * vf).sum() / vf.sum().clamp(min=1.0)

    def token_loss(self, cur, tgt, negs):
        if self.beta <= 0:
            return nce_token_loss(cur, tgt, negs, self.tau(), self.thres)                               # :236-239
        return ihn_token_loss(cur, tgt, negs, self.tau(), self.thres, self.beta)
