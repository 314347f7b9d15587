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
        h = torch.tanh(y @ self.p["attention_net.0.weight"].t() + self.p["attention_net.0.bias"])   # comirec.py:92-96
        return h @ self.p["attention_net.3.weight"].t()                                                # [..., K]

    def interests(self, y, valid):
        """y [B, L, D], valid [B, L] -> u [B, L, K, D]: u[b, l, k] = sum_{l' <= l, valid} softmax_l'(a[b, l', k]) y[b, l']
        (comirec.py:236-268; a fully masked window gives 0 through nan_to_num)."""
        a = self.attention_logits(y)                                                                    # [B, L, K]
        Lq = y.shape[1]
        causal = torch.tril(torch.ones(Lq, Lq, dtype=torch.bool))                                       # [l, l']
        keep = causal[None, :, :] & valid[:, None, :]                                                   # [B, l, l']
        w = a.permute(0, 2, 1)[:, None, :, :].expand(-1, Lq, -1, -1)                                    # [B, l, K, l']
        w = torch.where(keep[:, :, None, :], w, torch.finfo(w.dtype).min)
        w = torch.nan_to_num(F.softmax(w, dim=-1), nan=0.0)
        w = torch.where(keep[:, :, None, :].any(-1, keepdim=True), w, torch.zeros_like(w))              # all-masked window
        return torch.einsum("blkm,bmd->blkd", w, y)

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
        lt, logits = nce_token_loss(cur[tok], tgt[tok], negs, self.tau(), self.thres)                   # :309-314
        p_of = torch.arange(P)[None, :, None].expand(B, P, L)[tok]
        s = torch.zeros(P, dtype=lt.dtype).index_add_(0, p_of, lt)
        c = torch.zeros(P, dtype=lt.dtype).index_add_(0, p_of, torch.ones_like(lt))
        out = {"loss": (self.lam.to(lt.dtype) * (s / c.clamp_min(1.0))).sum()}                          # :321-333
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
