"""ComiRec-SA on the HSTU body: drop-in for REC/model/IDNet/comirec.py `ComiRec(config, dataload)` --
same constructor keys, batch tuple, output keys and parameter names (`position_embedding`, `_hstu.*`,
`attention_net.0.{weight,bias}`, `attention_net.3.weight`, `item_embedding`, `logit_scale`), so a reference state dict
loads.  Everything up to the body output and everything from the sampled-softmax loss on is the HSTU path of hstu.py;
what ComiRec adds sits in between (comirec.py:232-300):

    a = attention_net(y)                       Linear -> tanh -> Linear, [T, K]          (fp32 GEMMs + csrc/comirec.cu)
    u[t, k] = causal softmax pooling of y       one online-softmax pass per (sequence, k)  (b200rec_comi_pool_fwd)
    q[t, p] = u[t, argmax_k <u[t, k], target(t, p)>]   hard readout per offset            (b200rec_comi_select_fwd)

and the P per-offset query sets run through the HSTU NCE machinery as P "heads" (one job per offset), whose gradient
comes back through the select / pooling / attention-net backward kernels into the body.  predict() scores the K
interests of the whole sequence against the catalogue ([B, K, N], comirec.py:351-414).

REMI (remi.py) is the same model with two training-time additions, both built here: the routing regulariser on the
attention-net logits (b200rec_comi_rr) and the interest-aware hard-negative loss (b200rec_nce_ihn_loss_fwd).
Not built: `skip_hstu`, dropout inside attention_net during training (p > 0 raises), item_embedding_size !=
hstu_embedding_size.  Module registration order
differs from the reference (attention_net is created after the body's parameters), so a same-seed init is not
bit-identical; loading a reference state dict is.
"""
import torch
import torch.nn as nn

from . import _lib as L
from .hstu import HSTU, _Job, truncated_normal_


class _Readout(object):
    """The multi-interest readout, called from HSTU._train_forward / _train_backward / user_heads."""

    def __init__(self, K, hidden):
        self.K, self.hidden = K, hidden

    def _logits(self, model, y, T):
        """a = tanh(y W1^T + b1) W2^T -> (h [T, Hd], a [T, K]); fp32 (the SIMT GEMM takes any shape)."""
        D, Hd, K = y.shape[1], self.hidden, self.K
        lin1, lin2 = model.attention_net[0], model.attention_net[3]
        h = torch.empty((T, Hd), dtype=torch.float32, device=y.device)
        if lin1.bias is not None:
            L.gemm(y, lin1.weight.data, h, T, Hd, D, lda=D, ldb=D, ldc=Hd, epilogue=L.EPI_BIAS_RESID, bias=lin1.bias.data)
        else:                                               # REMI: attention_net_bias = False (remi.py:91)
            L.gemm(y, lin1.weight.data, h, T, Hd, D, lda=D, ldb=D, ldc=Hd)
        L.call("b200rec_comi_tanh", h.data_ptr(), h.numel(), L.stream())
        a = torch.empty((T, K), dtype=torch.float32, device=y.device)
        L.gemm(h, lin2.weight.data, a, T, K, Hd, lda=Hd, ldb=Hd, ldc=K)
        return h, a

    def _pool(self, y, a, seq_off, B, T):
        D, K = y.shape[1], self.K
        u = torch.empty((T, K, D), dtype=torch.float32, device=y.device)
        M = torch.empty((T, K), dtype=torch.float32, device=y.device)
        S = torch.empty((T, K), dtype=torch.float32, device=y.device)
        L.call("b200rec_comi_pool_fwd", a.data_ptr(), y.data_ptr(), seq_off.data_ptr(), B, K, D, u.data_ptr(), M.data_ptr(),
               S.data_ptr(), L.stream())
        return u, M, S

    def _routing_reg(self, model, a, seq_off, B_real, T, need_grad):
        """REMI's routing regulariser (remi.py:156-196, 356-372) from the routing logits a [T, K]: (rr, d rr / d a)."""
        K, dev = self.K, a.device
        var2 = torch.zeros((T, K), dtype=torch.float32, device=dev)            # dummy tokens (static mode) stay 0
        da = torch.zeros((T, K), dtype=torch.float32, device=dev) if need_grad else None
        scratch = torch.empty((T, K, 3), dtype=torch.float32, device=dev) if need_grad else None
        L.call("b200rec_comi_rr", a.data_ptr(), seq_off.data_ptr(), B_real, K, model._hstu_embedding_dim, var2.data_ptr(),
               L.ptr(scratch), L.ptr(da), L.stream())
        n_valid = seq_off[B_real].to(torch.float32).clamp(min=1.0)             # device scalar: no host sync
        return var2.sum() / n_valid, da

    def train_forward(self, model, y, W, items, tok_b, tok_pos, seq_off, B, LP, T, B_real=None, need_grad=True):
        D, P, K = y.shape[1], model.pred_len, self.K
        if model.training and model.attention_net[2].p > 0:
            raise NotImplementedError("ComiRec: dropout inside attention_net (hidden_dropout_prob > 0 in training) is not built")
        h, a = self._logits(model, y, T)
        rr = da_rr = None
        if float(getattr(model, "lambda_rr", 0.0)) > 0:
            rr, da_rr = self._routing_reg(model, a, seq_off, B if B_real is None else B_real, T, need_grad)
        u, M, S = self._pool(y, a, seq_off, B, T)
        traw = torch.empty((B * LP, D), dtype=torch.float32, device=y.device)
        L.call("b200rec_gather_rows", W.data_ptr(), D, items.reshape(-1).data_ptr(), B * LP, traw.data_ptr(), L.F32, L.stream())
        hd = torch.empty((T, P, D), dtype=torch.float32, device=y.device)
        sel = torch.empty((T, P), dtype=torch.int32, device=y.device)
        L.call("b200rec_comi_select_fwd", u.data_ptr(), traw.data_ptr(), tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P, K, D,
               hd.data_ptr(), sel.data_ptr(), L.stream())
        return hd, dict(y=y, h=h, a=a, u=u, M=M, S=S, sel=sel, rr=rr, da_rr=da_rr)

    def train_backward(self, model, d_hd, ctx, grads):
        c = ctx["comi"]
        y, h, a, u = c["y"], c["h"], c["a"], c["u"]
        T, D = y.shape
        P, K, Hd = model.pred_len, self.K, self.hidden
        B = ctx["seq_off"].numel() - 1
        st, dev = L.stream(), y.device
        lin1, lin2 = model.attention_net[0], model.attention_net[3]
        du = torch.empty((T, K, D), dtype=torch.float32, device=dev)
        L.call("b200rec_comi_select_bwd", d_hd.data_ptr(), c["sel"].data_ptr(), T, P, K, D, du.data_ptr(), st)
        dy = torch.zeros((T, D), dtype=torch.float32, device=dev)
        da = torch.zeros((T, K), dtype=torch.float32, device=dev)
        L.call("b200rec_comi_pool_bwd", du.data_ptr(), u.data_ptr(), y.data_ptr(), a.data_ptr(), c["M"].data_ptr(),
               c["S"].data_ptr(), ctx["seq_off"].data_ptr(), B, K, D, dy.data_ptr(), da.data_ptr(), st)
        if c.get("da_rr") is not None:                      # + lambda_rr * d rr / d a, scaled like every other gradient
            da.add_(c["da_rr"] * (ctx["gscale"].reshape(()) * float(model.lambda_rr)))
        # attention net: a = h W2^T, h = tanh(y W1^T + b1)
        dW2 = torch.empty((K, Hd), dtype=torch.float32, device=dev)
        L.gemm(da, h, dW2, K, Hd, T, lda=K, a_major=1, ldb=Hd, b_major=1, ldc=Hd)
        dh = torch.empty((T, Hd), dtype=torch.float32, device=dev)
        L.gemm(da, lin2.weight.data, dh, T, Hd, K, lda=K, ldb=Hd, b_major=1, ldc=Hd)
        L.call("b200rec_comi_tanh_bwd", h.data_ptr(), dh.data_ptr(), dh.numel(), st)          # dh -> dz1
        dW1 = torch.empty((Hd, D), dtype=torch.float32, device=dev)
        L.gemm(dh, y, dW1, Hd, D, T, lda=Hd, a_major=1, ldb=D, b_major=1, ldc=D)
        if lin1.bias is not None:
            db1 = torch.empty(Hd, dtype=torch.float32, device=dev)
            L.colsum(dh, T, Hd, Hd, db1)
            grads[lin1.bias] = db1
        L.gemm(dh, lin1.weight.data, dy, T, D, Hd, lda=Hd, ldb=D, b_major=1, ldc=D, epilogue=L.EPI_ACCUM)
        grads[lin1.weight], grads[lin2.weight] = dW1, dW2
        return dy

    def predict_heads(self, model, y, seq_off, B, T):
        """The K interests of the whole sequence = the pooled interests at every sequence's last token."""
        _, a = self._logits(model, y, T)
        u, _, _ = self._pool(y, a, seq_off, B, T)
        last = (seq_off[1:] - 1).long()
        KD = self.K * y.shape[1]
        out = torch.empty((B, self.K, y.shape[1]), dtype=torch.float32, device=y.device)
        L.call("b200rec_gather_rows", u.data_ptr(), KD, last.data_ptr(), B, out.data_ptr(), L.F32, L.stream())
        return out


class ComiRec(HSTU):
    """REC/model/IDNet/comirec.py:21 `ComiRec(config, dataload)`."""

    def __init__(self, config, dataload, compute_dtype=torch.bfloat16):
        cfg = _hstu_config(config)
        if cfg["item_embedding_size"] != cfg["hstu_embedding_size"]:
            raise NotImplementedError("ComiRec: item_embedding_size != hstu_embedding_size (tower) is not built")
        if config.get("skip_hstu", False):
            raise NotImplementedError("ComiRec: skip_hstu is not built")
        D = cfg["hstu_embedding_size"]
        self._comi_K = int(config.get("interest_num", None) or 4)                       # comirec.py:89
        self._comi_hidden, net_bias = self._attention_net_shape(config, D)
        self._readout = _Readout(self._comi_K, self._comi_hidden)
        super().__init__(cfg, dataload, compute_dtype)
        self.num_interest = self._comi_K
        self.medusa_num_heads = self._comi_K            # eval: one score row per interest (comirec.py:409-411)
        self.attention_net = nn.Sequential(             # comirec.py:91-96
            nn.Linear(D, self._comi_hidden, bias=net_bias), nn.Tanh(), nn.Dropout(self._linear_dropout_rate),
            nn.Linear(self._comi_hidden, self._comi_K, bias=False))
        for p in self.attention_net.parameters():       # comirec.py:135-146 (reset_params)
            truncated_normal_(p.data, mean=0.0, std=0.02)

    @staticmethod
    def _attention_net_shape(config, D):
        """(hidden width, first Linear has a bias) of the attention net (comirec.py:88, 92)."""
        return int(config.get("interest_hidden", None) or D // 2), True

    def _build_jobs(self):
        """One NCE job per prediction offset: offset p has its own queries (the interest chosen against ITS target),
        the global negative set, loss weight horizon_discount[p] (comirec.py:303-333)."""
        return [_Job("nce", -1, p, 0, 1 << p, 0, 1.0, 0) for p in range(self.pred_len)]

    def forward(self, interaction, n_tokens=None, prepared=None):
        out = super().forward(interaction, n_tokens=n_tokens, prepared=prepared)
        out.pop("seg_0_loss", None)                     # the HSTU path's per-segment log has no ComiRec counterpart
        return out


class REMI(ComiRec):
    """REC/model/IDNet/remi.py:14 `REMI(config, dataload)`: ComiRec-SA on the HSTU body trained with routing
    regularisation (lambda_rr, remi.py:156-196, 356-372: b200rec_comi_rr) and interest-aware hard negatives (beta_ihn,
    remi.py:198-277: b200rec_nce_ihn_loss_fwd on the un-fused logits; beta_ihn <= 0 is the plain sampled softmax of
    ComiRec on the fused path).  predict() is ComiRec's (remi.py:439-496).  Same state-dict names; the attention net's
    first Linear has no bias when `attention_net_bias` is False, its width defaults to D * interest_hidden_ratio."""

    def __init__(self, config, dataload, compute_dtype=torch.bfloat16):
        lam, beta = config.get("lambda_rr", None), config.get("beta_ihn", None)
        self.lambda_rr = 100.0 if lam is None else float(lam)                           # remi.py:38
        self.beta_ihn = 1.0 if beta is None else float(beta)                            # remi.py:40
        self._ihn_beta = max(self.beta_ihn, 0.0)
        super().__init__(config, dataload, compute_dtype)

    @staticmethod
    def _attention_net_shape(config, D):
        ratio = config.get("interest_hidden_ratio", None)
        hidden = config.get("interest_hidden", None) or int(D * (0.5 if ratio is None else float(ratio)))   # remi.py:87
        bias = config.get("attention_net_bias", None)
        return int(hidden), True if bias is None else bool(bias)                                        # remi.py:91


def _hstu_config(config):
    """ComiRec's config keys -> the HSTU keys the shared path reads (single identity head, plain NCE)."""
    class _Cfg(dict):
        def __getitem__(self, k):
            return dict.get(self, k, None)

        def get(self, k, d=None):
            v = dict.get(self, k, None)
            return d if v is None else v
    cfg = _Cfg(dict(config))
    cfg.update(loss=config["loss"] or "nce", num_segment_head=1, num_prior_head=1, head_interaction="multiplicative",
               medusa_num_layers=0, neg_sample_by_cat=False, prior_switch=None, pos_sample_mix_ratio=0)
    if cfg["loss"] != "nce":
        raise NotImplementedError(f"loss={cfg['loss']} is not supported")          # comirec.py:112-113
    return cfg
