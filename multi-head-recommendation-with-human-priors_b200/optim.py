"""Fused AdamW over the HSTU parameters (torch.optim.AdamW maths, reference trainer.py:292-299).

Dense parameters take one fused pass each.  The item-embedding table takes the dense-equivalent
row kernel fed by the compact (unique id, row) gradient of the sorted-segment scatter-add, so the
[N, D] dense gradient is never materialised (set `sparse_embedding_grad=True` on the model).
"""
import torch

from . import _lib as L


class FusedAdamW(object):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.state = {}
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self._row_slot = None

    def _st(self, p):
        s = self.state.get(p)
        if s is None:
            s = (torch.zeros_like(p.data), torch.zeros_like(p.data))
            self.state[p] = s
        return s

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None
        self.model.emb_grad = None

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.step_count += 1
        g = self.param_groups[0]
        lr, (b1, b2), eps, wd = g["lr"], g["betas"], g["eps"], g["weight_decay"]
        st = L.stream()
        emb = self.model.item_embedding.weight
        for p in self.model.parameters():
            if p is emb and p.grad is None and self.model.emb_grad is not None:
                m, v = self._st(p)
                uniq_ids, uniq_rows, n_uniq = self.model.emb_grad
                N, D = p.shape
                if self._row_slot is None or self._row_slot.numel() != N:
                    self._row_slot = torch.empty(N, dtype=torch.int32, device=p.device)
                L.call("b200rec_adamw_rows", p.data_ptr(), m.data_ptr(), v.data_ptr(), N, D, uniq_ids.data_ptr(),
                       uniq_rows.data_ptr(), n_uniq.data_ptr(), self._row_slot.data_ptr(), lr, b1, b2, eps, wd,
                       self.step_count, grad_scale, st)
                continue
            if p.grad is None:
                continue
            m, v = self._st(p)
            gr = p.grad.contiguous()
            if gr.dtype != torch.float32:
                gr = gr.float()
            L.call("b200rec_adamw", p.data_ptr(), m.data_ptr(), v.data_ptr(), gr.data_ptr(), p.numel(), lr, b1, b2,
                   eps, wd, self.step_count, grad_scale, st)
