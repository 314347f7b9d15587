"""Fused AdamW over the HSTU parameters (torch.optim.AdamW maths, reference trainer.py:292-299).

Dense parameters are updated by ONE multi-tensor launch; the item-embedding table takes the
dense-equivalent row kernel fed by the compact (unique id, row) gradient of the sorted-segment
scatter-add, so the [N, D] dense gradient is never materialised (`sparse_embedding_grad=True`).
With `device_step=True` the step counter / bias corrections / lr live in device memory and are
advanced by a tick kernel, so the whole optimizer step can sit inside a captured CUDA graph.
"""
import struct

import torch

from . import _lib as L


class FusedAdamW(object):
    # The lazy table update keeps the scalars of the last HIST_CAP optimizer steps in a ring; every HIST_CAP // 2 steps
    # all rows are brought up to date (one dense pass) so that no row ever needs an entry that was overwritten.
    HIST_CAP = 1 << 16

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, device_step=False,
                 lazy_table=False):
        """lazy_table: defer the dense-equivalent update of item-table rows nobody reads (exact AdamW values, see
        include/b200rec.h); the model brings rows up to date before every lookup / evaluation / state_dict()."""
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.state = {}
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self._row_slot = None
        self.device_step = device_step or lazy_table
        self.lazy_table = bool(lazy_table)
        self._last = self._hist = None
        if self.lazy_table:
            model._table_sync, model._table_flush = self.sync_rows, self.flush
        self._coef = None          # device fp32[4] = {lr, bc1, bc2_sqrt, step}
        self._table_key = None
        self._table = self._blocks = None
        self._keepalive = []
        self._half_done = False

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

    def set_lr(self, lr):
        self.param_groups[0]["lr"] = lr
        if self._coef is not None:
            self._coef[0:1].fill_(lr)

    def _dense_table(self, dense):
        """Device pointer table for the multi-tensor kernel; rebuilt only when a pointer changes."""
        shadow_of = getattr(self.model, "shadow_of", lambda p: None)
        shadows = [shadow_of(p) for p in dense]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel(), 0 if sh is None else sh.data_ptr())
                    for p, sh in zip(dense, shadows))
        if key != self._table_key:
            recs, blocks = [], []
            for t, p in enumerate(dense):
                m, v = self._st(p)
                sh = shadows[t]
                recs.append(struct.pack("QQQQQq", p.data_ptr(), m.data_ptr(), v.data_ptr(), p.grad.data_ptr(),
                                        0 if sh is None else sh.data_ptr(), p.numel()))
                for beg in range(0, p.numel(), 4096):
                    blocks += [t, beg]
            dev = dense[0].device
            # pinned staging buffers: legal inside CUDA-graph capture (the copy becomes a memcpy node that
            # re-reads the pinned buffer at replay, so the buffers are kept alive for the optimizer's life)
            raw_h = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).pin_memory()
            blk_h = torch.tensor(blocks, dtype=torch.int64).pin_memory()
            if torch.cuda.is_current_stream_capturing():
                self._keepalive.append((raw_h, blk_h))     # a graph re-reads its staging buffers at every replay
            else:
                self._eager_staging = (raw_h, blk_h)       # eager: the host allocator defers reuse until the copy ran
            self._table = raw_h.to(dev, non_blocking=True)
            self._blocks = blk_h.to(dev, non_blocking=True)
            self._table_key = key
        return self._table, self._blocks

    @torch.no_grad()
    def step(self, grad_scale=1.0, rows=True, dense=True):
        """rows / dense select the item-table pass and the multi-tensor pass; a caller that overlaps the dense
        all-reduce with the table update calls step(dense=False) then step(rows=False) (ONE optimizer step: the
        counter advances on the first of the two calls)."""
        first_half = rows or not self._half_done
        if first_half:
            self.step_count += 1
        self._half_done = rows and not dense
        g = self.param_groups[0]
        lr, (b1, b2), eps, wd = g["lr"], g["betas"], g["eps"], g["weight_decay"]
        st = L.stream()
        emb = self.model.item_embedding.weight
        coef = None
        if self.device_step:
            if self._coef is None:
                self._coef = torch.tensor([lr, 0.0, 0.0, float(self.step_count - 1)], dtype=torch.float32, device=emb.device)
            if first_half:
                if self.lazy_table:
                    self._lazy_state(completed=self.step_count - 1)
                    self.rebase_if_due(self.step_count - 1)
                    L.call("b200rec_adamw_tick_hist", self._coef.data_ptr(), self._hist.data_ptr(), self.HIST_CAP, b1, b2, wd, st)
                else:
                    L.call("b200rec_adamw_tick", self._coef.data_ptr(), b1, b2, st)
            coef = self._coef.data_ptr()
        if self.device_step and rows and not dense:
            # the dense half may run after the next step's lr was set: it keeps the scalars of ITS step
            if getattr(self, "_coef_pending", None) is None:
                self._coef_pending = torch.empty_like(self._coef)
            self._coef_pending.copy_(self._coef)
        elif self.device_step and dense and not rows and getattr(self, "_coef_pending", None) is not None:
            coef = self._coef_pending.data_ptr()
        dense_list = []
        for p in self.model.parameters():
            if p is emb and self.lazy_table and p.grad is not None:
                raise L.B200RecError("lazy_table needs the compact table gradient (config['sparse_embedding_grad'] = True)")
            if p is emb and p.grad is None and self.model.emb_grad is not None:
                if not rows:
                    continue
                m, v = self._st(p)
                uniq_ids, uniq_rows, n_uniq = self.model.emb_grad
                N, D = p.shape
                if self.lazy_table:
                    L.call("b200rec_adamw_rows_lazy", p.data_ptr(), m.data_ptr(), v.data_ptr(), N, D, uniq_ids.data_ptr(),
                           uniq_rows.data_ptr(), n_uniq.data_ptr(), uniq_ids.numel(), self._last.data_ptr(),
                           self._hist.data_ptr(), self.HIST_CAP, coef, b1, b2, eps, wd, grad_scale, st)
                    continue
                if self._row_slot is None or self._row_slot.numel() != N:
                    self._row_slot = torch.empty(N, dtype=torch.int32, device=p.device)
                L.call("b200rec_adamw_rows", p.data_ptr(), m.data_ptr(), v.data_ptr(), N, D, uniq_ids.data_ptr(),
                       uniq_rows.data_ptr(), n_uniq.data_ptr(), self._row_slot.data_ptr(), lr, b1, b2, eps, wd,
                       self.step_count, grad_scale, coef, st)
                continue
            if p.grad is None:
                continue
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                p.grad = p.grad.float().contiguous()
            dense_list.append(p)
        if dense_list and dense:
            dense = dense_list
            table, blocks = self._dense_table(dense)
            L.call("b200rec_adamw_multi", table.data_ptr(), blocks.data_ptr(), blocks.numel() // 2, lr, b1, b2, eps, wd,
                   self.step_count, grad_scale, coef, st)

    # ---- lazy item-table update -------------------------------------------------------------------------
    def _lazy_state(self, completed=None):
        """Creates last[row] / hist / the device step counter on first use; returns the table Parameter.
        `completed` = optimizer steps already applied to every row (defaults to step_count)."""
        emb = self.model.item_embedding.weight
        completed = self.step_count if completed is None else completed
        if self._coef is None:                     # coef[3] = number of completed optimizer steps
            g = self.param_groups[0]
            self._coef = torch.tensor([g["lr"], 0.0, 0.0, float(completed)], dtype=torch.float32, device=emb.device)
        if self._last is None or self._last.numel() != emb.shape[0] or self._last.device != emb.device:
            self._last = torch.full((emb.shape[0],), int(completed), dtype=torch.int32, device=emb.device)
            self._hist = torch.zeros((self.HIST_CAP, 4), dtype=torch.float32, device=emb.device)
        return emb

    @torch.no_grad()
    def sync_rows(self, ids, id_stride=1, id_offset=0):
        """Bring the table rows `ids` (int64 tensor, duplicates fine; None = every row) up to the current step.
        id_stride > 1: `ids` are global ids of a row-sharded table; only ids % id_stride == id_offset are mine."""
        if not self.lazy_table or (self.step_count == 0 and self._last is None):
            return
        emb = self._lazy_state()
        g = self.param_groups[0]
        (b1, b2), eps, wd = g["betas"], g["eps"], g["weight_decay"]
        m, v = self._st(emb)
        N, D = emb.shape
        if ids is not None:
            ids = ids.reshape(-1).contiguous()
        L.call("b200rec_adamw_rows_catchup", emb.data_ptr(), m.data_ptr(), v.data_ptr(), N, D, L.ptr(ids),
               N if ids is None else ids.numel(), self._last.data_ptr(), self._hist.data_ptr(), self.HIST_CAP,
               self._coef.data_ptr(), b1, b2, eps, wd, int(id_stride), int(id_offset), L.stream())

    def flush(self):
        self.sync_rows(None)

    def rebase_if_due(self, completed):
        """Called (outside any graph capture / replay) before step `completed + 1`: when half of the history ring
        has been consumed since the last full pass, every row is brought up to date so the ring can wrap."""
        if self.lazy_table and completed > 0 and completed % (self.HIST_CAP // 2) == 0 and \
                not torch.cuda.is_current_stream_capturing():
            self.flush()

    def mark_all_current(self):
        """After loading a checkpoint: every row holds the state of step `step_count`."""
        if self.lazy_table:
            self._lazy_state()
            self._last.fill_(int(self.step_count))
