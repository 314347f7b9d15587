"""Data-parallel plumbing over torch.distributed (one process per GPU; NCCL on the GPU box, gloo in
the CPU tests).  Dense HSTU / head gradients are averaged with one flat all-reduce; the item-table
gradient travels in compact (unique id, row) form: every rank all-gathers the others' rows and
re-runs the deterministic sorted-segment reduction, so all replicas apply the same update.
(Reference: DDP / ZeRO-2 gradient averaging, trainer.py:434-453)
"""
import torch
import torch.distributed as dist

from . import _lib as L


def flatten_dense_grads(params):
    """Returns (flat fp32 buffer, views) over the gradients of `params` (grads are re-pointed to views)."""
    with_grad = [p for p in params if p.grad is not None]
    if not with_grad:
        return None, []
    total = sum(p.grad.numel() for p in with_grad)
    flat = torch.empty(total, dtype=torch.float32, device=with_grad[0].grad.device)
    off = 0
    for p in with_grad:
        n = p.grad.numel()
        flat[off:off + n].copy_(p.grad.reshape(-1))
        p.grad = flat[off:off + n].view_as(p.grad)
        off += n
    return flat, with_grad


def merge_compact_rows(ids_list, rows_list, D, scale):
    """Concatenate per-rank (ids, rows), reduce duplicates deterministically (rank order, then row order)."""
    ids = torch.cat(ids_list)
    rows = torch.cat(rows_list)
    if scale != 1.0:
        rows = rows * scale
    n = ids.numel()
    dev = ids.device
    ws_bytes = L.lib().b200rec_scatter_add_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    uid = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    urows = torch.empty((max(n, 1), D), dtype=torch.float32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("b200rec_scatter_add_sorted", ids.data_ptr(), n, rows.data_ptr(), D, uid.data_ptr(), urows.data_ptr(),
           nu.data_ptr(), ws.data_ptr(), ws_bytes, L.stream())
    return uid, urows, nu


class DataParallel(object):
    def __init__(self, model, optimizer=None, group=None):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group)

    def sync_gradients(self):
        W = self.world
        emb = self.model.item_embedding.weight
        dense = [p for p in self.model.parameters() if p is not emb or p.grad is not None]
        flat, _ = flatten_dense_grads(dense)
        if flat is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / W)
        if self.model.emb_grad is not None and emb.grad is None:
            uid, urows, nu = self.model.emb_grad
            D = urows.shape[1]
            k = nu.to(torch.int64)
            ks = [torch.zeros_like(k) for _ in range(W)]
            dist.all_gather(ks, k, group=self.group)
            ks = [int(x.item()) for x in ks]
            kmax = max(max(ks), 1)
            send_ids = torch.full((kmax,), -1, dtype=torch.int64, device=uid.device)
            send_rows = torch.zeros((kmax, D), dtype=torch.float32, device=uid.device)
            mine = ks[dist.get_rank(self.group)]
            send_ids[:mine] = uid[:mine]
            send_rows[:mine] = urows[:mine]
            all_ids = [torch.empty_like(send_ids) for _ in range(W)]
            all_rows = [torch.empty_like(send_rows) for _ in range(W)]
            dist.all_gather(all_ids, send_ids, group=self.group)
            dist.all_gather(all_rows, send_rows, group=self.group)
            self.model.emb_grad = merge_compact_rows(all_ids, all_rows, D, 1.0 / W)
