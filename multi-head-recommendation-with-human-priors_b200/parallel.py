"""Multi-GPU plumbing over torch.distributed (one process per GPU; NCCL on the GPU box, gloo in the
CPU tests).  SURVEY §8(e):

  * dense HSTU / head weights: data parallel, gradients averaged with one flat all-reduce
    (reference: DDP / ZeRO-2, trainer.py:434-453);
  * negatives: ids are all-gathered so every rank scores against the global negative set
    (reference all-gathers the embeddings, basemodel.py:11-22 / hstu.py:673,755);
  * item table: either replicated (compact (id,row) gradients all-gathered and re-reduced
    identically on every rank) or ROW-SHARDED by `id % world` (`ShardedTable`): lookups and
    gradient rows travel by all-to-all, the owner runs the deterministic sorted-segment reduction
    and the optimizer on its shard only;
  * eval with a sharded table: every rank scores all gathered users against its rows, local
    top-K lists are all-gathered and merged (`merge_topk`).

The row gather / segment reduction used on the local shard are the CUDA kernels of libb200rec; the
CPU tests inject torch equivalents through `row_gather` / `segment_reduce` to exercise the routing.
"""
import torch
import torch.distributed as dist

from . import _lib as L


# ------------------------------------------------------------------------------- local kernels
def cuda_row_gather(table, idx):
    out = torch.empty((idx.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    L.call("b200rec_gather_rows", table.data_ptr(), table.shape[1], idx.data_ptr(), idx.numel(), out.data_ptr(),
           L.F32, L.stream())
    return out


def cuda_segment_reduce(ids, rows):
    """Deterministic sum of rows per id (ids <= 0 dropped).  Returns (uniq_ids, uniq_rows, n_uniq dev int32)."""
    n, D = ids.numel(), rows.shape[1]
    dev = ids.device
    ws_bytes = L.lib().b200rec_scatter_add_workspace_bytes(max(n, 1))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    uid = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    urows = torch.empty((max(n, 1), D), dtype=torch.float32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("b200rec_scatter_add_sorted", ids.data_ptr(), n, rows.data_ptr(), D, uid.data_ptr(), urows.data_ptr(),
           nu.data_ptr(), ws.data_ptr(), ws_bytes, L.stream())
    return uid, urows, nu


def merge_compact_rows(ids_list, rows_list, D, scale):
    ids = torch.cat(ids_list)
    rows = torch.cat(rows_list)
    if scale != 1.0:
        rows = rows * scale
    return cuda_segment_reduce(ids, rows)


# ------------------------------------------------------------------------------- helpers
def world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def gather_negative_ids(neg_items, group=None):
    """[B, sets, n] per rank -> [W*B, sets, n] in rank order (the order of the reference's
    torch.stack(all_gather(...)).reshape(-1, D), basemodel.py:17-18)."""
    W = world(group)
    if W == 1:
        return neg_items
    neg_items = neg_items.contiguous()
    out = [torch.empty_like(neg_items) for _ in range(W)]
    dist.all_gather(out, neg_items, group=group)
    return torch.cat(out, dim=0)


def flatten_dense_grads(params):
    """Returns (flat fp32 buffer, params) over the gradients of `params` (grads are re-pointed to views)."""
    with_grad = [p for p in params if p.grad is not None]
    if not with_grad:
        return None, []
    # every view starts on a 16-byte boundary (vector loads in the fused optimizer)
    total = sum((p.grad.numel() + 3) // 4 * 4 for p in with_grad)
    flat = torch.zeros(total, dtype=torch.float32, device=with_grad[0].grad.device)
    off = 0
    for p in with_grad:
        n = p.grad.numel()
        flat[off:off + n].copy_(p.grad.reshape(-1))
        p.grad = flat[off:off + n].view_as(p.grad)
        off += (n + 3) // 4 * 4
    return flat, with_grad


def _a2a(send, send_counts, recv_counts, group):
    """all_to_all_single with per-peer row counts (python lists)."""
    out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(recv_counts),
                           input_split_sizes=list(send_counts), group=group)
    return out


# ------------------------------------------------------------------------------- sharded table
class ShardedTable(object):
    """Item table sharded by rows: global id g lives on rank g % W at local row g // W.

    fetch(ids)        -> rows of arbitrary global ids (ids deduplicated by the caller), via two
                         all-to-alls (ids out, rows back); keeps the routing plan.
    push_grads(rows)  -> sends one gradient row per fetched id back to its owner along the same
                         plan; the owner reduces duplicates deterministically and returns the
                         compact gradient of ITS shard in local row indices.
    """

    def __init__(self, local_weight, item_num, group=None, row_gather=cuda_row_gather,
                 segment_reduce=cuda_segment_reduce):
        self.weight = local_weight
        self.item_num = item_num
        self.group = group
        self.W = world(group)
        self.rank = dist.get_rank(group) if self.W > 1 else 0
        self.row_gather, self.segment_reduce = row_gather, segment_reduce
        self.pre_gather = None     # hook: called with the local row indices right before they are read
        self.plan = None

    @staticmethod
    def local_rows(item_num, W, rank):
        return (item_num - rank + W - 1) // W

    @staticmethod
    def shard_of(full_weight, W, rank):
        return full_weight[rank::W].contiguous()

    def fetch(self, ids, out=None):
        """Rows of `ids` (deduplicated by the caller) in the order of `ids`; written into out[:len(ids)] if given."""
        W = self.W
        owner = ids % W
        order = torch.argsort(owner, stable=True)
        sorted_ids = ids[order]
        send_counts = torch.bincount(owner, minlength=W)
        recv_counts = torch.empty_like(send_counts)
        if W > 1:
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        else:
            recv_counts.copy_(send_counts)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        req = _a2a(sorted_ids, sc, rc, self.group) if W > 1 else sorted_ids
        local_idx = torch.div(req, W, rounding_mode="floor")
        if self.pre_gather is not None:
            self.pre_gather(local_idx)
        rows = self.row_gather(self.weight, local_idx)
        back = _a2a(rows, rc, sc, self.group) if W > 1 else rows
        out = torch.empty_like(back) if out is None else out[:back.shape[0]]
        out[order] = back
        self.plan = (order, sc, rc, local_idx)
        return out

    def push_grads(self, grad_rows, scale=1.0):
        """grad_rows[i] belongs to the i-th id of the last fetch.  Returns (local_row_ids, rows, n_uniq)
        with ids shifted by +1 removed (plain local row indices); padding id 0 never gets a gradient."""
        order, sc, rc, local_idx = self.plan
        send = grad_rows[order]
        recv = _a2a(send, sc, rc, self.group) if self.W > 1 else send
        if scale != 1.0:
            recv = recv * scale
        # +1 shift: the segment reduction drops keys <= 0, and local row 0 of rank 0 IS padding id 0
        keys = local_idx + 1
        if self.rank == 0:
            keys = torch.where(local_idx == 0, torch.zeros_like(keys), keys)
        uid, urows, nu = self.segment_reduce(keys, recv)
        return uid - 1, urows, nu


# ------------------------------------------------------------------------------- peer (CUDA IPC) access
_IPC_OPENED = {}     # process-wide: handle bytes -> mapped base (a handle can be opened once per process)


class PeerMap(object):
    """Maps same-shaped device buffers of every rank into this process with CUDA IPC (one NVLink / NVSwitch node).

    `table(t)` returns a device int64[W] tensor of pointers: entry r addresses rank r's buffer (entry `rank` is the
    local pointer).  Kernels index it directly (b200rec_gather_rows_sharded, b200rec_scatter_add_sorted_peer), so
    lookups and gradient rows cross NVLink inside our own kernels: no all-to-all, no host-synchronised split sizes.
    The exchange of the 72-byte (handle, offset) records is the only collective, once per buffer."""

    def __init__(self, group=None):
        self.group = group
        self.W = world(group)
        self.rank = dist.get_rank(group) if self.W > 1 else 0
        self._opened = _IPC_OPENED
        self._keep = []            # exported tensors must outlive the peers' mappings

    def table(self, t):
        import ctypes as C
        assert t.is_cuda and t.is_contiguous()
        self._keep.append(t)
        ptrs = [0] * self.W
        ptrs[self.rank] = t.data_ptr()
        if self.W > 1:
            handle = (C.c_char * 64)()
            off = C.c_int64(0)
            L.call("b200rec_ipc_export", t.data_ptr(), C.addressof(handle), C.addressof(off))
            rec = torch.frombuffer(bytearray(bytes(handle) + int(off.value).to_bytes(8, "little")), dtype=torch.uint8)
            rec = rec.to(t.device)
            recs = [torch.empty_like(rec) for _ in range(self.W)]
            dist.all_gather(recs, rec, group=self.group)
            for r in range(self.W):
                if r == self.rank:
                    continue
                raw = bytes(recs[r].cpu().numpy().tobytes())
                h, o = raw[:64], int.from_bytes(raw[64:72], "little")
                base = self._opened.get(h)
                if base is None:
                    out = C.c_void_p(0)
                    hb = (C.c_char * 64).from_buffer_copy(h)
                    L.call("b200rec_ipc_import", C.addressof(hb), C.addressof(out))
                    base = int(out.value)
                    self._opened[h] = base
                ptrs[r] = base + o
        return torch.tensor(ptrs, dtype=torch.int64, device=t.device)

    def release(self):
        """Drops the references that keep the exported buffers alive (mappings stay until the process exits)."""
        self._keep.clear()


def cuda_gather_rows_sharded(shard_ptrs, W, D, ids, out):
    """out[i] = shard[ids[i] % W][ids[i] // W] over the peer pointer table (ids < 0: zero row)."""
    L.call("b200rec_gather_rows_sharded", shard_ptrs.data_ptr(), W, D, ids.data_ptr(), ids.numel(), out.data_ptr(),
           L.stream())
    return out


def cuda_segment_reduce_peer(ids_all, n_per, W, rank, src_ptrs, D, max_uniq):
    """Owner-side reduction over every rank's gradient-row buffer (see b200rec_scatter_add_sorted_peer).
    Returns (local_row_ids, rows, n_uniq)."""
    n = n_per * W
    dev = ids_all.device
    ws_bytes = L.lib().b200rec_scatter_add_workspace_bytes(max(n, 1))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    uid = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    urows = torch.empty((max(1, min(n, max_uniq)), D), dtype=torch.float32, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("b200rec_scatter_add_sorted_peer", ids_all.data_ptr(), n_per, W, rank, src_ptrs.data_ptr(), D, uid.data_ptr(),
           urows.data_ptr(), nu.data_ptr(), ws.data_ptr(), ws_bytes, L.stream())
    return uid, urows, nu


# ------------------------------------------------------------------------------- data parallel
class DataParallel(object):
    """Gradient synchronisation for one training step (call between backward and optimizer.step)."""

    def __init__(self, model, optimizer=None, group=None):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = world(group)

    def sync_gradients(self):
        W = self.world
        if W == 1:
            return
        emb = self.model.item_embedding.weight
        dense = [p for p in self.model.parameters() if p is not emb or p.grad is not None]
        flat, _ = flatten_dense_grads(dense)
        if flat is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / W)
        if getattr(self.model, "sharded_table", None) is not None:
            return  # the owner already holds the rank-summed, 1/W-scaled gradient of its shard
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


# ------------------------------------------------------------------------------- sharded eval
def merge_topk(val_list, idx_list, head_list, K):
    """Cross-shard merge of per-shard top-K lists ([B, K] each, global item ids): value desc, id asc.
    The lists are disjoint in item id, so no de-duplication is needed."""
    v = torch.cat(val_list, dim=1)
    i = torch.cat(idx_list, dim=1)
    h = torch.cat(head_list, dim=1)
    # two stable sorts = lexicographic (value desc, id asc)
    o1 = torch.argsort(i, dim=1, stable=True)
    v, i, h = v.gather(1, o1), i.gather(1, o1), h.gather(1, o1)
    o2 = torch.argsort(v, dim=1, descending=True, stable=True)
    return i.gather(1, o2)[:, :K].contiguous(), v.gather(1, o2)[:, :K].contiguous(), h.gather(1, o2)[:, :K].contiguous()
