"""CUDA-graph training step: forward + backward + fused AdamW captured once per token-count bucket and
replayed with ~zero host work (the eager step issues ~500 C-ABI launches and ~700 small torch ops, which
at B200 speed is as long as the GPU work itself).

Shapes must be static inside a graph, so the jagged token list is padded with dummy tokens up to a
multiple of `bucket` (<= bucket-1 wasted rows); the number of valid context tokens is host metadata the
collate function knows (`n_tokens = int(mask[:, :L].sum())` on the CPU copy of the batch), so no device
sync is needed to pick the graph.  All graphs share one memory pool.
"""
import torch

from . import _lib as L


class GraphedTrainStep(object):
    def __init__(self, model, optimizer, example_batch, bucket=128, warmup=2, instrument=False):
        """instrument=True captures an external CUDA event pair around every GEMM launch (bench.py reads them after a
        replay with `gemm_times()`): kernel times of the REPLAYED graph, not of an eager re-run."""
        assert optimizer.device_step, "GraphedTrainStep needs FusedAdamW(device_step=True)"
        self.instrument = instrument
        self.gemm_events = {}
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # the captured step holds no gradient exchange: replicas would silently diverge
            raise L.B200RecError("GraphedTrainStep is single-GPU; with several ranks shard the table and use "
                                 "GraphedShardedStep, or run the eager step with parallel.DataParallel")
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.static = tuple(torch.empty_like(t) for t in example_batch)
        self.graphs = {}
        self.pool = None
        self.warmup = warmup
        self.max_tokens = example_batch[0].shape[0] * model.max_seq_length

    def _eager(self, T_b):
        self.opt.zero_grad()
        out = self.model(self.static, n_tokens=T_b)
        out["loss"].backward()
        self.opt.step()
        return out

    def _state_tensors(self):
        opt = self.opt
        ts = []
        for p in self.model.parameters():
            m, v = opt._st(p)
            ts += [p.data, m, v]
        if opt._coef is None:
            g = opt.param_groups[0]
            opt._coef = torch.tensor([g["lr"], 0.0, 0.0, float(opt.step_count)], dtype=torch.float32,
                                     device=next(self.model.parameters()).device)
        ts.append(opt._coef)
        if opt.lazy_table:
            opt._lazy_state()
            ts += [opt._last, opt._hist]
        return ts

    def _capture(self, T_b):
        # the warm-up steps really train: snapshot every piece of training state and restore it afterwards
        state = self._state_tensors()
        snap = [t.clone() for t in state]
        step0 = self.opt.step_count
        # warm-up on a side stream (allocator / lazy-init effects must not be captured)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self.warmup):
                self._eager(T_b)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.launches
        if self.instrument:
            L.gemm_timing, L.gemm_timing_external = [], True
        try:
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                out = self._eager(T_b)
        finally:
            if self.instrument:
                self.gemm_events[T_b] = L.gemm_timing
                L.gemm_timing, L.gemm_timing_external = None, False
        if self.pool is None:
            self.pool = g.pool()
        for t, c in zip(state, snap):
            t.copy_(c)
        self.opt.step_count = step0
        # the bf16 weight shadows followed the warm-up updates: re-cast them from the restored masters
        self.model.invalidate_shadows()
        self.model._cast_weights()
        return g, out, L.launches - n0

    def gemm_times(self, n_tokens):
        T_b = min(self.max_tokens, max(self.bucket, (int(n_tokens) + self.bucket - 1) // self.bucket * self.bucket))
        return _read_gemm_events(self.gemm_events[T_b])

    def __call__(self, batch, n_tokens):
        """batch: (items, neg_items, mask, tags) on host (pinned) or device; n_tokens: host int."""
        T_b = min(self.max_tokens, max(self.bucket, (int(n_tokens) + self.bucket - 1) // self.bucket * self.bucket))
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        entry = self.graphs.get(T_b)
        if entry is None:
            entry = self._capture(T_b)
            self.graphs[T_b] = entry
        g, out, n_launch = entry
        if self.model.shadows_stale():           # parameters changed outside the graph (load_state_dict, ...)
            self.model.invalidate_shadows()
            self.model._cast_weights()
        self.opt.rebase_if_due(self.opt.step_count)   # lazy table: the history ring wraps every HIST_CAP steps
        g.replay()
        self.opt.step_count += 1
        L.launches += n_launch
        return out


def _read_gemm_events(events):
    """[(ms, flops)] of the GEMM launches of the last replay (call after a synchronize)."""
    return [(ev[0].elapsed_time(ev[1]), ev[2]) + tuple(ev[3:]) for ev in events]


class GraphedShardedStep(object):
    """Multi-GPU step with a row-sharded item table (`HSTU.shard_item_table`), one NVLink / NVSwitch node, with NO host
    synchronisation and no data-dependent shapes (the old id / row / gradient all-to-alls with
    host-read split sizes serialised the ranks).  Every rank maps its peers' table shards and gradient-row buffers
    once with CUDA IPC (`parallel.PeerMap`); per step:

      pre   one fixed-size all-gather of this rank's ids (items + its own negatives); the owners bring the requested
            rows up to date (lazy AdamW catch-up over the gathered list, filtered by id % W); a one-element all-reduce
            orders that against the readers; `gather_rows_sharded` then reads every needed row straight out of its
            owner's HBM over NVLink into the step's row cache (one row per requested position: no unique, no sort);
      graph forward + backward on static buffers -> dense gradients in one flat buffer, one gradient row per cache row
            in the peer-mapped buffer `g`;
      post  a one-element all-reduce (all ranks' `g` complete), then each owner reduces the gradient rows of ITS ids by
            reading them out of every rank's `g` while summing (`scatter_add_sorted_peer`, deterministic (rank, index)
            order) and runs the fused AdamW on its shard; the flat all-reduce of the dense gradients runs on its own
            communicator and the dense AdamW of step k is applied at the start of step k+1 after that step's
            pre-phase, so it hides behind the id gather / row fetch (`flush()` before evaluating / saving).

    The collectives that remain are the reference's own (negative sharing, basemodel.py:11-22; gradient averaging,
    trainer.py:434-453) plus two one-element barriers; the host only enqueues (~30 launches) and runs ahead."""

    def __init__(self, model, optimizer, example_batch, bucket=128, warmup=2, group=None, instrument=False):
        import torch.distributed as dist
        from . import parallel
        assert model.sharded_table is not None, "call model.shard_item_table() first"
        assert not model._has_tower(), "item_id_proj_tower: use the eager sharded step"
        assert model.share_negatives, "the sharded step shares negatives across ranks (reference semantics)"
        self.instrument = instrument
        self.gemm_events = {}
        self.dist, self.group, self.parallel = dist, group, parallel
        self.model, self.opt, self.bucket, self.warmup = model, optimizer, bucket, warmup
        items, neg, mask, tags = example_batch
        dev = items.device
        st = model.sharded_table
        self.world, self.rank = st.W, st.rank
        W = self.world
        B, LP = items.shape
        n_sets, n = neg.shape[1], neg.shape[2]
        self.B, self.LP, self.n_sets = B, LP, n_sets
        self.n_items = (B + 1) * LP                          # +1: the all-padding dummy row of static-token mode
        self.n_negloc = n_sets * B * n                       # this rank's negatives, set-major
        self.n_neg = W * B * n                               # global negatives per set
        self.U = self.n_items + n_sets * self.n_neg          # row-cache rows = gradient rows per rank
        D = model.item_embedding.weight.shape[1]
        self.D = D
        self.static = tuple(torch.empty_like(t) for t in example_batch)
        i64 = dict(dtype=torch.int64, device=dev)
        self.id_block = torch.zeros(self.n_items + self.n_negloc, **i64)
        self.id_all = torch.zeros((W, self.n_items + self.n_negloc), **i64)
        self.cache_ids = torch.zeros(self.U, **i64)          # global id behind every cache row of THIS rank
        self.cache_ids_all = torch.zeros((W, self.U), **i64)  # ... of every rank (the owner-side reduction reads it)
        self.g = torch.zeros((self.U, D), dtype=torch.float32, device=dev)
        self.peers = parallel.PeerMap(group)
        self.shard_ptrs = self.peers.table(model.item_embedding.weight.data)
        self.g_ptrs = self.peers.table(self.g)
        self._shard_ptr0 = model.item_embedding.weight.data_ptr()
        idx_items = torch.arange(self.n_items, **i64).view(B + 1, LP)
        idx_neg = (self.n_items + torch.arange(n_sets * self.n_neg, **i64)).view(n_sets, self.n_neg)
        self.prep = dict(W=torch.zeros((self.U, D), dtype=torch.float32, device=dev), items_idx=idx_items,
                         neg_idx=idx_neg, gl_items=torch.zeros((B + 1, LP), **i64),
                         gl_neg=torch.zeros((n_sets, self.n_neg), **i64), cached=True, push=False, n_rows=self.U,
                         grad_out=self.g, info=None)
        self.token = torch.zeros(1, dtype=torch.float32, device=dev)
        self.graphs = {}
        self.pool = None
        self.max_tokens = B * model.max_seq_length
        self.dense = self.flat = self.views = None
        self._pending = None                     # async all-reduce of the previous step's dense gradients
        self.ar_group = dist.new_group() if (self.world > 1 and group is None) else group

    def _barrier(self):
        """Stream-ordered cross-rank ordering point (one-element all-reduce): when it completes on this rank, every
        rank's stream has reached it, i.e. everything they enqueued before it has finished."""
        if self.world > 1:
            self.dist.all_reduce(self.token, op=self.dist.ReduceOp.SUM, group=self.group)

    def _fill(self, batch):
        model, W = self.model, self.world
        assert model.item_embedding.weight.data_ptr() == self._shard_ptr0, "the table shard moved: rebuild the stepper"
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        items, neg = self.static[0], self.static[1]
        B, LP, ni = self.B, self.LP, self.n_items
        # my id block: items (+ dummy row of zeros) | my negatives, set-major
        self.id_block[:B * LP].copy_(items.reshape(-1))
        self.id_block[ni:].copy_(neg.permute(1, 0, 2).reshape(-1))
        if W > 1:
            self.dist.all_gather(list(self.id_all.unbind(0)), self.id_block, group=self.group)
        else:
            self.id_all[0].copy_(self.id_block)
        # global negative set per category, rank-major (the order of the reference's all_gather, basemodel.py:17-18)
        gl_neg = self.id_all[:, ni:].reshape(W, self.n_sets, -1).permute(1, 0, 2).reshape(self.n_sets, self.n_neg)
        self.prep["gl_neg"].copy_(gl_neg)
        self.prep["gl_items"].copy_(self.id_all[self.rank, :ni].view(B + 1, LP))
        self.cache_ids_all[:, :ni].copy_(self.id_all[:, :ni])
        self.cache_ids_all[:, ni:].copy_(gl_neg.reshape(1, -1).expand(W, -1))
        self.cache_ids.copy_(self.cache_ids_all[self.rank])
        # owners bring every row somebody is about to read up to date (lazy exact AdamW), then all ranks may read
        if model._table_sync is not None:
            self.opt.sync_rows(self.id_all, id_stride=W, id_offset=self.rank)
        self._barrier()
        self.parallel.cuda_gather_rows_sharded(self.shard_ptrs, W, self.D, self.cache_ids, self.prep["W"])

    def _fwd_bwd(self, T_b):
        self.opt.zero_grad()
        out = self.model(self.static, n_tokens=T_b, prepared=self.prep)
        out["loss"].backward()
        if self.flat is None:
            emb = self.model.item_embedding.weight
            self.dense = [p for p in self.model.parameters() if p is not emb and p.grad is not None]
            sizes = [(p.numel() + 3) // 4 * 4 for p in self.dense]    # 16-byte aligned views
            self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=emb.device)
            self.views, off = [], 0
            for p, n in zip(self.dense, sizes):
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += n
        torch._foreach_copy_(self.views, [p.grad for p in self.dense])
        assert self.model.cache_grad.data_ptr() == self.g.data_ptr()
        return out

    def _capture(self, T_b):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self.warmup):
                self._fwd_bwd(T_b)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.launches
        if self.instrument:
            L.gemm_timing, L.gemm_timing_external = [], True
        try:
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                out = self._fwd_bwd(T_b)
        finally:
            if self.instrument:
                self.gemm_events[T_b] = L.gemm_timing
                L.gemm_timing, L.gemm_timing_external = None, False
        if self.pool is None:
            self.pool = g.pool()
        return g, out, L.launches - n0

    def gemm_times(self, n_tokens):
        T_b = min(self.max_tokens, max(self.bucket, (int(n_tokens) + self.bucket - 1) // self.bucket * self.bucket))
        return _read_gemm_events(self.gemm_events[T_b])

    def flush(self):
        """Dense AdamW of the last step (its all-reduce was left running)."""
        if self._pending is not None:
            work, self._pending = self._pending, None
            if work is not True:
                work.wait()
            self.opt.step(grad_scale=1.0 / self.world, rows=False)

    def __call__(self, batch, n_tokens):
        T_b = min(self.max_tokens, max(self.bucket, (int(n_tokens) + self.bucket - 1) // self.bucket * self.bucket))
        self._fill(batch)                        # pre-phase: needs the table only
        self.flush()                             # previous step's dense update (its all-reduce ran meanwhile)
        entry = self.graphs.get(T_b)
        if entry is None:
            entry = self._capture(T_b)
            self.graphs[T_b] = entry
        g, out, n_launch = entry
        if self.model.shadows_stale():
            self.model.invalidate_shadows()
            self.model._cast_weights()
        g.replay()
        L.launches += n_launch
        model, W = self.model, self.world
        for p, v in zip(self.dense, self.views):
            p.grad = v
        model.item_embedding.weight.grad = None
        # the flat all-reduce of the dense gradients starts as soon as the graph is done, on its own communicator: it
        # runs under this step's table update AND the next step's pre-phase; the dense update waits for it in flush()
        self._pending = True
        if W > 1:
            self._pending = self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.ar_group, async_op=True)
        # every rank's gradient rows are complete -> each owner reduces the rows of ITS ids out of all ranks' buffers
        # (NVLink reads inside the reduction) and updates its shard
        self._barrier()
        model.emb_grad = self.parallel.cuda_segment_reduce_peer(self.cache_ids_all, self.U, W, self.rank, self.g_ptrs,
                                                                self.D, model.item_embedding.weight.shape[0])
        self.opt.step(grad_scale=1.0 / W, dense=False)
        return out
