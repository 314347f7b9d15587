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
    return [(e0.elapsed_time(e1), fl) for (e0, e1, fl) in events]


class GraphedShardedStep(object):
    """Multi-GPU step with a row-sharded item table (`HSTU.shard_item_table`), split in three:

      pre   (eager)  negative-id all-gather, unique, all-to-all row fetch into a fixed-capacity row cache
                     (`HSTU.prepare_rows`): collectives and data-dependent sizes live here;
      graph (replay) forward + backward on static buffers -> dense gradients in ONE flat buffer shared by
                     all buckets, and one gradient row per cache row;
      post  (eager)  all-to-all push of the cache-row gradients to their owners (deterministic segment reduce
                     there) and the table update; the flat all-reduce of the dense gradients is launched on its
                     own communicator and the dense AdamW of step k runs at the start of step k+1, AFTER that
                     step's pre-phase: the all-reduce hides behind the id exchange / row fetch, which only need
                     the table.  `flush()` applies the pending dense update (call it before evaluating or saving).

    The eager part is ~40 launches instead of ~1200, so the ranks stay GPU-bound.
    """

    def __init__(self, model, optimizer, example_batch, bucket=128, warmup=2, group=None, instrument=False):
        import torch.distributed as dist
        self.instrument = instrument
        self.gemm_events = {}
        assert model.sharded_table is not None, "call model.shard_item_table() first"
        assert not model._has_tower(), "item_id_proj_tower: use the eager sharded step"
        self.dist, self.group = dist, group
        self.model, self.opt, self.bucket, self.warmup = model, optimizer, bucket, warmup
        items, neg, mask, tags = example_batch
        dev = items.device
        self.world = model.sharded_table.W
        B, LP = items.shape
        n_sets = neg.shape[1]
        n_neg = neg.shape[0] * neg.shape[2] * (self.world if model.share_negatives else 1)
        D = model.item_embedding.weight.shape[1]
        self.cap = (B + 1) * LP + n_sets * n_neg             # every requested id distinct
        self.static = tuple(torch.empty_like(t) for t in example_batch)
        i64 = dict(dtype=torch.int64, device=dev)
        self.prep = dict(W=torch.zeros((self.cap, D), dtype=torch.float32, device=dev),
                         items_idx=torch.zeros((B + 1, LP), **i64), neg_idx=torch.zeros((n_sets, n_neg), **i64),
                         gl_items=torch.zeros((B + 1, LP), **i64), gl_neg=torch.zeros((n_sets, n_neg), **i64),
                         cached=True, push=False, n_rows=None)
        self.graphs = {}
        self.pool = None
        self.max_tokens = B * model.max_seq_length
        self.dense = self.flat = self.views = None
        self._pending = None                     # async all-reduce of the previous step's dense gradients
        self.ar_group = dist.new_group() if (self.world > 1 and group is None) else group

    def _fill(self, batch):
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        p = self.model.prepare_rows(self.static[0], self.static[1], static=True, cache_out=self.prep["W"])
        U = p["n_rows"]                                   # rows were fetched straight into the static cache
        for k in ("items_idx", "neg_idx", "gl_items", "gl_neg"):
            self.prep[k].copy_(p[k])
        return U

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
        return out, self.model.cache_grad

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
                out, cache_grad = self._fwd_bwd(T_b)
        finally:
            if self.instrument:
                self.gemm_events[T_b] = L.gemm_timing
                L.gemm_timing, L.gemm_timing_external = None, False
        if self.pool is None:
            self.pool = g.pool()
        return g, out, cache_grad, L.launches - n0

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
        U = self._fill(batch)                    # pre-phase: needs the table only
        self.flush()                             # previous step's dense update (its all-reduce ran meanwhile)
        entry = self.graphs.get(T_b)
        if entry is None:
            entry = self._capture(T_b)
            self.graphs[T_b] = entry
        g, out, cache_grad, n_launch = entry
        if self.model.shadows_stale():
            self.model.invalidate_shadows()
            self.model._cast_weights()
        g.replay()
        L.launches += n_launch
        model, W = self.model, self.world
        for p, v in zip(self.dense, self.views):
            p.grad = v
        model.item_embedding.weight.grad = None
        # gradient rows to their owners first, then the dense all-reduce runs on the NCCL stream WHILE the owner
        # reduces its rows and updates its table shard (HBM-bound); the dense update waits for the all-reduce
        model.emb_grad = model.sharded_table.push_grads(cache_grad[:U], scale=1.0)
        self.opt.step(grad_scale=1.0 / W, dense=False)
        self._pending = True
        if W > 1:
            self._pending = self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.ar_group, async_op=True)
        return out
