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
    def __init__(self, model, optimizer, example_batch, bucket=128, warmup=2):
        assert optimizer.device_step, "GraphedTrainStep needs FusedAdamW(device_step=True)"
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
        with torch.cuda.graph(g, pool=self.pool):
            out = self._eager(T_b)
        if self.pool is None:
            self.pool = g.pool()
        for t, c in zip(state, snap):
            t.copy_(c)
        self.opt.step_count = step0
        return g, out, L.launches - n0

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
        g.replay()
        self.opt.step_count += 1
        L.launches += n_launch
        return out
