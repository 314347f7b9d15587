"""Full-catalogue eval scoring + top-K sweep (BASELINE.json configs[4]; reference hstu.py:979 + collector.py:191-275):
N synthetic items x H decode heads x B_e users, K = 200, at 1/2/4/8 GPUs (`bench.py --eval-sweep`).

The sweep times the scoring + masks + cross-head merge + top-K stage on its own (user embeddings are given: random unit
rows) through `HSTU.predict_topk`'s kernels: the streamed path reads the bf16 table exactly once per user
batch regardless of N.  With W GPUs the catalogue is row-sharded (`id % W`), every rank scores all users against its
rows and the per-shard lists are all-gathered and merged.  One JSON line per (N, H) point:
users/s, the table-stream count, achieved TFLOP/s and GB/s against the measured peaks.
"""
import json
import math
import os

import torch

from . import _lib as L
from . import parallel


def _unit_rows(n, D, seed, dev, dtype=torch.bfloat16, chunk=1 << 20):
    """Random unit rows generated on the device in chunks (a 50 M x 256 fp32 staging copy would not be needed twice)."""
    out = torch.empty((n, D), dtype=dtype, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        x = torch.randn((r1 - r0, D), generator=g, device=dev, dtype=torch.float32)
        out[r0:r1] = (x / x.norm(dim=1, keepdim=True)).to(dtype)
    return out


def topk_point(table, U, K, rank, W, tag_bits=None, head_cat=None, streamed=True, cap=8192):
    """U bf16 [B, H, D] (already gathered over ranks), table bf16 [N_local, D] -> local (idx, val, head) [B, K] with
    GLOBAL item ids; returns also the number of times the table was streamed."""
    dev = U.device
    B, H, D = U.shape
    N = table.shape[0]
    by_items = H <= 16                 # FOLD_ITEMS (rows = items): heads padded to a multiple of 4 only
    hp = (H if H <= 2 else (H + 3) // 4 * 4) if by_items else 32
    Up = torch.zeros((B, hp, D), dtype=U.dtype, device=dev)
    Up[:, :H] = U
    on = torch.zeros((B, hp), dtype=torch.uint8, device=dev)
    on[:, :H] = 1
    on_bits = (on.to(torch.int64) << torch.arange(hp, device=dev)).sum(dim=1).to(torch.int32)
    cat = None
    if head_cat is not None:
        cat = torch.full((hp,), -1, dtype=torch.int32, device=dev)
        cat[:H] = head_cat
    idx = torch.empty((B, K), dtype=torch.int64, device=dev)
    val = torch.empty((B, K), dtype=torch.float32, device=dev)
    hsrc = torch.empty((B, K), dtype=torch.int32, device=dev)

    def fold_gemm(b0, b1, n_rows, fval, fhead, ld, stream_args=()):
        nb = b1 - b0
        if by_items:
            L.gemm(table, Up[b0:b1].reshape(nb * hp, D), fval, n_rows, nb * hp, D, lda=D, ldb=D, ldc=ld,
                   epilogue=L.EPI_FOLD_ITEMS, C2=fhead, ldc2=ld,
                   fold_items=(hp, on_bits[b0:b1].contiguous(), cat, tag_bits, rank, W) + tuple(stream_args))
        else:
            L.gemm(Up[b0:b1].reshape(nb * hp, D), table, fval, nb * hp, n_rows, D, lda=D, ldb=D, ldc=ld,
                   epilogue=L.EPI_FOLD_HEADS, C2=fhead, ldc2=ld,
                   fold=(hp, on[b0:b1].reshape(-1).contiguous(), cat, tag_bits, rank, W) + tuple(stream_args))

    N0 = min(N, max(32768, ((N + 15) // 16 + 255) // 256 * 256))
    if streamed and N >= 4 * N0 and N < (1 << 27):
        ldn0 = (N0 + 3) // 4 * 4
        fval = torch.empty((B, ldn0), dtype=torch.float32, device=dev)
        fhead = torch.empty((B, ldn0), dtype=torch.uint8, device=dev)
        fold_gemm(0, B, N0, fval, fhead, ldn0)
        L.call("b200rec_topk_select", fval.data_ptr(), fhead.data_ptr(), B, N0, ldn0, K, None, None, rank, W,
               idx.data_ptr(), val.data_ptr(), hsrc.data_ptr(), L.stream())
        thr = val[:, K - 1].contiguous()
        cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        keys = torch.empty((B, cap), dtype=torch.int64, device=dev)
        ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        fold_gemm(0, B, N, fval, None, ldn0, (thr, cnt, keys, cap))
        L.call("b200rec_topk_from_candidates", keys.data_ptr(), cnt.data_ptr(), cap, B, K, 0, None, None,
               rank, W, idx.data_ptr(), val.data_ptr(), hsrc.data_ptr(), ovf.data_ptr(), L.stream())
        return (idx, val, hsrc), 1.0 + N0 / N, ovf
    # materialising path, user-chunked so that fval + fhead stay below 8 GB: the table is re-streamed per chunk
    ldn = (N + 3) // 4 * 4
    chunk = max(1, min(B, int((8 << 30) // max(1, ldn * 5))))
    passes = 0
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        fval = torch.empty((nb, ldn), dtype=torch.float32, device=dev)
        fhead = torch.empty((nb, ldn), dtype=torch.uint8, device=dev)
        fold_gemm(b0, b1, N, fval, fhead, ldn)
        L.call("b200rec_topk_select", fval.data_ptr(), fhead.data_ptr(), nb, N, ldn, K, None, None, rank, W,
               idx[b0:b1].data_ptr(), val[b0:b1].data_ptr(), hsrc[b0:b1].data_ptr(), L.stream())
        passes += 1
    return (idx, val, hsrc), float(passes), None


def run(args, rank, world, local_rank, pk):
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, K, Be = args.sweep_dim, 200, args.eval_users
    items = [int(x) for x in args.sweep_items.split(",") if x]
    heads = [int(x) for x in args.sweep_heads.split(",") if x]
    free_b = torch.cuda.mem_get_info(dev)[0]
    for N in items:
        n_local = (N - rank + world - 1) // world
        if n_local * D * 2 > 0.8 * free_b or n_local >= (1 << 27) + (1 << 26):
            if rank == 0:
                print(json.dumps({"metric": "eval_users_per_sec", "skipped": f"N={N} needs {n_local * D * 2 / 1e9:.1f} GB "
                                  f"per GPU at {world} GPU(s): row-shard over more GPUs", "items": N, "n_gpus": world}))
            continue
        table = _unit_rows(n_local, D, 7 + rank, dev)               # this rank's rows id % W == rank
        for H in heads:
            gen = torch.Generator(device=dev).manual_seed(100 + rank)
            U_loc = torch.randn((Be, H, D), generator=gen, device=dev)
            U_loc = (U_loc / U_loc.norm(dim=-1, keepdim=True)).to(torch.bfloat16)

            def one():
                U = U_loc
                if world > 1:
                    parts = [torch.empty_like(U_loc) for _ in range(world)]
                    dist.all_gather(parts, U_loc)
                    U = torch.cat(parts, dim=0)
                (idx, val, hs), passes, ovf = topk_point(table, U, K, rank, world)
                if world > 1:
                    outs = []
                    for t in (val, idx, hs):
                        r = torch.empty_like(t)
                        dist.all_to_all_single(r, t.contiguous())
                        outs.append(r.view(world, Be, K))
                    v, i, h = outs
                    idx, val, hs = parallel.merge_topk(list(v), list(i), list(h), K)
                return idx, passes, ovf

            for _ in range(2):
                idx, passes, ovf = one()
            torch.cuda.synchronize()
            overflow = bool(ovf is not None and int(ovf.item()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 if N >= 10_000_000 else 5
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            if rank == 0:
                users = Be * world
                flops = 2.0 * users * H * N * D                     # useful scoring FLOPs (no head padding counted)
                tbytes = N * D * 2.0                                # one stream of the bf16 table over all GPUs
                print(json.dumps({
                    "metric": "eval_users_per_sec", "value": users / (ms / 1e3), "unit": "users/s", "n_gpus": world,
                    "items": N, "heads": H, "dim": D, "users_per_batch": users, "K": K, "ms_per_batch": ms,
                    "table_streams_per_batch": passes, "candidate_overflow": overflow,
                    "useful_tflops": flops / (ms / 1e3) / 1e12,
                    "frac_of_tensor_peak": flops / (ms / 1e3) / 1e12 / (world * pk.get("bf16_tflops_sustained", 1370.0)),
                    "table_gbs": tbytes * passes / (ms / 1e3) / 1e9,
                    "frac_of_hbm_peak": tbytes * passes / (ms / 1e3) / 1e9 / (world * pk.get("hbm_gbs", 6446.0)),
                    "bound": "tensor" if users * H >= 212 else "hbm",
                    "sharding": "item rows id % W + cross-GPU top-K merge" if world > 1 else "none",
                    "data": "synthetic unit rows (seeded), random unit user heads"}), flush=True)
        del table
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
