#!/usr/bin/env python
"""HSTU multi-head train (+ eval) throughput on synthetic data of BASELINE.json's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config B] [--impl reference]

Own arm: one process per GPU (torchrun for N > 1).  A step = forward + backward + fused AdamW over
one synthetic batch of the named config (default B = HSTU-Pixel8M-prior, BASELINE.json configs[1]).
`value` = samples/s with the batch resident in HBM; `e2e` = the same step driven from pinned HOST
buffers (H2D of the batch and D2H of the loss inside the timed region).  `roofline` is for the
dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of every GEMM launch in the timed region /
their CUDA-event durations.  `cpu_baseline` / `--impl reference`: the reference algorithm on the
host cores (the unmodified reference when /root/reference is mounted, else the oracle port).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


WORKLOADS = {"A": "HSTU-Pixel8M-base-small (A)", "A2": "HSTU-Pixel8M-base-small 2 attn heads (A2)",
             "B": "HSTU-Pixel8M-prior (B)", "C": "HSTU-MerRec-prior (C)", "D": "HSTU-EBNerd-prior-mult (D)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dense-table-update", action="store_true",
                    help="touch every table row every step (the dense pass) instead of the lazy exact update")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="B")
    ap.add_argument("--impl", default="b200rec", choices=["b200rec", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch (0 = config value)")
    ap.add_argument("--layers", type=int, default=0, help="override n_layers (debug only; marks the line invalid)")
    ap.add_argument("--cpu-batch", type=int, default=4, help="samples per CPU-baseline step")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--eval-users", type=int, default=256)
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the CUDA-graph step")
    ap.add_argument("--replicate-table", action="store_true", help="N>1: keep the item table replicated")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel time table of one step (torch.profiler)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def train_flops_per_sample(cfg, valid_frac=1.0):
    """SURVEY §8(d) forward FLOPs per sample x3 (dedup counts, all rows valid upper bound)."""
    D, Lc, P = cfg["hstu_embedding_size"], cfg["MAX_ITEM_LIST_LENGTH"], cfg["pred_len"]
    NL, S, C = cfg["n_layers"], cfg["num_segment_head"], cfg["num_prior_head"]
    H = S + C if cfg["head_interaction"] == "additive" else S * C
    nneg = cfg["num_negatives"]
    body = NL * (10 * D * D * Lc + 4 * Lc * Lc * D)
    heads = H * 2 * D * D * Lc * (1 if cfg["medusa_num_layers"] else 0)
    by_cat = bool(cfg["neg_sample_by_cat"]) and cfg["loss"] == "prior"
    sets = (C if by_cat else 0) + (1 if (not by_cat or cfg["head_interaction"] == "additive") else 0)
    q_rows = H * Lc if cfg["loss"] == "prior" else (P // (P // S if cfg["medusa_num_layers"] else P)) * Lc
    nce_q = 2 * D * nneg * q_rows
    fix = 2 * D * nneg * (Lc + P) * sets
    return 3 * (body + heads + nce_q) + fix


def make_cfg(args):
    from b200rec import synth
    over = {}
    if args.batch:
        over["train_batch_size"] = args.batch
        # keep n negatives per sample as in the named config
        base = synth.PRESETS[args.config]
        over["num_negatives"] = base["num_negatives"] // base["train_batch_size"] * args.batch
    if args.layers:
        over["n_layers"] = args.layers
    return synth.make_config(args.config, **over)


# ------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(cfg, batch_size, steps, warmup=1):
    """Times the reference algorithm on the host cores (fwd + bwd + dense torch AdamW)."""
    from b200rec import synth
    from oracle import ref_harness as rh
    from oracle import hstu_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_per = cfg["num_negatives"] // cfg["train_batch_size"]
    c = synth.Config(cfg)
    c["train_batch_size"] = batch_size
    c["num_negatives"] = n_per * batch_size
    dl = synth.make_dataload(c)
    batch = synth.make_train_batch(c, seed=1)
    if rh.available():
        kind = "reference"
        model = rh.build_reference_model(dict(c), c["item_num"], dl.category_counts, dl.category_to_int)
        model.train()  # training mode (dropout at the preset's rate), like the GPU arm
        params = [p for p in model.parameters() if p.requires_grad]
        step_fn = lambda: model(batch)["loss"]
    else:
        kind = "port"
        from b200rec.hstu import HSTU
        torch.manual_seed(2020)
        host = HSTU(c, dl, compute_dtype=torch.float32)          # parameter container only (CPU, no kernels)
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "_rel_attn_bias" not in k)
              for k, v in host.state_dict().items()}
        oracle = orc.OracleHSTU(c, sd, dl.category_counts, dl.category_to_int)
        params = [v for v in sd.values() if v.requires_grad]
        step_fn = lambda: oracle.forward(batch)["loss"]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = step_fn()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": batch_size / sec, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"{steps} steps of batch {batch_size} (config {cfg['name']} at reduced batch, "
                      f"{n_per} negatives/sample/set, fp32, dense AdamW)", "ms_per_step": sec * 1e3}


# dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM launches of one training step (ncu, round 1)
GEMM_DRAM_BYTES_PER_STEP = {"B": 14.4e9}


# ------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = make_cfg(args)
    base = {"metric": "train_samples_per_sec", "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic", "config": {
                "workload": f"{WORKLOADS.get(cfg['name'], cfg['name'])}: {cfg['n_layers']} blocks D={cfg['hstu_embedding_size']} "
                            f"L={cfg['MAX_ITEM_LIST_LENGTH']} P={cfg['pred_len']} heads={cfg['num_segment_head']}+"
                            f"{cfg['num_prior_head']} {cfg['head_interaction']} negatives={cfg['num_negatives']}/set "
                            f"items={cfg['item_num']}",
                "per_gpu_batch": cfg["train_batch_size"], "global_batch": cfg["train_batch_size"] * world,
                "parallelism": f"dp{world}" + ("+row-sharded-table(a2a)" if world > 1 and not args.replicate_table else ""), "l2": "working set >> 126 MB L2 (activations + 1.8 GB table); no flush",
                "negatives": "re-drawn every step (64 pre-drawn sets rotated; by category where the config does)",
                "table_update": "dense pass over all rows" if args.dense_table_update else
                                "lazy exact AdamW (rows brought up to date when next read)",
                "step": ("eager" if (args.no_graph or args.profile or (world > 1 and args.replicate_table)) else
                         "cuda-graph replay per 128-token bucket" if world == 1 else
                         "eager id exchange + row fetch, cuda-graph fwd/bwd per 128-token bucket, eager all-reduce / "
                         "gradient-row push / AdamW")}}
    if args.impl == "reference":
        if rank != 0:
            return
        base["config"]["step"] = "cpu eager (reference algorithm on host cores)"
        r = cpu_reference_run(cfg, args.cpu_batch, max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
        line = dict(base)
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"], "dtype": "f32",
                     "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                     "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line))
        return

    from b200rec import synth, _lib as L
    from b200rec.hstu import HSTU
    from b200rec.optim import FusedAdamW
    from b200rec import parallel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    cfg["sparse_embedding_grad"] = True
    dl = synth.make_dataload(cfg)
    torch.manual_seed(2020)
    model = HSTU(cfg, dl, compute_dtype=dtype).to(dev).train()  # training mode: Philox dropout at the preset's rate
    use_graph = not args.no_graph and not args.profile and not (world > 1 and args.replicate_table)
    # lazy_table: exact dense-equivalent AdamW on the item table, rows nobody reads are updated when next read
    opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, device_step=use_graph and world == 1,
                     lazy_table=not args.dense_table_update)
    if world > 1 and not args.replicate_table:
        model.shard_item_table()          # rows id % W == rank; lookups / gradient rows by all-to-all
    dp = parallel.DataParallel(model, opt) if world > 1 else None
    item_tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
    n_batches = 4
    host_batches = [tuple(t.pin_memory() for t in synth.make_train_batch(cfg, seed=10 + i, rank=rank, world_size=world,
                                                                         item_tags=item_tags))
                    for i in range(n_batches)]
    dev_batches = [tuple(t.to(dev) for t in b) for b in host_batches]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_batches[0])

    Lc = cfg["MAX_ITEM_LIST_LENGTH"]
    n_tok = [int(b[2][:, :Lc].sum()) for b in host_batches]     # host metadata (the collate fn knows it)
    stepper = None
    if use_graph:
        from b200rec.graphed import GraphedTrainStep, GraphedShardedStep
        stepper = (GraphedTrainStep if world == 1 else GraphedShardedStep)(model, opt, dev_batches[0], bucket=128)

    def eager_step(batch):
        opt.zero_grad()
        out = model(batch)
        out["loss"].backward()
        if dp is not None:
            dp.sync_gradients()
        opt.step()
        return out["loss"]

    # Fresh negatives EVERY step (like the reference's sampler, trainset.py:126-137): the four synthetic batches
    # only fix the sequences / token counts; re-using their negative ids would let the lazy table update skip the
    # rows a real run keeps touching.  64 independent negative sets are drawn before the timed region (by
    # category where the config samples by category) and rotated in, one 0.6 MB copy per step.
    N_items, n_neg_sets = cfg["item_num"], 64
    by_cat = bool(cfg["neg_sample_by_cat"]) and cfg["loss"] == "prior" and cfg["category_by"] == "item"
    gen = torch.Generator().manual_seed(555 + rank)
    pools = [torch.nonzero(item_tags[1:, c]).flatten().add_(1) for c in range(cfg["eval_num_cats"])] if by_cat else []
    shape = tuple(host_batches[0][1].shape)
    host_negs = []
    for _ in range(n_neg_sets):
        neg = torch.randint(1, N_items, shape, generator=gen)
        for c, pool in enumerate(pools):                            # set c: uniform over the items of category c
            neg[:, c] = pool[torch.randint(0, pool.numel(), tuple(neg[:, c].shape), generator=gen)]
        host_negs.append(neg.pin_memory())
    dev_negs = [t.to(dev) for t in host_negs]
    neg_ctr = [0]

    def fresh_negatives(neg):
        src = (dev_negs if neg.is_cuda else host_negs)[neg_ctr[0] % n_neg_sets]
        neg_ctr[0] += 1
        neg.copy_(src, non_blocking=True)

    def step(batch, i=None):
        fresh_negatives(batch[1])
        if stepper is not None:
            return stepper(batch, n_tok[i])["loss"]
        return eager_step(batch)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, n_batches if use_graph else 0)):   # graph mode: capture every bucket first
        step(dev_batches[i % n_batches], i % n_batches)
    barrier()
    if (args.profile or os.environ.get("B200REC_PROFILE_STEP")) and (rank == 0 or world > 1):
        from torch.profiler import profile, ProfilerActivity
        acts = [ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if os.environ.get("B200REC_PROFILE_STEP") else [])
        with profile(activities=acts) as prof:
            for i in range(2):
                step(dev_batches[i], i)
            torch.cuda.synchronize()
        if rank == 0 and os.environ.get("B200REC_PROFILE_STEP"):
            prof.export_chrome_trace("gpurun_out/trace_rank0.json")
    if args.profile and rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60),
              file=sys.stderr)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- timed region 1: batch resident in HBM
    launches0 = L.launches
    if not use_graph:
        L.gemm_timing = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(dev_batches[i % n_batches], i % n_batches)
    host_ms = (time.perf_counter() - t_host0) * 1e3      # host enqueue time (no sync): > GPU time means launch-bound
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launches - launches0
    gemm_events = L.gemm_timing
    L.gemm_timing = None
    # ---- timed region 2: end to end from pinned host buffers
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    last = None
    for i in range(args.steps):
        j = i % n_batches
        if use_graph:
            b = host_batches[j]                  # H2D of the pinned batch happens inside the stepper
            last = float(step(b, j).item())      # D2H read of the loss
        else:
            fresh_negatives(host_batches[j][1])
            b = tuple(t.to(dev, non_blocking=True) for t in host_batches[j])
            last = float(eager_step(b).item())
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if stepper is not None and hasattr(stepper, "flush"):
        stepper.flush()                          # pending dense update of the sharded step (overlapped all-reduce)
    if use_graph:
        # kernels inside a replayed graph cannot be bracketed by events: time every GEMM launch of the same
        # step in an instrumented eager pass right after the timed regions (same kernels, same shapes)
        L.gemm_timing = []
        for i in range(args.steps):
            eager_step(dev_batches[i % n_batches])
        torch.cuda.synchronize()
        gemm_events = L.gemm_timing
        L.gemm_timing = None
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    eval_res = None
    if not args.no_eval:
        # every rank takes part: with a row-sharded table the eval path exchanges users and merges top-K lists
        try:
            eval_res = eval_bench(cfg, model, item_tags, dev, args.eval_users, rank, world)
        except Exception as ex:  # eval is a secondary number; never lose the train line
            if world > 1:
                raise
            eval_res = {"error": repr(ex)[:200]}
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    B = cfg["train_batch_size"]
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    pk, pk_src = peaks()
    gflops = sum(f for (_, _, f) in gemm_events)
    gms = sum(s.elapsed_time(e) for (s, e, _) in gemm_events)
    ach = gflops / (gms / 1e3) / 1e12 if gms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    line = dict(base)
    line.update({
        "value": value, "ms_per_step": ms / args.steps, "dtype": args.dtype,
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms / args.steps,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05, all GEMM launches of the step)",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                     "traffic": GEMM_DRAM_BYTES_PER_STEP.get(cfg["name"]) if world == 1 else None,
                     "traffic_note": "dram read+write bytes of all GEMM launches of one step, ncu launch list "
                                     "profiles/r01_launches_dram_metrics.csv (profiles/r01_kernel_evidence.md)",
                     "peak_source": pk_src + ", sustained bf16",
                     "gemm_ms_per_step": gms / args.steps, "gemm_share_of_step": gms / ms if ms else None,
                     "timing": ("instrumented eager pass after the timed region (the timed region replays CUDA graphs)"
                                if use_graph else "CUDA events around every GEMM launch inside the timed region"),
                     "gemm_launches_per_step": len(gemm_events) / args.steps},
        "model_flops_utilisation": train_flops_per_sample(cfg) * value / 1e12 / peak if peak else None,
        "clocks": sampler.summary() if sampler else None,
        "loss_last": last,
    })
    if args.layers:
        line["invalid"] = "n_layers overridden (debug run)"
    if eval_res is not None:
        line["eval"] = eval_res
    if not args.no_cpu:
        r = cpu_reference_run(cfg, args.cpu_batch, args.cpu_steps)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def eval_bench(cfg, model, item_tags, dev, users, rank=0, world=1):
    """Secondary metric: eval users/s = predict + masks + cross-head merge + top-200 + hit matrix.  `users` per
    rank; with a row-sharded table every rank scores all ranks' users against its rows and the per-shard top-K
    lists are merged."""
    from b200rec import synth
    from b200rec.evaluator import Collector
    ev = synth.make_eval_batch(cfg, seed=3 + rank, batch_size=users, item_tags=item_tags)
    C = cfg["eval_num_cats"]
    tags = item_tags.t().contiguous().to(dev) if cfg["category_by"] == "item" else \
        torch.ones(C, cfg["item_num"], dtype=torch.bool, device=dev)
    if model.sharded_table is not None:
        tags = tags[:, rank::world].contiguous()
    feat = model.compute_item_all()
    seq, tt = ev["item_seq"].to(dev), ev["target_tags"].to(dev)
    hist = (ev["history_index"][0].to(dev), ev["history_index"][1].to(dev))
    tgt = ev["item_target"].to(dev)
    coll = Collector(cfg)

    def one():
        top = model.predict_topk(seq, feat, tags, tt, history_index=hist, K=max(cfg["topk"]))
        coll.eval_batch_collect(None, ev["positive_u"], tgt, None, topk=top)

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    if os.environ.get("B200REC_PROFILE_EVAL"):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=60), file=sys.stderr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    return {"metric": "eval_users_per_sec", "value": users * world / (ms / 1e3), "unit": "users/s",
            "users_per_batch": users * world, "items": cfg["item_num"], "heads": model.medusa_num_heads,
            "K": max(cfg["topk"]), "ms_per_batch": ms,
            "sharding": "item rows id % W, cross-GPU top-K merge" if model.sharded_table is not None else "none"}


if __name__ == "__main__":
    main()
