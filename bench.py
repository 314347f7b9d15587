#!/usr/bin/env python
"""HSTU multi-head train (+ eval) throughput on synthetic data of BASELINE.json's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config B|C|D|A] [--impl reference] [--eval-sweep]

Own arm: one process per GPU (torchrun for N > 1).  A step = forward + backward + fused AdamW over one synthetic
batch of the named config (default B = HSTU-Pixel8M-prior, BASELINE.json configs[1]).
  value      samples/s over exactly K steps with the batch resident in HBM (CUDA events, max over ranks)
  e2e        the same K steps driven from pinned HOST buffers (H2D of the batch, D2H of the loss inside the region)
  sustained  the same step repeated for >= --sustained-seconds (default 6 s): the power / thermal steady state
  roofline   the dominant kernel family (tcgen05 GEMMs): algorithmic FLOPs of every GEMM launch of a step / their
             durations, read from CUDA events captured INSIDE the replayed CUDA graph (external event-record nodes of an
             instrumented copy of the step graph, replayed right after the sustained region, i.e. in the same clock
             state); the peak is chosen from the SM clock sampled during those replays (burst when un-capped)
  flush_ms   the one pass that brings every lazily-updated table row up to date (paid before eval / checkpoint)
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference model on the host cores (oracle/_ref, written by
oracle/build_ref.py; the restated oracle `kind: port` only if that copy is absent), at a reduced batch which the line states.
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
    ap.add_argument("--items", type=int, default=0, help="override the catalogue size (0 = config value)")
    ap.add_argument("--layers", type=int, default=0, help="override n_layers (debug only; marks the line invalid)")
    ap.add_argument("--cpu-batch", type=int, default=4, help="samples per CPU-baseline step")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--eval-users", type=int, default=256)
    ap.add_argument("--sustained-seconds", type=float, default=6.0, help="0 disables the sustained sub-record")
    ap.add_argument("--instrument-steps", type=int, default=8, help="replays of the event-instrumented graph")
    ap.add_argument("--gemm-table", default="", help="write a per-shape table of the timed GEMM launches to this file")
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the CUDA-graph step")
    ap.add_argument("--replicate-table", action="store_true", help="N>1: keep the item table replicated")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel time table of one step (torch.profiler)")
    ap.add_argument("--eval-sweep", action="store_true",
                    help="full-catalogue eval scoring + top-K sweep (BASELINE configs[4]) instead of the training bench")
    ap.add_argument("--sweep-items", default="1000000,5000000,10000000,50000000")
    ap.add_argument("--sweep-heads", default="1,7,12")
    ap.add_argument("--sweep-dim", type=int, default=256)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, \
        "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.marks = {}

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append((time.perf_counter(), [x.strip() for x in out.strip().split(",")]))
            except Exception:
                pass
            time.sleep(0.15)

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def summary(self, t0=None, t1=None):
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        ok = [r for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        sm = sorted(int(float(r[0])) for r in ok)
        mx = [int(float(r[1])) for r in ok if r[1].replace(".", "").isdigit()]
        pw = sorted(float(r[2]) for r in ok if r[2].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in ok)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": pw[len(pw) // 2] if pw else None, "reasons": reasons, "samples": len(ok)}


# ------------------------------------------------------------------------------------- FLOP model
def train_flops_upper(cfg):
    """Forward FLOPs per sample x3 with every row valid (dedup counts): the printed upper bound."""
    D, Lc, P = cfg["hstu_embedding_size"], cfg["MAX_ITEM_LIST_LENGTH"], cfg["pred_len"]
    return _flops(cfg, float(Lc), float(Lc * Lc), float(Lc + P))


def _flops(cfg, n_ctx, n_ctx_sq, n_tgt):
    """n_ctx = valid context tokens, n_ctx_sq = sum of squared sequence lengths, n_tgt = valid target positions
    (per sample averages).  body 10 D^2 l + 4 l^2 D per block, heads 2 D^2 l per head, NCE queries 2 D Nneg per
    (job, token), false-negative filter 2 D Nneg per (target position, negative set); x3 for forward + backward
    except the filter (no gradient)."""
    D, P = cfg["hstu_embedding_size"], cfg["pred_len"]
    NL, S, C = cfg["n_layers"], cfg["num_segment_head"], cfg["num_prior_head"]
    H = S + C if cfg["head_interaction"] == "additive" else S * C
    nneg = cfg["num_negatives"]
    body = NL * (10 * D * D * n_ctx + 4 * n_ctx_sq * D)
    heads = H * 2 * D * D * n_ctx * (1 if cfg["medusa_num_layers"] else 0)
    by_cat = bool(cfg["neg_sample_by_cat"]) and cfg["loss"] == "prior"
    sets = (C if by_cat else 0) + (1 if (not by_cat or cfg["head_interaction"] == "additive") else 0)
    if cfg["loss"] == "prior":
        jobs = (S if cfg["head_interaction"] == "additive" else 0) + C * (1 if cfg["head_interaction"] == "additive" else S)
    else:
        jobs = P // (P // S if cfg["medusa_num_layers"] else P)
    nce_q = 2 * D * nneg * jobs * n_ctx
    fix = 2 * D * nneg * n_tgt * sets
    return 3 * (body + heads + nce_q) + fix


def train_flops_valid(cfg, host_batches):
    """Same model on the ACTUAL masks of the synthetic batches (valid tokens only), per sample."""
    Lc = cfg["MAX_ITEM_LIST_LENGTH"]
    tot, n = 0.0, 0
    for b in host_batches:
        m = b[2].bool()
        ell = m[:, :Lc].sum(1).double()
        B = m.shape[0]
        tot += _flops(cfg, float(ell.mean()), float((ell * ell).mean()), float(m.sum(1).double().mean())) * B
        n += B
    return tot / n


def make_cfg(args):
    from b200rec import synth
    over = {}
    if args.batch:
        over["train_batch_size"] = args.batch
        base = synth.PRESETS[args.config]
        over["num_negatives"] = base["num_negatives"] // base["train_batch_size"] * args.batch   # same n per sample
    if args.layers:
        over["n_layers"] = args.layers
    if args.items:
        over["item_num"] = args.items
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
    c["item_num"] = min(cfg["item_num"], 450000)          # dense fp32 table + dense AdamW state on the host
    dl = synth.make_dataload(c)
    batch = synth.make_train_batch(c, seed=1)
    if rh.available():
        kind = "reference"
        model = rh.build_reference_model(dict(c), c["item_num"], dl.category_counts, dl.category_to_int)
        model.train()  # training mode (dropout at the preset's rate), like the GPU arm
        params = [p for p in model.parameters() if p.requires_grad]
        step_fn = lambda: model(batch)["loss"]
        src = "unmodified reference modules from " + ("/root/reference" if rh.REFERENCE_CODE.startswith("/root/reference")
                                                      else "oracle/_ref (verbatim copy, oracle/build_ref.py)")
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
        src = "oracle/hstu_oracle.py restatement (oracle/_ref absent)"
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
            "sample": f"{steps} steps of batch {batch_size} (config {cfg['name']} at reduced batch: {n_per} negatives/"
                      f"sample/set = {n_per * batch_size} per set, {c['item_num']} items, fp32, dense torch AdamW; {src})",
            "ms_per_step": sec * 1e3, "ran": {"per_gpu_batch": batch_size, "negatives_per_set": n_per * batch_size,
                                              "items": c["item_num"]}}


def gemm_traffic_record(name):
    """dram read+write bytes of all GEMM launches of one training step, from the committed ncu summary of this
    round (profiles/r02_gemm_traffic.json, written by scripts/ncu_traffic.py from the ncu launch list)."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            if name in d:
                return d[name].get("gemm_dram_bytes_per_step"), d[name].get("source")
        except Exception:
            pass
    return None, None


def workload_text(cfg):
    return (f"{WORKLOADS.get(cfg['name'], cfg['name'])}: {cfg['n_layers']} blocks D={cfg['hstu_embedding_size']} "
            f"L={cfg['MAX_ITEM_LIST_LENGTH']} P={cfg['pred_len']} heads={cfg['num_segment_head']}+"
            f"{cfg['num_prior_head']} {cfg['head_interaction']} negatives={cfg['num_negatives']}/set "
            f"items={cfg['item_num']}")


# ------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.eval_sweep:
        return eval_sweep(args, rank, world, local_rank)
    cfg = make_cfg(args)
    graph_mode = not (args.no_graph or args.profile or (world > 1 and args.replicate_table))
    base = {"metric": "train_samples_per_sec", "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic", "config": {
                "workload": workload_text(cfg),
                "per_gpu_batch": cfg["train_batch_size"], "global_batch": cfg["train_batch_size"] * world,
                "parallelism": f"dp{world}" + ("+row-sharded-table" if world > 1 and not args.replicate_table else ""),
                "l2": "working set >> 126 MB L2 (activations + item table); no flush",
                "negatives": "re-drawn every step (64 pre-drawn sets rotated; by category where the config does)",
                "table_update": "dense pass over all rows" if args.dense_table_update else
                                "lazy exact AdamW (rows brought up to date when next read)",
                "step": ("eager" if not graph_mode else
                         "cuda-graph replay per 128-token bucket" if world == 1 else
                         "device-side id exchange + row fetch, cuda-graph fwd/bwd per 128-token bucket, gradient-row "
                         "push / AdamW, dense all-reduce overlapped")}}
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(cfg, args.cpu_batch, max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
        line = dict(base)
        line["config"] = dict(base["config"])
        line["config"].update({"step": "cpu eager (reference model on host cores)", "per_gpu_batch": r["ran"]["per_gpu_batch"],
                               "global_batch": r["ran"]["per_gpu_batch"],
                               "workload": workload_text(cfg) + f" -- RUN AT batch {r['ran']['per_gpu_batch']}, "
                                           f"{r['ran']['negatives_per_set']} negatives/set, {r['ran']['items']} items",
                               "same_config": False, "parallelism": f"cpu x{r['cores']} threads",
                               "table_update": "dense torch.optim.AdamW", "negatives": "one fixed synthetic batch"})
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"], "dtype": "f32",
                     "same_config": False,
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
    use_graph = graph_mode
    # lazy_table: exact dense-equivalent AdamW on the item table, rows nobody reads are updated when next read
    opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, device_step=use_graph and world == 1,
                     lazy_table=not args.dense_table_update)
    if world > 1 and not args.replicate_table:
        model.shard_item_table()          # rows id % W == rank; lookups / gradient rows travel over NVLink
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
    Stepper = None
    if use_graph:
        from b200rec.graphed import GraphedTrainStep, GraphedShardedStep
        Stepper = GraphedTrainStep if world == 1 else GraphedShardedStep
        stepper = Stepper(model, opt, dev_batches[0], bucket=128)

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
    # category where the config samples by category) and rotated in, one small copy per step.
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

    def step(batch, i=None, st=None):
        fresh_negatives(batch[1])
        if st is not None or stepper is not None:
            return (st or stepper)(batch, n_tok[i])["loss"]
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
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60),
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
    t_reg0 = time.perf_counter()
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
    t_reg1 = time.perf_counter()
    ms_e2e = f0.elapsed_time(f1)
    # ---- sustained region: the same device-resident step for >= sustained_seconds (power / thermal steady state)
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(ms / args.steps, 1e-3)) + 1)
        if world > 1:                                     # every rank must run the same number of steps
            t = torch.tensor([n_sus], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            n_sus = int(t.item())
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_sus0 = time.perf_counter()
        g0.record()
        for i in range(n_sus):
            step(dev_batches[i % n_batches], i % n_batches)
        g1.record()
        barrier()
        t_sus1 = time.perf_counter()
        sustained = {"steps": n_sus, "ms": g0.elapsed_time(g1)}
    # ---- GEMM launch times inside the replayed graph (instrumented copy of the step graph, same clock state)
    gemm_src = "CUDA events around every GEMM launch inside the timed region (eager step)"
    t_ins0 = t_ins1 = None
    if use_graph and args.instrument_steps > 0:
        if hasattr(stepper, "flush"):
            stepper.flush()                               # sharded step: no dense update may be pending across steppers
        ins = Stepper(model, opt, dev_batches[0], bucket=128, instrument=True)
        for i in range(4):                                # capture (state snapshotted / restored), then back to the
            step(dev_batches[0], 0, st=ins)               # loaded clock state after the capture pause
        torch.cuda.synchronize()
        gemm_events = []
        t_ins0 = time.perf_counter()
        for _ in range(args.instrument_steps):
            step(dev_batches[0], 0, st=ins)
            torch.cuda.synchronize()
            gemm_events += ins.gemm_times(n_tok[0])
        t_ins1 = time.perf_counter()
        if hasattr(ins, "flush"):
            ins.flush()
        n_gemm_steps = args.instrument_steps
        gemm_src = ("external CUDA event-record nodes around every GEMM launch inside the replayed step graph "
                    f"({args.instrument_steps} replays right after the sustained region)")
        del ins
    else:
        gemm_events = [(ev[0].elapsed_time(ev[1]), ev[2]) + tuple(ev[3:]) for ev in (gemm_events or [])]
        n_gemm_steps = args.steps
    if stepper is not None and hasattr(stepper, "flush"):
        stepper.flush()                          # pending dense update of the sharded step (overlapped all-reduce)
    # ---- lazy table: the catch-up pass over every row (paid before evaluation / checkpoints)
    flush_ms = None
    if opt.lazy_table:
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        h0.record()
        opt.flush()
        h1.record()
        torch.cuda.synchronize()
        flush_ms = h0.elapsed_time(h1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    t = torch.tensor([ms, ms_e2e, sustained["ms"] if sustained else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e, ms_sus = float(t[0]), float(t[1]), float(t[2])
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
    gflops = sum(ev[1] for ev in gemm_events)
    gms = sum(ev[0] for ev in gemm_events)
    if args.gemm_table:
        # per-shape table of the in-graph GEMM launches (M, N, K, groups, epilogue, a_major, b_major)
        tab = {}
        for ev in gemm_events:
            r = tab.setdefault(ev[2] if len(ev) > 2 else None, [0, 0.0, 0.0])
            r[0] += 1; r[1] += ev[0]; r[2] += ev[1]
        with open(args.gemm_table, "w") as fh:
            fh.write("M N K groups epi a_major b_major | launches/step us/launch TF/s share_of_gemm\n")
            for k, r in sorted(tab.items(), key=lambda kv: -kv[1][1]):
                fh.write("%s | %.1f %.1f %.0f %.3f\n" % (" ".join(map(str, k)) if k else "?", r[0] / n_gemm_steps,
                                                      r[1] / r[0] * 1e3, r[2] / (r[1] / 1e3) / 1e12, r[1] / gms))
    ach = gflops / (gms / 1e3) / 1e12 if gms > 0 else 0.0
    clk_gemm = None
    if sampler:
        # the instrumented replays take ~0.1 s (shorter than one nvidia-smi poll): they run back to back with the
        # sustained region, whose samples describe the same clock / power state
        clk_gemm = sampler.summary(t_ins0, t_ins1) if t_ins0 else sampler.summary(t_reg0, t_reg1)
        if not clk_gemm["samples"] and sustained:
            clk_gemm = sampler.summary(t_sus0, t_ins1 if t_ins1 else t_sus1)
            clk_gemm["window"] = "sustained region + instrumented replays"
    sm_max = pk.get("sm_max_mhz") or (clk_gemm or {}).get("sm_max_mhz") or 1965.0
    uncapped = bool(clk_gemm and clk_gemm["sm_mhz"] and clk_gemm["sm_mhz"] >= 0.93 * sm_max and
                    "sw_power_cap" not in clk_gemm["reasons"])
    peak_burst, peak_sus = pk.get("bf16_tflops"), pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    peak = peak_burst if uncapped else peak_sus
    traffic, traffic_src = gemm_traffic_record(cfg["name"]) if world == 1 else (None, None)
    flops_valid = train_flops_valid(cfg, host_batches)
    line = dict(base)
    line.update({
        "value": value, "ms_per_step": ms / args.steps, "dtype": args.dtype,
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms / args.steps,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel / gemm_tc_grouped_kernel (tcgen05; all GEMM launches of a step)",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak else None,
                     "frac_of_burst": ach / peak_burst if peak_burst else None,
                     "frac_of_sustained": ach / peak_sus if peak_sus else None,
                     "peak_choice": ("burst: SM clock %s MHz of %s during the GEMM timing, no power cap" if uncapped else
                                     "sustained: SM clock %s MHz of %s during the GEMM timing") %
                                    ((clk_gemm or {}).get("sm_mhz"), sm_max),
                     "peak_source": pk_src, "traffic": traffic, "traffic_source": traffic_src,
                     "gemm_ms_per_step": gms / n_gemm_steps, "gemm_share_of_step": (gms / n_gemm_steps) / (ms / args.steps),
                     "gemm_flops_per_step": gflops / n_gemm_steps, "timing": gemm_src,
                     "gemm_launches_per_step": len(gemm_events) / n_gemm_steps,
                     "clocks_during_gemm_timing": clk_gemm},
        # per GPU: (valid-token FLOPs of this rank's batch / step time) / peak -- identical on every rank, so no /N needed
        "model_flops_utilisation": flops_valid * B / (ms / args.steps / 1e3) / 1e12 / peak_sus if peak_sus else None,
        "model_flops_note": "SURVEY 8(d) FLOP model on the ACTUAL valid tokens of the synthetic batches, per GPU, over the "
                            "sustained bf16 peak; all-rows-valid upper bound per sample: %.3e" % train_flops_upper(cfg),
        "clocks": sampler.summary(t_reg0, t_reg1) if sampler else None,
        "loss_last": last,
    })
    if sustained:
        line["sustained"] = {"value": B * world * sustained["steps"] / (ms_sus / 1e3), "unit": "samples/s",
                             "steps": sustained["steps"], "seconds": ms_sus / 1e3,
                             "ms_per_step": ms_sus / sustained["steps"],
                             "clocks": sampler.summary(t_sus0, t_sus1) if sampler else None}
    if flush_ms is not None:
        line["flush_ms"] = flush_ms
        line["flush_note"] = "one catch-up pass over all item rows (lazy exact AdamW), paid before eval / checkpoint, " \
                             "outside the per-step timed regions"
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
    n = 5
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


def eval_sweep(args, rank, world, local_rank):
    """BASELINE configs[4]: full-catalogue scoring + top-K over N synthetic items x H heads (one JSON line per point).
    The catalogue is random unit rows; with N GPUs it is row-sharded (`id % W`) and the per-shard lists
    are merged across GPUs."""
    from b200rec import evalsweep
    evalsweep.run(args, rank, world, local_rank, peaks()[0])


if __name__ == "__main__":
    main()
