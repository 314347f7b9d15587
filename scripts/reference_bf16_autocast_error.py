"""How much precision does the REFERENCE itself lose in its production dtype?  The reference trains under Lightning's
`bf16-mixed` (overall/ID.yaml:47-49): matmuls / linears run in bf16 autocast, the rest in fp32.  This script runs the
unmodified reference HSTU on CPU twice on the same batch — fp32, and under torch.autocast(cpu, bfloat16) — at the real
width / depth of a BASELINE config (small batch) and prints the per-tensor gradient cosine between the two.  It is the
yardstick for the bf16 tolerance of the CUDA path (tests/test_gpu_real_shapes.py, profiles/r02_bf16_parity.md).
Usage (build container, needs /root/reference or oracle/_ref): python scripts/reference_bf16_autocast_error.py B"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from b200rec import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "B"
over = dict(train_batch_size=4, item_num=100000, hidden_dropout_prob=0.0)
over["num_negatives"] = synth.PRESETS[name]["num_negatives"]
cfg = synth.make_config(name, **over)
dl = synth.make_dataload(cfg)
item_tags = synth.make_item_tags(cfg, torch.Generator().manual_seed(4242))
batch = synth.make_train_batch(cfg, seed=11, item_tags=item_tags)
torch.set_num_threads(os.cpu_count() or 1)
grads, losses = {}, {}
for mode in ("fp32", "bf16-mixed"):
    model = rh.build_reference_model(dict(cfg), cfg["item_num"], dl.category_counts, dl.category_to_int)
    model.eval()
    if mode == "fp32":
        out = model(batch)
    else:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = model(batch)
    out["loss"].float().backward()
    losses[mode] = float(out["loss"])
    grads[mode] = {k: p.grad.float().clone() for k, p in model.named_parameters() if p.grad is not None}
rep = {"config": name, "loss_fp32": losses["fp32"], "loss_bf16_mixed": losses["bf16-mixed"], "cosine": {}}
for k, g in grads["fp32"].items():
    if g.numel() < 2:
        continue
    a, b = g.flatten().double(), grads["bf16-mixed"][k].flatten().double()
    rep["cosine"][k] = float((a @ b) / (a.norm() * b.norm() + 1e-300))
worst = min(rep["cosine"].items(), key=lambda kv: kv[1])
rep["cosine_min"] = {"tensor": worst[0], "value": worst[1]}
print(json.dumps(rep, indent=1))
