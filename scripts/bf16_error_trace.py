"""Where does the bf16 production mode lose precision?  Runs the SAME batch through the fp32 verification mode and the
bf16 production mode of the CUDA path (config given on the command line, real width / depth, small batch) and prints the
relative error and cosine of the intermediates the model stashes in `HSTU._debug` (forward: embedding, body output, head
outputs, normalised queries; backward: d q_hat, d heads, d body output, d embedding rows).
Usage (GPU box): python scripts/bf16_error_trace.py B"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200rec import synth  # noqa: E402
from b200rec.hstu import HSTU  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "B"
over = dict(train_batch_size=4, item_num=100000, hidden_dropout_prob=0.0)
over["num_negatives"] = synth.PRESETS[name]["num_negatives"]
cfg = synth.make_config(name, **over)
cfg["sparse_embedding_grad"] = True
dl = synth.make_dataload(cfg)
torch.manual_seed(2020)
host = HSTU(cfg, dl, compute_dtype=torch.float32)
batch = tuple(t.cuda() for t in synth.make_train_batch(cfg, seed=11))
dbg = {}
for dt in (torch.float32, torch.bfloat16):
    m = HSTU(cfg, dl, compute_dtype=dt)
    m.load_state_dict(host.state_dict())
    m = m.cuda().eval()
    m._debug = {}
    out = m(batch)
    out["loss"].backward()
    dbg[dt] = m._debug
    dbg[dt]["loss"] = out["loss"].detach().reshape(1)
a, b = dbg[torch.float32], dbg[torch.bfloat16]
for k in ("loss", "x0", "y", "hd", "qhat", "that", "dqhat", "dthat", "d_hd", "dy", "dx0"):
    x, y = a[k].double().flatten(), b[k].double().flatten()
    rel = float((x - y).norm() / (x.norm() + 1e-300))
    cos = float((x @ y) / (x.norm() * y.norm() + 1e-300))
    print(f"{k:8s} rel_err {rel:.3e}  cos {cos:.6f}  |fp32| {float(x.norm()):.3e}")
for i, (p, q) in enumerate(zip(a["per_p"], b["per_p"])):
    print("job", i, "per_p rel", float((p - q).abs().max() / (p.abs().max() + 1e-30)))
