"""Generates tests/golden/comirec_*.pt from the LIVE, UNMODIFIED reference ComiRec
(/root/reference/code/REC/model/IDNet/comirec.py; build container only).  Usage: python tests/golden/make_golden_comirec.py

Each fixture: config, reference-initialised state dict, one seeded train batch with the reference's loss / logging
scalars / every parameter gradient, one eval batch with predict() scores [B, K interests, N]."""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from b200rec import synth  # noqa: E402

TINY = dict(n_layers=2, n_heads=2, item_embedding_size=32, hstu_embedding_size=32, MAX_ITEM_LIST_LENGTH=12,
            train_batch_size=6, num_negatives=30, item_num=300, eval_batch_size=5, hidden_dropout_prob=0.0)
CASES = {
    "comirec_p1": dict(TINY, interest_num=4),
    "comirec_p4": dict(TINY, pred_len=4, eval_pred_len=4, interest_num=3, interest_hidden=20),
    # REMI (remi.py): routing regularisation + interest-aware hard negatives; the shipped remi.yaml values first
    "remi_p1": dict(TINY, model="REMI", interest_num=4, lambda_rr=100.0, beta_ihn=1.0, attention_net_bias=False,
                    interest_hidden_ratio=0.5),
    "remi_p3_beta4": dict(TINY, model="REMI", pred_len=3, eval_pred_len=3, interest_num=3, lambda_rr=10.0, beta_ihn=4.0,
                          attention_net_bias=True, interest_hidden=20),
    "remi_p2_beta0": dict(TINY, model="REMI", pred_len=2, eval_pred_len=2, interest_num=4, lambda_rr=100.0, beta_ihn=0.0),
}


def _assert_margin(ref, batch, cfg):
    """Top-2 gap of the readout similarities on the valid (token, offset) pairs: the selection must not hinge on rounding."""
    from oracle.comirec_oracle import OracleComiRec
    o = OracleComiRec(dict(cfg), {k: v.detach().clone() for k, v in ref.state_dict().items()})
    items, _, mask, _ = batch
    L, P = o.L, o.P
    m = mask.bool()
    E = o.embed(items)
    y = o.body(E[:, :L] + o.p["position_embedding.weight"][:L], m[:, :L])
    u = o.interests(y, m[:, :L])
    widx = torch.arange(L)[None, :] + 1 + torch.arange(P)[:, None]
    tok = m[:, None, :L] & m[:, widx]
    sim = torch.einsum("blkd,bpld->bplk", u, E[:, widx])[tok]
    prob = torch.softmax(sim, dim=-1)
    top2 = prob.topk(2, dim=-1).values
    exact_tie = (sim == sim[:, :1]).all(dim=-1)       # a sequence's first token: one-element prefix, all interests equal
    gap = float((top2[:, 0] - top2[:, 1])[~exact_tie].min())
    assert gap > 1e-6, f"readout arg-max margin {gap:.2e} too small for a portable fixture"
    print(f"   readout margin (min top-2 softmax gap): {gap:.3e}, selections used: {sorted(set(prob.argmax(-1).tolist()))}")


def run_case(name, over):
    cfg = synth.make_config("A2", **over)
    rh.load()
    if over.get("model") == "REMI":
        from REC.model.IDNet.remi import REMI as ComiRec
    else:
        from REC.model.IDNet.comirec import ComiRec
    torch.manual_seed(2020)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ComiRec(rh.RefConfig(dict(cfg)), rh.RefDataload(cfg["item_num"]))
    ref.eval()
    # A fresh init (std 0.02) makes the K interests identical to ~1e-6: the reference's arg-max over softmax(similarity)
    # (comirec.py:287-291) is then decided by fp32 rounding of the softmax, which no other implementation (not even
    # torch's own CUDA softmax) reproduces.  Trained-scale weights separate the interests; the generator asserts the margin.
    with torch.no_grad():
        ref.attention_net[0].weight.mul_(60.0)
        ref.attention_net[3].weight.mul_(60.0)
        ref.item_embedding.weight.mul_(10.0)
    batch = synth.make_train_batch(cfg, seed=3, zipf=False)
    out = ref(batch)
    out["loss"].backward()
    _assert_margin(ref, batch, cfg)
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in ref.named_parameters()}
    logs = {k: float(v) for k, v in out.items()}
    ev = synth.make_eval_batch(cfg, seed=5)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        feat = ref.compute_item_all()
        scores, _, _, _ = ref.predict(ev["item_seq"], None, feat, None, None)
    fx = dict(name=name, cfg=dict(cfg), state_dict={k: v.clone() for k, v in ref.state_dict().items()},
              train_batch=batch, loss=float(out["loss"]), logs=logs, grads=grads, eval_batch=ev, item_feature=feat,
              scores=scores)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.pt")
    torch.save(fx, path)
    print(f"{name}: loss={fx['loss']:.6f} scores={tuple(scores.shape)} logs={sorted(logs)} -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")
    print("   params:", [(k, tuple(v.shape)) for k, v in ref.named_parameters() if "_hstu" not in k])


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, over in CASES.items():
        if not only or any(name.startswith(o) for o in only):
            run_case(name, over)
