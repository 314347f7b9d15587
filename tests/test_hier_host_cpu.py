"""Host logic of the hierarchical decode heads and their options (hstu.py:444-483, 652-663) in the CPU tier: the tape the
forward records, the reverse walk, gradient accumulation over shared tensors / shared weights — against the oracle's
autograd.  The C-ABI calls are served by tests/cabi_cpu_shim.py (a restatement of include/b200rec.h on host memory); the
kernels themselves are checked on the GPU by test_gpu_model.py[hier_2x7 / hier_options]."""
import pytest
import torch

import cabi_cpu_shim as shim
from conftest import load_golden
from oracle import hstu_oracle as orc
from b200rec import synth
from b200rec.hstu import HSTU

OPTS = [
    dict(),
    dict(head_norm=True),
    dict(cat_bottleneck=True),
    dict(cat_bottleneck=True, cat_bottleneck_dim=8, head_norm=True),
    dict(share_seg_weights=True),
    dict(segment_embed=True),
    dict(head_norm=True, cat_bottleneck=True, share_seg_weights=True, segment_embed=True),
]


@pytest.mark.parametrize("opts", OPTS, ids=lambda o: "+".join(sorted(o)) or "plain")
@pytest.mark.parametrize("layers", [1, 2])
def test_hier_heads_forward_backward_match_oracle(opts, layers):
    cfg = synth.make_config("D", n_layers=1, n_heads=2, item_embedding_size=32, hstu_embedding_size=32,
                            MAX_ITEM_LIST_LENGTH=8, train_batch_size=3, num_negatives=8, item_num=100,
                            head_interaction="hierarchical", num_segment_head=2, pred_len=4, eval_pred_len=4,
                            medusa_num_layers=layers, **opts)
    dl = synth.make_dataload(cfg)
    torch.manual_seed(7)
    model = HSTU(cfg, dl, compute_dtype=torch.float32)
    with torch.no_grad():                    # LayerNorm / bias parameters away from their tiny init so mistakes show
        for n, p in model.named_parameters():
            if "medusa" in n or "segment_emb" in n:
                p.add_(0.3 * torch.randn_like(p))
    rows, D = 13, 32
    H = model.medusa_num_heads
    y = torch.randn(rows, D)
    d_hd = torch.randn(rows, H, D)
    with shim.installed():
        hd = model._hier_forward(y, rows)
        grads = {}
        dy = model._hier_backward(d_hd.clone(), model._hier_tape, y, grads)
    # oracle: same parameters (aliases kept), autograd
    sd = orc.state_dict_from_module(model, requires_grad=True)
    yo = y.clone().requires_grad_(True)
    ho = orc.OracleHSTU(cfg, sd, dl.category_counts, dl.category_to_int).heads(yo)        # [H, rows, D]
    assert torch.allclose(hd, ho.permute(1, 0, 2), rtol=1e-5, atol=1e-5)
    (ho.permute(1, 0, 2) * d_hd).sum().backward()
    assert torch.allclose(dy, yo.grad, rtol=1e-4, atol=1e-5 * max(1.0, float(dy.abs().max())))
    names = {p: n for n, p in model.named_parameters()}
    seen = set()
    for p, g in grads.items():
        n = names[p]
        seen.add(n)
        assert torch.allclose(g, sd[n].grad, rtol=1e-4, atol=1e-5 * max(1.0, float(g.abs().max()))), n
    expect = {n for n, _ in model.named_parameters() if n.startswith(("medusa_cat_head", "medusa_seg_head", "segment_emb"))}
    assert seen == expect


def test_hier_options_state_dict_is_the_reference_layout():
    """Same parameter names, shapes and aliasing as the reference module (fixture from the live class)."""
    fx = load_golden("hier_options")
    cfg = synth.Config(fx["cfg"])
    dl = synth.make_dataload(cfg)
    model = HSTU(cfg, dl, compute_dtype=torch.float32)
    ref_sd = fx["state_dict"]
    sd = model.state_dict()
    assert set(ref_sd) == set(sd), (sorted(set(ref_sd) - set(sd)), sorted(set(sd) - set(ref_sd)))
    for k, v in ref_sd.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    model.load_state_dict(ref_sd)                                    # strict
    assert {k for k, _ in model.named_parameters()} == set(fx["grads"])
    a = model.medusa_seg_head[0][0][0].linear.weight
    assert all(model.medusa_seg_head[c][s][0].linear.weight is a for c in range(cfg["num_prior_head"]) for s in range(2))
