"""ComiRec-SA on the HSTU body (b200rec.comirec, csrc/comirec.cu) against fixtures produced by the LIVE
reference class (tests/golden/make_golden_comirec.py): loss, logging scalars, every gradient, predict() scores; the
readout kernels alone against plain torch."""
import pytest
import torch

from conftest import load_golden
from b200rec import synth, _lib as L

pytestmark = pytest.mark.gpu
CASES = ["comirec_p1", "comirec_p4"]


def dev():
    return torch.device("cuda:0")


def build(fx, dtype):
    from b200rec.comirec import ComiRec
    cfg = synth.Config(fx["cfg"])
    model = ComiRec(cfg, synth.Dataload(cfg["item_num"], {}, {}), compute_dtype=dtype)
    missing = model.load_state_dict(fx["state_dict"], strict=True)
    return cfg, model.to(dev()).eval()


@pytest.mark.parametrize("name", CASES)
def test_train_step_fp32_matches_reference(name):
    fx = load_golden(name)
    cfg, model = build(fx, torch.float32)
    out = model(tuple(t.to(dev()) for t in fx["train_batch"]))
    loss = float(out["loss"])
    assert abs(loss - fx["loss"]) <= 2e-5 * max(1.0, abs(fx["loss"])), (loss, fx["loss"])
    assert set(k for k in out if k != "loss") == set(k for k in fx["logs"] if k != "loss")
    out["loss"].backward()
    for k, p in model.named_parameters():
        g_ref = fx["grads"][k]
        if g_ref is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        scale = max(1e-6, g_ref.abs().max().item())
        err = (p.grad.cpu() - g_ref).abs().max().item() / scale
        assert err < 1e-3, (k, err)
    for k, v in fx["logs"].items():
        if k != "loss":
            assert abs(float(out[k]) - v) <= 1e-4 * max(1.0, abs(v)), (k, float(out[k]), v)


@pytest.mark.parametrize("name", CASES)
def test_predict_scores_match_reference(name):
    fx = load_golden(name)
    cfg, model = build(fx, torch.float32)
    feat = model.compute_item_all()
    assert torch.allclose(feat.cpu(), fx["item_feature"], rtol=1e-5, atol=1e-6)
    scores, _, _, _ = model.predict(fx["eval_batch"]["item_seq"].to(dev()), None, feat, None, None)
    assert scores.shape == fx["scores"].shape
    assert torch.allclose(scores.cpu(), fx["scores"], rtol=1e-4, atol=2e-5)


def test_train_step_bf16_within_tolerance():
    fx = load_golden("comirec_p4")
    cfg, model = build(fx, torch.bfloat16)
    out = model(tuple(t.to(dev()) for t in fx["train_batch"]))
    assert abs(float(out["loss"]) - fx["loss"]) <= 1e-2 * max(1.0, abs(fx["loss"]))
    out["loss"].backward()
    for k, p in model.named_parameters():
        g_ref = fx["grads"][k]
        if g_ref is None or g_ref.numel() < 2:
            continue
        g, r = p.grad.cpu().flatten().double(), g_ref.flatten().double()
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        # the hard readout is an arg-max: a bf16 body can flip near-ties, which moves whole gradient rows
        assert cos > 0.97, (k, cos)


def test_pooling_kernels_against_torch():
    """pool_fwd / pool_bwd on jagged sequences (one of length 1) vs the dense softmax definition + autograd."""
    torch.manual_seed(0)
    lens, K, D = [7, 1, 12, 5], 3, 40
    T, B = sum(lens), len(lens)
    off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev())
    y = torch.randn(T, D, device=dev(), requires_grad=True)
    a = (torch.randn(T, K, device=dev()) * 3).requires_grad_(True)
    u = torch.empty(T, K, D, device=dev())
    M, S = torch.empty(T, K, device=dev()), torch.empty(T, K, device=dev())
    L.call("b200rec_comi_pool_fwd", a.data_ptr(), y.data_ptr(), off.data_ptr(), B, K, D, u.data_ptr(), M.data_ptr(),
           S.data_ptr(), L.stream())
    ref = torch.zeros(T, K, D, device=dev())
    rows = []
    for b in range(B):
        s, e = int(off[b]), int(off[b + 1])
        for t in range(s, e):
            w = torch.softmax(a[s:t + 1], dim=0)                       # [t-s+1, K]
            rows.append(torch.einsum("lk,ld->kd", w, y[s:t + 1]))
    ref = torch.stack(rows)
    assert torch.allclose(u, ref.detach(), rtol=1e-5, atol=1e-6)
    du = torch.randn(T, K, D, device=dev())
    ref.backward(du)
    dy = torch.zeros(T, D, device=dev())
    da = torch.zeros(T, K, device=dev())
    L.call("b200rec_comi_pool_bwd", du.data_ptr(), u.data_ptr(), y.data_ptr(), a.data_ptr(), M.data_ptr(), S.data_ptr(),
           off.data_ptr(), B, K, D, dy.data_ptr(), da.data_ptr(), L.stream())
    assert torch.allclose(dy, y.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(da, a.grad, rtol=1e-4, atol=1e-5)
