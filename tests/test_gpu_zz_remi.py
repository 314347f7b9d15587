"""REMI on the HSTU body (b200rec.comirec.REMI; reference REC/model/IDNet/remi.py) against fixtures produced
by the LIVE reference class (tests/golden/make_golden_comirec.py: remi_*): loss incl. the routing regulariser and the
interest-aware hard-negative loss, logging scalars, every gradient, predict() scores; the two new kernels alone against
the host build of the same source (routing regulariser) and a float64 torch restatement (hard-negative loss)."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import load_golden
from b200rec import synth, _lib as L

pytestmark = pytest.mark.gpu
CASES = ["remi_p1", "remi_p3_beta4", "remi_p2_beta0"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev():
    return torch.device("cuda:0")


def build(fx, dtype):
    from b200rec.comirec import REMI
    cfg = synth.Config(fx["cfg"])
    model = REMI(cfg, synth.Dataload(cfg["item_num"], {}, {}), compute_dtype=dtype)
    model.load_state_dict(fx["state_dict"], strict=True)
    return cfg, model.to(dev()).eval()


@pytest.mark.parametrize("name", CASES)
def test_train_step_fp32_matches_reference(name):
    fx = load_golden(name)
    cfg, model = build(fx, torch.float32)
    out = model(tuple(t.to(dev()) for t in fx["train_batch"]))
    loss = float(out["loss"].detach())
    assert abs(float(out["rr_loss"]) - fx["logs"]["rr_loss"]) <= 1e-4 * max(1e-3, abs(fx["logs"]["rr_loss"])), \
        (float(out["rr_loss"]), fx["logs"]["rr_loss"])
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


def test_predict_scores_match_reference():
    fx = load_golden("remi_p1")
    cfg, model = build(fx, torch.float32)
    feat = model.compute_item_all()
    scores, _, _, _ = model.predict(fx["eval_batch"]["item_seq"].to(dev()), None, feat, None, None)
    assert scores.shape == fx["scores"].shape
    assert torch.allclose(scores.cpu(), fx["scores"], rtol=1e-4, atol=2e-5)


def test_train_step_bf16_within_tolerance():
    """bf16 body + the IHN row kernel on bf16 operands (un-fused fp32 logits).  Like ComiRec, the hard readout is an
    arg-max: a bf16 body can flip near-ties, which moves whole gradient rows."""
    fx = load_golden("remi_p3_beta4")
    cfg, model = build(fx, torch.bfloat16)
    out = model(tuple(t.to(dev()) for t in fx["train_batch"]))
    assert abs(float(out["loss"].detach()) - fx["loss"]) <= 2e-2 * max(1.0, abs(fx["loss"])), (float(out["loss"]), fx["loss"])
    out["loss"].backward()
    worst = {}
    for k, p in model.named_parameters():
        g_ref = fx["grads"][k]
        if g_ref is None or g_ref.numel() < 2:
            continue
        g, r = p.grad.cpu().flatten().double(), g_ref.flatten().double()
        worst[k] = float((g @ r) / (g.norm() * r.norm() + 1e-30))
    print("REMI bf16 gradient cosines (min 5):", sorted(worst.items(), key=lambda kv: kv[1])[:5])
    bad = {k: c for k, c in worst.items() if c <= 0.99}                 # measured on B200: >= 0.99995
    assert not bad, bad


def test_rr_kernel_equals_host_build_of_its_source(tmp_path):
    """b200rec_comi_rr on the GPU vs csrc/remi_core.cuh compiled for the host (which the CPU tier pins on the reference's
    dense formulation): same scans, so the results agree to rounding of expf."""
    out = tmp_path / "libremi_host.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "remi_host.cpp"),
                           "-o", str(out)])
    host = ctypes.CDLL(str(out))
    host.remi_rr_host.restype = None
    host.remi_rr_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    torch.manual_seed(1)
    lens, K, D = [7, 1, 50, 0, 12, 33], 4, 64
    n_dummy = 6
    T = sum(lens) + n_dummy
    off = torch.tensor([0] + list(np.cumsum(lens)) + [T], dtype=torch.int32)
    a = (torch.randn(T, K) * 4).contiguous()
    var2_h, da_h, sc_h = torch.zeros(T, K), torch.zeros(T, K), torch.empty(T, K, 3)
    host.remi_rr_host(a.data_ptr(), off.data_ptr(), len(lens), K, D, var2_h.data_ptr(), sc_h.data_ptr(), da_h.data_ptr())
    a_d, off_d = a.to(dev()), off.to(dev())
    var2 = torch.zeros(T, K, device=dev())
    da = torch.zeros(T, K, device=dev())
    sc = torch.empty(T, K, 3, device=dev())
    L.call("b200rec_comi_rr", a_d.data_ptr(), off_d.data_ptr(), len(lens), K, D, var2.data_ptr(), sc.data_ptr(),
           da.data_ptr(), L.stream())
    assert torch.allclose(var2.cpu(), var2_h, rtol=1e-4, atol=1e-12)
    assert torch.allclose(da.cpu(), da_h, rtol=1e-3, atol=1e-6 * float(da_h.abs().max()))
    assert float(da[T - n_dummy:].abs().max()) == 0.0
    var2b = torch.zeros(T, K, device=dev())
    L.call("b200rec_comi_rr", a_d.data_ptr(), off_d.data_ptr(), len(lens), K, D, var2b.data_ptr(), None, None, L.stream())
    assert torch.equal(var2b, var2)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("beta", [1.0, 0.3, 4.0])
def test_ihn_row_kernel_against_torch(act, beta):
    """b200rec_nce_ihn_loss_fwd: several offsets per row (p_mask 0b111), filtered negatives, a row whose negatives are ALL
    filtered (loss 0), an unused row; against remi.py:198-277 restated in float64 with autograd."""
    torch.manual_seed(int(beta * 10))
    B, Lc, P, D, n_neg = 3, 6, 3, 32, 77
    LP = Lc + P
    n_words = (n_neg + 31) // 32
    ld = n_words * 32
    tok = [(b, pos) for b in range(B) for pos in range(Lc) if not (b == 1 and pos < 2)]
    T = len(tok)
    tok_b = torch.tensor([x[0] for x in tok], dtype=torch.int32)
    tok_pos = torch.tensor([x[1] for x in tok], dtype=torch.int32)
    l2 = lambda x: x / x.norm(dim=-1, keepdim=True)
    q = l2(torch.randn(T, D)).to(act)
    th = l2(torch.randn(B * LP, D)).to(act)
    cos = (torch.rand(T, ld) * 2 - 1).float()
    tok_ok = (torch.rand(B * LP, 1) > 0.15).to(torch.uint8)
    tok_ok[2 * LP + Lc - 1 + 1: 2 * LP + Lc - 1 + 1 + P] = 0                   # token (2, Lc-1): no valid offset at all
    same = torch.rand(B * LP, n_neg) < 0.05
    same[0 * LP + 2 + 1 + 1] = True                                             # token (0, 2), offset 1: every negative filtered
    bits = torch.zeros(B * LP, n_words, dtype=torch.int64)
    for j in range(n_neg):
        bits[:, j // 32] |= same[:, j].long() << (j % 32)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    coef = torch.tensor([0.5, 0.3, 0.2]) / 7.0
    log_tau = torch.tensor(math.log(15.0))
    d = dev()
    loss = torch.empty(T, P, device=d)
    g0 = torch.empty(T, P, device=d)
    dsc = torch.empty(T, P, device=d)
    rank0 = torch.empty(T, P, dtype=torch.int32, device=d)
    nval = torch.empty(T, P, dtype=torch.int32, device=d)
    G = torch.empty(T, ld, dtype=act, device=d)
    args = [x.to(d) for x in (cos, bits, q, th, tok_b, tok_pos, tok_ok, coef, log_tau)]
    cos_d, bits_d, q_d, th_d, tb_d, tp_d, ok_d, coef_d, lt_d = args
    L.call("b200rec_nce_ihn_loss_fwd", cos_d.data_ptr(), ld, n_neg, bits_d.data_ptr(), q_d.data_ptr(), D, th_d.data_ptr(),
           L.dt(act), D, tb_d.data_ptr(), tp_d.data_ptr(), T, LP, P, 0b111, ok_d.data_ptr(), 1, 0, coef_d.data_ptr(),
           lt_d.data_ptr(), float(beta), loss.data_ptr(), g0.data_ptr(), dsc.data_ptr(), rank0.data_ptr(), nval.data_ptr(),
           G.data_ptr(), ld, L.stream())
    torch.cuda.synchronize()
    # float64 restatement on the same (rounded) operands
    qf, tf = q.double(), th.double()
    c64 = cos[:, :n_neg].double().requires_grad_(True)
    lt64 = log_tau.double().requires_grad_(True)
    tau = lt64.exp()
    total = torch.zeros((), dtype=torch.float64)
    e_loss = torch.zeros(T, P, dtype=torch.float64)
    e_g0 = torch.zeros(T, P, dtype=torch.float64)
    e_rank = torch.full((T, P), -1, dtype=torch.int64)
    e_nval = torch.zeros(T, P, dtype=torch.int64)
    zps = {}
    for t, (b, pos) in enumerate(tok):
        for p in range(P):
            r = b * LP + pos + 1 + p
            if not tok_ok[r, 0]:
                continue
            zp = (tau * (qf[t] * tf[r]).sum()).detach().requires_grad_(True)
            zps[(t, p)] = zp
            neg = (tau * c64[t]).masked_fill(same[r], float("-inf"))
            keep = ~same[r]
            e_rank[t, p] = int((neg.detach()[keep] > zp.detach()).sum())
            e_nval[t, p] = int(keep.sum()) + 1
            if keep.any():
                ln = torch.logsumexp((beta + 1) * neg, 0) - (torch.logsumexp(beta * neg, 0) - math.log(n_neg))
                lv = coef[p].double() * (torch.logaddexp(zp, ln) - zp)
                # d loss / d log tau includes the positive's share: zp depends on tau
                lv_tau = coef[p].double() * (torch.logaddexp(tau * (qf[t] * tf[r]).sum(), ln) - tau * (qf[t] * tf[r]).sum())
                e_loss[t, p] = lv.detach()
                total = total + lv_tau
                e_g0[t, p] = torch.autograd.grad(lv, zp, retain_graph=True)[0]
    total.backward()
    tol = dict(rtol=2e-4, atol=2e-6) if act == torch.float32 else dict(rtol=2e-2, atol=2e-4)
    assert torch.allclose(loss.cpu().double(), e_loss, rtol=2e-5, atol=1e-6)
    assert torch.allclose(g0.cpu().double(), e_g0, rtol=2e-4, atol=1e-7)
    assert torch.equal(rank0.cpu().long(), e_rank)
    assert torch.equal(nval.cpu().long(), e_nval)
    assert abs(float(dsc.sum()) - float(lt64.grad)) <= 2e-4 * max(1.0, abs(float(lt64.grad)))
    assert torch.allclose(G.float().cpu().double()[:, :n_neg], c64.grad, **tol)
    t_all = tok.index((0, 2))                                                    # all negatives filtered at offset 1
    if tok_ok[0 * LP + 2 + 1 + 1, 0]:
        assert float(loss[t_all, 1]) == 0.0 and float(g0[t_all, 1]) == 0.0
