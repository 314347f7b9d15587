"""csrc/remi_core.cuh (REMI routing regulariser: one scan per (sequence, interest), closed-form gradient) compiled for the
host and checked against the reference's dense formulation (remi.py:156-196, 356-372) + autograd.  The CUDA kernel in
comirec.cu calls the same function with the same arguments; its launch is checked on the GPU (test_gpu_zz_remi.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("remi") / "libremi_host.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "remi_host.cpp"),
                           "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    lib.remi_rr_host.restype = None
    lib.remi_rr_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def reference_rr(a_list, D):
    """remi.py:156-196 on right-aligned, left-padded windows, + the masked mean over valid steps (:356-372)."""
    Lmax = max(x.shape[0] for x in a_list)
    K = a_list[0].shape[1]
    B = len(a_list)
    a = torch.zeros(B, Lmax, K, dtype=a_list[0].dtype)
    valid = torch.zeros(B, Lmax, dtype=torch.bool)
    for b, x in enumerate(a_list):                      # left padding like the reference batches
        a[b, Lmax - x.shape[0]:] = x
        valid[b, Lmax - x.shape[0]:] = True
    causal = torch.tril(torch.ones(Lmax, Lmax, dtype=torch.bool))
    keep = causal[None] & valid[:, None, :]                                   # [B, l, l']
    w = a.permute(0, 2, 1)[:, None].expand(-1, Lmax, -1, -1)                  # [B, l, K, l']
    w = torch.where(keep[:, :, None, :], w, torch.finfo(w.dtype).min)
    att = torch.nan_to_num(torch.softmax(w, dim=-1), nan=0.0)
    mk = keep[:, :, None, :].to(att.dtype).expand_as(att)
    lens = mk.sum(-1, keepdim=True).clamp(min=1.0)
    am = att * mk
    dev = (am - am.sum(-1, keepdim=True) / lens) * mk
    B_, L_, K_, _ = dev.shape
    d2 = dev.reshape(B_ * L_, K_, -1)
    cov = torch.bmm(d2, d2.transpose(1, 2)) / D
    var = torch.diagonal(cov, dim1=-2, dim2=-1)
    per_step = (torch.norm(var, dim=1) ** 2).reshape(B_, L_)
    vf = valid.to(att.dtype)
    return (per_step * vf).sum() / vf.sum().clamp(min=1.0)


@pytest.mark.parametrize("lens,K,D,scale", [([7, 1, 12, 5], 3, 32, 3.0), ([1], 4, 64, 1.0), ([50, 33, 2], 4, 32, 8.0),
                                            ([9, 0, 4], 2, 16, 0.02)])
def test_rr_scan_matches_dense_reference(host_lib, lens, K, D, scale):
    torch.manual_seed(len(lens) * 100 + K)
    a_list = [(torch.randn(n, K, dtype=torch.float64) * scale).requires_grad_(True) for n in lens]
    ref = reference_rr([x for x in a_list if x.shape[0] > 0], D)
    ref.backward()
    ref = ref.detach()
    a = torch.cat([x.detach() for x in a_list]).float().contiguous()
    T = a.shape[0]
    off = np.array([0] + list(np.cumsum(lens)), dtype=np.int32)
    # one trailing dummy sequence (static-token mode of the graphed step): must be ignored, its da stays zero
    n_dummy = 3
    a_all = torch.cat([a, torch.randn(n_dummy, K)]).contiguous()
    off_all = np.concatenate([off, [off[-1] + n_dummy]]).astype(np.int32)
    var2 = torch.zeros(T + n_dummy, K)
    da = torch.zeros(T + n_dummy, K)
    scratch = torch.empty(T + n_dummy, K, 3)
    host_lib.remi_rr_host(a_all.data_ptr(), off_all.ctypes.data, len(lens), K, D, var2.data_ptr(), scratch.data_ptr(),
                          da.data_ptr())
    rr = float(var2.sum()) / max(T, 1)
    assert abs(rr - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-12, (rr, float(ref))
    g_ref = torch.cat([x.grad if x.grad is not None else torch.zeros_like(x) for x in a_list]).float()
    assert torch.allclose(da[:T], g_ref, rtol=2e-3, atol=2e-6 * float(g_ref.abs().max()) + 1e-12), \
        float((da[:T] - g_ref).abs().max())
    assert float(da[T:].abs().max()) == 0.0 and float(var2[T:].abs().max()) == 0.0
    # forward-only call leaves da untouched
    var2b = torch.zeros_like(var2)
    host_lib.remi_rr_host(a_all.data_ptr(), off_all.ctypes.data, len(lens), K, D, var2b.data_ptr(), None, None)
    assert torch.equal(var2b, var2)


@pytest.mark.parametrize("name", ["remi_p1", "remi_p3_beta4", "remi_p2_beta0"])
def test_remi_state_dict_is_the_reference_layout(name):
    """b200rec.comirec.REMI takes the reference's config keys and loads the live class's state dict (strict)."""
    from conftest import load_golden
    from b200rec import synth
    from b200rec.comirec import REMI
    fx = load_golden(name)
    cfg = synth.Config(fx["cfg"])
    model = REMI(cfg, synth.Dataload(cfg["item_num"], {}, {}), compute_dtype=torch.float32)
    model.load_state_dict(fx["state_dict"], strict=True)
    assert {k for k, _ in model.named_parameters()} == set(fx["grads"])
    assert model.lambda_rr == float(cfg["lambda_rr"]) and model.beta_ihn == float(cfg["beta_ihn"])
    assert (model.attention_net[0].bias is None) == (cfg["attention_net_bias"] is False)
    assert model._ihn_beta == max(0.0, float(cfg["beta_ihn"]))


def test_routing_regulariser_host_glue(host_lib):
    """comirec._Readout._routing_reg (argument order of the b200rec_comi_rr call, the device-side mean, static-token
    dummy sequence) with the C-ABI call served by the host build of the kernel's source; against the oracle's dense
    restatement of remi.py:156-196 + autograd."""
    import cabi_cpu_shim as shim
    from conftest import load_golden
    from oracle.comirec_oracle import OracleREMI
    from b200rec import synth
    from b200rec.comirec import REMI
    fx = load_golden("remi_p1")
    cfg = synth.Config(fx["cfg"])
    model = REMI(cfg, synth.Dataload(cfg["item_num"], {}, {}), compute_dtype=torch.float32)
    model.load_state_dict(fx["state_dict"], strict=True)
    K, D, Lq = model.num_interest, cfg["hstu_embedding_size"], 9
    lens = [9, 4, 1, 6]
    torch.manual_seed(5)
    y = torch.randn(len(lens), Lq, D, dtype=torch.float32)
    valid = torch.zeros(len(lens), Lq, dtype=torch.bool)
    for b, n in enumerate(lens):
        valid[b, Lq - n:] = True
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    orc = OracleREMI(fx["cfg"], sd)
    a_pad = orc.attention_logits(y).detach().requires_grad_(True)                  # [B, L, K]
    orc.attention_logits = lambda _y: a_pad
    lam, rr_ref = orc.routing_loss(y, valid)
    rr_ref.backward()
    a_tok = a_pad.detach()[valid].contiguous()                                       # jagged [T, K]
    T = a_tok.shape[0]
    n_dummy = 5                                                                     # static-token mode: one dummy sequence
    a_all = torch.cat([a_tok, torch.randn(n_dummy, K)]).contiguous()
    seq_off = torch.tensor([0] + list(np.cumsum(lens)) + [T + n_dummy], dtype=torch.int32)

    def comi_rr(a, off, B_real, K_, D_, var2, scratch, da, _stream):
        host_lib.remi_rr_host(a, off, B_real, K_, D_, var2, scratch, da)

    shim.EXTRA["b200rec_comi_rr"] = comi_rr
    try:
        with shim.installed():
            rr, da = model._readout._routing_reg(model, a_all, seq_off, len(lens), T + n_dummy, True)
            rr_fwd, da_none = model._readout._routing_reg(model, a_all, seq_off, len(lens), T + n_dummy, False)
    finally:
        shim.EXTRA.pop("b200rec_comi_rr")
    assert lam == model.lambda_rr == 100.0
    assert abs(float(rr) - float(rr_ref)) <= 2e-5 * abs(float(rr_ref))
    assert da_none is None and float(rr_fwd) == float(rr)
    g_ref = a_pad.grad[valid]
    assert torch.allclose(da[:T], g_ref, rtol=2e-3, atol=2e-6 * float(g_ref.abs().max()))
    assert float(da[T:].abs().max()) == 0.0
