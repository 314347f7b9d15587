"""Fused sampled-softmax path (bf16 production mode): positives + reference exponent -> logits GEMM with the NCE_EXP
epilogue (bf16 softmax numerators + per-row partial sums, no fp32 [T, Nneg] tensor) -> combine.  Checked against
(a) a plain PyTorch fp32 restatement of hstu.py:600-629 + cross_entropy on the same bf16-rounded operands, including the
false-negative filter and several offsets sharing one query row, and (b) the unfused kernels end to end."""
import math

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from b200rec import _lib as L  # noqa: E402
from b200rec import synth  # noqa: E402
from b200rec.hstu import HSTU  # noqa: E402

DEV = torch.device("cuda:0")


def _unit(n, D, gen):
    x = torch.randn(n, D, generator=gen)
    return (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(DEV)


@pytest.mark.parametrize("T,B,LP,P,n_neg,D,scale", [(37, 4, 14, 4, 96, 64, math.log(20.0)),
                                                     (300, 8, 58, 8, 8192, 256, math.log(20.0)),
                                                     (130, 6, 30, 2, 1000, 128, math.log(100.0)),
                                                     (64, 4, 20, 1, 512, 1024, 0.3)])
def test_fused_nce_forward_and_gradient_against_torch(T, B, LP, P, n_neg, D, scale):
    gen = torch.Generator().manual_seed(T + n_neg)
    Lc = LP - P
    # tokens: T distinct (b, pos) pairs with pos < Lc
    flat = torch.randperm(B * Lc, generator=gen)[:T].sort().values
    tok_b = (flat // Lc).to(torch.int32).to(DEV)
    tok_pos = (flat % Lc).to(torch.int32).to(DEV)
    qhat = _unit(T, D, gen)
    that = _unit(B * LP, D, gen)
    nhat = _unit(n_neg, D, gen)
    # plant false negatives: some negatives are (bf16 copies of) target rows -> cos = 1 > 0.99; some twice
    plant = torch.randint(0, B * LP, (max(4, n_neg // 40),), generator=gen)
    slots = torch.randperm(n_neg, generator=gen)[:plant.numel()]
    nhat[slots.to(DEV)] = that[plant.to(DEV)]
    # make a few queries close to their p = 0 target so positives dominate somewhere (exercises the reference exponent)
    near = torch.arange(0, T, 7)
    r_near = (tok_b[near].long() * LP + tok_pos[near].long() + 1)
    qhat[near.to(DEV)] = that[r_near]
    tok_ok = (torch.rand(B * LP, 1, generator=gen) < 0.8).to(torch.uint8).to(DEV)
    p_mask = (1 << P) - 1 if P < 3 else ((1 << P) - 1) & ~2           # one offset not served by this head
    coef = (torch.rand(P, generator=gen) + 0.1).to(DEV)
    lscale = torch.tensor(scale, device=DEV)
    thres = 0.99
    n_words = (n_neg + 31) // 32
    ld_neg = n_words * 32
    # false-negative bits through the real GT_BITS GEMM
    bits = torch.empty((B * LP, n_words), dtype=torch.int32, device=DEV)
    row_any = torch.zeros(B * LP + P + 1, dtype=torch.uint8, device=DEV)
    L.gemm(that, nhat, bits, B * LP, n_neg, D, lda=D, ldb=D, ldc=n_words, epilogue=L.EPI_GT_BITS, alpha=thres, C2=row_any)
    # ---- fused path
    st = L.stream()
    pos_cos = torch.empty((T, P), device=DEV)
    mref, thr = torch.empty(T, device=DEV), torch.empty(T, device=DEV)
    L.call("b200rec_nce_pos_ref", qhat.data_ptr(), D, that.data_ptr(), D, tok_b.data_ptr(), tok_pos.data_ptr(), T, LP, P,
           p_mask, tok_ok.data_ptr(), 1, 0, lscale.data_ptr(), pos_cos.data_ptr(), mref.data_ptr(), thr.data_ptr(), st)
    n_parts = L.lib().b200rec_gemm_nce_parts(n_neg)
    E = torch.full((T, ld_neg), 7.0, dtype=torch.bfloat16, device=DEV)
    stats = torch.empty((T, n_parts, 4), device=DEV)
    # grouped launch with two identical problems exercises the per-group pointer tables; problem 0 is checked
    E2, stats2 = torch.empty_like(E), torch.empty_like(stats)
    L.gemm_grouped([(qhat, nhat, E), (qhat, nhat, E2)], T, n_neg, D, lda=D, ldb=D, ldc=ld_neg, epilogue=L.EPI_NCE_EXP,
                   nce=[(mref, thr, stats), (mref, thr, stats2)], nce_logit_scale=lscale)
    assert torch.equal(E[:, :n_neg], E2[:, :n_neg]) and torch.equal(stats, stats2)
    out = {k: torch.empty((T, P), device=DEV) for k in ("loss", "g0", "dsc")}
    rank0 = torch.empty((T, P), dtype=torch.int32, device=DEV)
    nval = torch.empty((T, P), dtype=torch.int32, device=DEV)
    rscale = torch.empty(T, device=DEV)
    qs = torch.empty((T, D), dtype=torch.bfloat16, device=DEV)
    L.call("b200rec_nce_combine", stats.data_ptr(), n_parts, E.data_ptr(), ld_neg, n_neg, bits.data_ptr(),
           row_any.data_ptr(), pos_cos.data_ptr(), mref.data_ptr(), qhat.data_ptr(), D, D, tok_b.data_ptr(),
           tok_pos.data_ptr(), T, LP, P, coef.data_ptr(), lscale.data_ptr(), out["loss"].data_ptr(), out["g0"].data_ptr(),
           out["dsc"].data_ptr(), rank0.data_ptr(), nval.data_ptr(), rscale.data_ptr(), qs.data_ptr(), D, st)
    torch.cuda.synchronize()
    # ---- torch fp32 restatement on the same bf16 operands
    tau = math.exp(min(max(scale, 0.0), math.log(100.0)))
    q32, t32, n32 = qhat.float(), that.float(), nhat.float()
    cos = q32 @ n32.t()                                                  # [T, n_neg]
    G_ref = torch.zeros_like(cos)
    r0 = tok_b.long() * LP + tok_pos.long() + 1
    same_all = (t32 @ n32.t()) > thres
    for p in range(P):
        r = r0 + p
        ok = (tok_ok[r, 0] != 0) & bool((p_mask >> p) & 1)
        zp = tau * (q32 * t32[r]).sum(1)
        z = tau * cos
        z = z.masked_fill(same_all[r], float("-inf"))
        lse = torch.logsumexp(torch.cat([zp[:, None], z], dim=1), dim=1)
        sm = torch.exp(z - lse[:, None])
        sm0 = torch.exp(zp - lse)
        c = coef[p]
        want_loss = torch.where(ok, c * (lse - zp), torch.zeros_like(lse))
        want_g0 = torch.where(ok, c * (sm0 - 1), torch.zeros_like(lse))
        zsafe = torch.where(torch.isfinite(z), z, torch.zeros_like(z))
        want_dsc = torch.where(ok, c * ((sm * zsafe).sum(1) + sm0 * zp - zp), torch.zeros_like(lse))
        assert torch.allclose(out["loss"][:, p], want_loss, rtol=2e-2, atol=2e-3 * float(c)), p
        assert torch.allclose(out["g0"][:, p], want_g0, rtol=2e-2, atol=3e-3 * float(c)), p
        assert torch.allclose(out["dsc"][:, p], want_dsc, rtol=3e-2, atol=3e-2 * float(c) * max(1.0, tau / 20)), p
        want_nv = torch.where(ok, (~same_all[r]).sum(1) + 1, torch.zeros_like(r)).to(torch.int32)
        assert torch.equal(nval[:, p], want_nv), p
        if p == 0:
            want_rank = (z > zp[:, None]).sum(1)
            got = rank0[:, 0]
            sel = ok
            # ranks count cosines above the positive's: with thousands of negatives a few sit within the fp32
            # summation-order noise (tensor-core vs torch accumulation) of the threshold and may flip
            assert (got[sel].long() - want_rank[sel]).abs().max().item() <= 3
            assert ((got[sel].long() - want_rank[sel]) != 0).float().mean().item() < 0.3
            assert (got[~sel] == -1).all()
        G_ref += torch.where(ok[:, None], tau * c * sm, torch.zeros_like(sm))
    G = rscale[:, None] * E[:, :n_neg].float()
    den = G_ref.abs().max().item()
    assert (G - G_ref).abs().max().item() <= 1.5e-2 * den, ((G - G_ref).abs().max().item(), den)
    # rows where no offset is served: nothing flows
    none = torch.isnan(pos_cos).all(1)
    assert (rscale[none] == 0).all() and torch.isfinite(E[:, :n_neg].float()).all() and torch.isfinite(rscale).all()
    assert torch.allclose(qs.float(), (rscale[:, None] * q32).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-6)
    # cosine of the whole gradient matrix
    cosine = float((G.double().flatten() @ G_ref.double().flatten()) / (G.double().norm() * G_ref.double().norm() + 1e-300))
    assert cosine > 0.9999, cosine


def _build(fx, fused):
    cfg = synth.Config(fx["cfg"])
    cfg["fused_nce"] = fused
    dl = synth.Dataload(cfg["item_num"], fx["category_counts"], fx["category_to_int"])
    m = HSTU(cfg, dl, compute_dtype=torch.bfloat16)
    m.load_state_dict(fx["state_dict"])
    return m.to(DEV).eval()


@pytest.mark.parametrize("name", ["prior_additive", "prior_mult", "nce_pred4", "prior_event_given", "nce_single"])
def test_fused_nce_model_matches_unfused_and_reference(name):
    """Whole model, bf16: fused path vs the unfused kernels (same operands: differences are the bf16 rounding of the
    stored numerators) and vs the reference fixture.  The tiny fixtures (300 items, 30 negatives) are dense in
    false-negative collisions, so the filtered-offset corrections are exercised on most rows."""
    fx = load_golden(name)
    res = {}
    for fused in (False, True):
        m = _build(fx, fused)
        out = m(tuple(t.to(DEV) for t in fx["train_batch"]))
        out["loss"].backward()
        res[fused] = (out, {k: p.grad.float().cpu() for k, p in m.named_parameters() if p.grad is not None})
    (o0, g0), (o1, g1) = res[False], res[True]
    assert abs(float(o1["loss"]) - float(o0["loss"])) <= 2e-3 * abs(float(o0["loss"]))
    assert abs(float(o1["loss"]) - fx["loss"]) <= 1e-2 * abs(fx["loss"])
    for k in o0:
        if k != "loss":
            assert abs(float(o1[k]) - float(o0[k])) <= 5e-3 * max(1.0, abs(float(o0[k]))), (k, float(o1[k]), float(o0[k]))
    for k, a in g0.items():
        b = g1[k]
        if a.numel() < 2:
            assert abs(float(a) - float(b)) <= 5e-2 * max(1e-3, abs(float(a))), k       # d logit_scale
            continue
        cos = float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm() + 1e-300))
        assert cos > 0.9995, (k, cos)
        r = fx["grads"][k]
        cr = float((b.double().flatten() @ r.double().flatten()) / (b.double().norm() * r.double().norm() + 1e-300))
        assert cr > 0.99, (k, cr)


@pytest.mark.parametrize("M,N,D", [(700, 8192, 1024), (123, 1000, 256), (58, 96, 512)])
def test_pruned_false_negative_filter_equals_full_product(M, N, D):
    """cos <= <64-dim prefix> + |tail_a| |tail_b| marks candidates, the verify kernel settles them with the full dot
    product: bit for bit the matrix (and row flags) of the full GT_BITS GEMM — on planted exact duplicates, planted
    near-duplicates on both sides of the threshold, and vectors sharing a large prefix component."""
    gen = torch.Generator().manual_seed(M + N)
    a = _unit(M, D, gen)
    b = _unit(N, D, gen)
    nd = max(3, N // 50)
    rows = torch.randint(0, M, (nd,), generator=gen).to(DEV)
    cols = torch.randperm(N, generator=gen)[:nd].to(DEV)
    b[cols] = a[rows]                                              # duplicates: cos = 1
    # near-duplicates: cos around the threshold (0.985 .. 0.995)
    k = max(3, N // 60)
    rows2 = torch.randint(0, M, (k,), generator=gen).to(DEV)
    cols2 = torch.randperm(N, generator=gen)[nd:nd + k].to(DEV)
    noise = torch.randn(k, D, generator=gen).to(DEV)
    noise = noise / noise.norm(dim=1, keepdim=True)
    eps = torch.linspace(0.10, 0.18, k, device=DEV)[:, None]
    nb = a[rows2].float() + eps * noise
    b[cols2] = (nb / nb.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    # adversarial for the bound: all the mass in the first 64 dims, parallel prefixes but different tails
    a[0, 64:] = 0
    a[0] = (a[0].float() / a[0].float().norm()).to(torch.bfloat16)
    thres = 0.99
    n_words = (N + 31) // 32
    full = torch.empty((M, n_words), dtype=torch.int32, device=DEV)
    ra_full = torch.zeros(M, dtype=torch.uint8, device=DEV)
    L.gemm(a, b, full, M, N, D, lda=D, ldb=D, ldc=n_words, epilogue=L.EPI_GT_BITS, alpha=thres, C2=ra_full)
    ta, tb = torch.empty(M, device=DEV), torch.empty(N, device=DEV)
    L.call("b200rec_tail_norm", a.data_ptr(), M, D, 64, ta.data_ptr(), L.stream())
    L.call("b200rec_tail_norm", b.data_ptr(), N, D, 64, tb.data_ptr(), L.stream())
    assert torch.allclose(ta, a[:, 64:].float().norm(dim=1), rtol=1e-5, atol=1e-6)
    pr = torch.empty((M, n_words), dtype=torch.int32, device=DEV)
    ra = torch.zeros(M, dtype=torch.uint8, device=DEV)
    L.gemm(a, b, pr, M, N, 64, lda=D, ldb=D, ldc=n_words, epilogue=L.EPI_GT_BITS, alpha=thres - 1e-5, gt=(ta, tb))
    cand = int(sum(bin(x & 0xffffffff).count("1") for x in pr.flatten().tolist())) if M * n_words < 50000 else None
    L.call("b200rec_gt_bits_verify", pr.data_ptr(), M, n_words, N, a.data_ptr(), b.data_ptr(), D, thres, ra.data_ptr(),
           L.stream())
    # production variant: the bound comes out of the tensor cores (operands augmented with the rounded-up tail norm)
    a_aug = torch.empty((M, 80), dtype=torch.bfloat16, device=DEV)
    b_aug = torch.empty((N, 80), dtype=torch.bfloat16, device=DEV)
    L.call("b200rec_prefix_aug", a.data_ptr(), M, D, 64, a_aug.data_ptr(), L.stream())
    L.call("b200rec_prefix_aug", b.data_ptr(), N, D, 64, b_aug.data_ptr(), L.stream())
    assert torch.equal(a_aug[:, :64], a[:, :64]) and (a_aug[:, 65:] == 0).all()
    assert (a_aug[:, 64].float() >= a[:, 64:].float().norm(dim=1)).all()           # rounded UP
    pr2 = torch.empty((M, n_words), dtype=torch.int32, device=DEV)
    ra2 = torch.zeros(M, dtype=torch.uint8, device=DEV)
    L.gemm(a_aug, b_aug, pr2, M, N, 80, lda=80, ldb=80, ldc=n_words, epilogue=L.EPI_GT_BITS, alpha=thres - 1e-5)
    L.call("b200rec_gt_bits_verify", pr2.data_ptr(), M, n_words, N, a.data_ptr(), b.data_ptr(), D, thres, ra2.data_ptr(),
           L.stream())
    cos = a.float() @ b.float().t()
    margin = (cos - thres).abs().min().item()
    assert margin > 2e-6, "a planted pair sits inside fp32 summation noise of the threshold: reseed"
    assert torch.equal(pr, full) and torch.equal(ra, ra_full)
    assert torch.equal(pr2, full) and torch.equal(ra2, ra_full)
    assert int(ra.sum()) >= 3
    if cand is not None:                                            # the bound prunes: few candidates beyond the true pairs
        true_pairs = int((cos > thres).sum())
        assert true_pairs <= cand <= true_pairs + max(4, N // 8)
