"""Prior construction on the GPU (b200rec.prior_graph, csrc/graph.cu) against the CPU oracle
(oracle/graph_oracle.py): edge sets bit-exact, memberships bit-exact (same deterministic rules, integer weights, IEEE
double gains), modularity, and the produced dictionary through the reference-format adapter (priors.PriorSpec)."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu


def _interactions(n_users, n_items, lo, hi, seed, zipf=True):
    rng = np.random.default_rng(seed)
    seqs = []
    p = 1.0 / np.arange(1, n_items) ** 1.05
    p /= p.sum()
    for _ in range(n_users):
        n = int(rng.integers(lo, hi))
        seqs.append((rng.choice(np.arange(1, n_items), size=n, p=p if zipf else None)).tolist())
    seqs[3] = []                                  # an empty user and a one-item user
    seqs[5] = seqs[5][:1]
    flat = torch.tensor([i for s in seqs for i in s], dtype=torch.int64, device="cuda")
    off = torch.tensor(np.concatenate([[0], np.cumsum([len(s) for s in seqs])]), dtype=torch.int64, device="cuda")
    return seqs, flat, off


@pytest.mark.parametrize("ctx,gap,budget", [(200, 0, 1 << 26), (7, 1, 1 << 26), (200, 0, 300)], ids=["all", "window", "chunked"])
def test_item_graph_edges_exact(ctx, gap, budget):
    from b200rec import prior_graph as pg
    seqs, flat, off = _interactions(80, 300, 2, 40, seed=3)
    want = go.item_graph_edges(seqs, 2, gap, ctx)
    got = pg.item_graph_edges(flat, off, 2, gap, ctx, pair_budget=budget)
    assert got.shape[0] == len(want)
    assert set(map(tuple, got.cpu().tolist())) == want
    assert bool((got[:, 0] < got[:, 1]).all())
    key = got[:, 0] * (1 << 32) + got[:, 1]
    assert bool((key[1:] > key[:-1]).all())                       # sorted, distinct


@pytest.mark.parametrize("ctx,cap", [(200, 2000), (5, 2000), (200, 3)], ids=["all", "window-quirk", "capped"])
def test_user_graph_edges_exact(ctx, cap):
    from b200rec import prior_graph as pg
    seqs, flat, off = _interactions(70, 60, 2, 25, seed=4, zipf=False)
    want = go.user_graph_edges(seqs, 2, 0, ctx, max_users_per_item=cap)
    got = pg.user_graph_edges(flat, off, 2, 0, ctx, max_users_per_item=cap)
    assert set(map(tuple, got.cpu().tolist())) == want


def test_large_group_row_path():
    """A group with more than 64 members takes the row-wise emit path."""
    from b200rec import prior_graph as pg
    members = torch.arange(1, 201, dtype=torch.int64, device="cuda")
    got = pg.pairs_from_groups(torch.zeros_like(members), members, 1)
    assert got.shape[0] == 200 * 199 // 2
    assert set(map(tuple, got.cpu().tolist())) == {(a, b) for a in range(1, 201) for b in range(a + 1, 201)}


@pytest.mark.parametrize("gamma", [1.0, 1.6])
def test_louvain_matches_oracle_bit_for_bit(gamma):
    from b200rec import prior_graph as pg
    rng = np.random.default_rng(7)
    n = 240
    block = rng.integers(0, 6, size=n)
    edges = [(i, j) for i in range(n) for j in range(i + 1, n) if rng.random() < (0.3 if block[i] == block[j] else 0.01)]
    want = go.louvain(n, edges, gamma)
    e = torch.tensor(edges, dtype=torch.int64, device="cuda")
    got, q = pg.louvain(n, e, gamma)
    assert got.cpu().tolist() == want
    assert abs(q - go.modularity(n, edges, want, gamma)) < 1e-9
    again, _ = pg.louvain(n, e, gamma)
    assert torch.equal(got, again)
    assert q > go.modularity(n, edges, [0] * n, gamma) + 0.2


def test_end_to_end_dictionary_loads_through_the_reference_format_adapter():
    from b200rec import prior_graph as pg
    from b200rec.priors import PriorSpec
    seqs, flat, off = _interactions(300, 200, 10, 30, seed=9)
    ttg, tag, q, n_edges = pg.build_prior_from_interactions(flat, off, graph="item", num_categories=4, gamma=1.2,
                                                            eval_pred_len=2, context_len=200)
    assert n_edges == len(go.item_graph_edges(seqs, 2, 0, 200))
    spec = PriorSpec.from_source(ttg, category_by="item", tag_version="v1")
    assert spec.num_categories == len(ttg["v1"]["category_counts"]) <= 4
    assert set(ttg["v1"]["tag_to_category"][0]) == set(ttg["v1"]["category_counts"])
    sizes = torch.bincount(tag[1:])
    for i, (cat, cnt) in enumerate(ttg["v1"]["category_counts"].items()):
        assert int(sizes[i + 1]) == cnt and ttg["v1"]["tag_to_category"][i + 1] == [cat]
    assert tag.shape[0] == int(flat.max()) + 1 and int(tag[0]) == 0
