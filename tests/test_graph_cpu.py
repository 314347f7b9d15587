"""CPU checks of the prior-construction oracle (oracle/graph_oracle.py): the edge builders against
hand-computed cases of the reference's loops (code/item-clustering.py:152-157, code/user-clustering.py:236-290), the
modularity against the dense-matrix definition, the local-moving + contraction on a planted partition."""
import numpy as np

from oracle import graph_oracle as go


def test_item_graph_edges_hand_case():
    # eval_pred_len 1: the last item is held out; user 3 has train_seq_len 1 -> no pair; duplicates collapse
    seqs = [[1, 2, 3, 9], [2, 3, 4, 4, 9], [5, 6], []]
    assert go.item_graph_edges(seqs, 1, 0, 200) == {(1, 2), (1, 3), (2, 3), (2, 4), (3, 4)}
    # max_user_seq_len 2: only the last two training items of every user
    assert go.item_graph_edges(seqs, 1, 0, 2) == {(2, 3)}              # user 2's window is [4, 4]: one distinct item
    # a train / test gap shortens the prefix further
    assert go.item_graph_edges(seqs, 1, 1, 200) == {(1, 2), (2, 3), (2, 4), (3, 4)}


def test_user_graph_edges_hand_case_and_slice_quirk():
    # users are 1-based.  item 7: users {1, 2, 3}; item 8: users {1, 3}
    seqs = [[7, 8, 1], [7, 2, 3], [8, 7, 4, 5]]
    assert go.user_graph_edges(seqs, 1, 0, 200) == {(1, 2), (1, 3), (2, 3)}
    # cap: only the 2 smallest user ids of item 7 form pairs; item 8 still links 1-3
    assert go.user_graph_edges(seqs, 1, 0, 200, max_users_per_item=2) == {(1, 2), (1, 3)}
    # context_len 1 < train_seq_len: list.slice(offset, train_seq_len) runs PAST the training prefix (reference quirk):
    # user 3 (len 4, train 3, offset 2) contributes items [4, 5], the held-out 5 included
    seqs2 = [[9, 9, 4, 6], [1, 2, 4, 5], [1, 2, 5, 7]]
    e = go.user_graph_edges(seqs2, 1, 0, 1)
    assert (1, 2) in e            # item 4: user 1 via [4, 6][:..], user 2 via [4, 5]
    assert (2, 3) in e            # item 5: user 2's window reaches its held-out item


def test_modularity_matches_dense_definition():
    rng = np.random.default_rng(0)
    n = 30
    A = np.triu((rng.random((n, n)) < 0.2).astype(float), 1)
    A = A + A.T
    edges = [(i, j) for i in range(n) for j in range(i + 1, n) if A[i, j]]
    memb = rng.integers(0, 4, size=n).tolist()
    k = A.sum(1)
    two_m = A.sum()
    for gamma in (1.0, 1.5):
        same = np.equal.outer(memb, memb)
        want = ((A - gamma * np.outer(k, k) / two_m) * same).sum() / two_m
        assert abs(go.modularity(n, edges, memb, gamma) - want) < 1e-12


def test_louvain_recovers_planted_partition_and_beats_trivial():
    rng = np.random.default_rng(1)
    n, blocks = 90, 3
    edges = [(i, j) for i in range(n) for j in range(i + 1, n)
             if rng.random() < (0.5 if i // 30 == j // 30 else 0.02)]
    memb = go.louvain(n, edges)
    assert len(set(memb)) == blocks
    for b in range(blocks):
        assert len(set(memb[30 * b: 30 * (b + 1)])) == 1
    assert go.modularity(n, edges, memb) > go.modularity(n, edges, [0] * n) + 0.3
    assert go.louvain(n, edges) == memb                     # deterministic
