"""CPU oracle of the prior-construction path.  TEST INFRASTRUCTURE ONLY: nothing in the product package
imports this file.

Restates, in plain Python / numpy:
  * the co-occurrence edge sets of the reference's clustering scripts -- `item_graph_edges` follows
    code/item-clustering.py:152-162 (and the identical loop at code/user-clustering.py:218-233), `user_graph_edges`
    follows code/user-clustering.py:236-290 -- line by line (sets + itertools.combinations);
  * the deterministic synchronous modularity local moving + contraction that `b200rec.prior_graph.louvain` runs on
    the GPU (rules in csrc/graph.cu), so that memberships can be compared bit for bit;
  * modularity with a resolution parameter, from the definition.

PARITY UNPINNED for the community step: the reference calls igraph's `Graph.community_leiden`
(code/item-clustering.py:240-245).  python-igraph is a third-party dependency that is neither vendored in
/root/reference nor installed in this image, and Leiden is randomised; there is no golden membership to compare
with.  What IS pinned: the edge construction (exact, against the reference's own loop restated here and checked on
hand-computed cases) and the modularity value of any membership (against the textbook formula in
tests/test_graph_cpu.py).
"""
import itertools

import numpy as np


def item_graph_edges(user_seq, eval_pred_len, train_test_gap, max_user_seq_len):
    """code/item-clustering.py:152-157: user_seq = list of per-user item-id lists."""
    edges = set()
    for seq in user_seq:
        train_seq_len = len(seq) - eval_pred_len - train_test_gap
        if train_seq_len > 1:
            edges.update(itertools.combinations(sorted(set(seq[max(0, train_seq_len - max_user_seq_len): train_seq_len])), 2))
    return edges


def user_graph_edges(user_seq, eval_pred_len, train_test_gap, context_len, max_users_per_item=2000):
    """code/user-clustering.py:236-290.  user ids are 1-based positions in user_seq (user 0 = [PAD]).  The reference
    slices `list.slice(offset, train_seq_len)`: `train_seq_len` ITEMS starting at `offset`, i.e. the window runs past
    the training prefix when offset > 0 -- replicated.  Its `unique()` leaves the order of an item's user list
    unspecified before the MAX_USERS_PER_ITEM slice; ascending order is the convention here."""
    by_item = {}
    for uid, seq in enumerate(user_seq, start=1):
        train_seq_len = len(seq) - eval_pred_len - train_test_gap
        offset = train_seq_len - context_len if train_seq_len > context_len else 0
        items = seq[offset: offset + train_seq_len] if train_seq_len > 0 else []
        for it in items:
            by_item.setdefault(it, set()).add(uid)
    edges = set()
    for users in by_item.values():
        if len(users) > 1:
            edges.update(itertools.combinations(sorted(users)[:max_users_per_item], 2))
    return edges


def modularity(n, edges, membership, gamma=1.0):
    """Q = sum_c [ in_c / 2m - gamma (tot_c / 2m)^2 ] for an unweighted simple graph (edges: iterable of (u, v))."""
    edges = list(edges)
    two_m = 2.0 * len(edges)
    deg = np.zeros(n)
    inside = {}
    for u, v in edges:
        deg[u] += 1
        deg[v] += 1
        if membership[u] == membership[v]:
            inside[membership[u]] = inside.get(membership[u], 0) + 2
    tot = {}
    for i in range(n):
        tot[membership[i]] = tot.get(membership[i], 0.0) + deg[i]
    return sum(inside.get(c, 0) / two_m - gamma * (t / two_m) ** 2 for c, t in tot.items())


def _local_moving(n, adj, deg, two_m, gamma, max_sweeps):
    """adj[i] = {j: w} (no self loops).  Same rules as csrc/graph.cu louvain_best_move_kernel / louvain_apply_kernel."""
    comm = list(range(n))
    tot = [int(d) for d in deg]
    csize = [1] * n
    idle = 0
    for sweep in range(max_sweeps):
        parity = sweep & 1
        best = list(comm)
        for i in range(n):
            if (i & 1) != parity:
                continue
            own = comm[i]
            runs = {}
            for j, w in adj[i].items():
                runs[comm[j]] = runs.get(comm[j], 0) + w
            ki = float(deg[i])
            scale = (gamma * ki) / float(two_m)
            stay = float(runs.get(own, 0)) - scale * float(tot[own] - deg[i])
            best_s, choice = stay, own
            for c in sorted(runs):
                if c == own:
                    continue
                if csize[own] == 1 and csize[c] == 1 and c > own:
                    continue
                s = float(runs[c]) - scale * float(tot[c])
                if s > best_s:
                    best_s, choice = s, c
            best[i] = choice
        moved = 0
        for i in range(n):
            a, b = comm[i], best[i]
            if a != b:
                comm[i] = b
                tot[a] -= deg[i]
                tot[b] += deg[i]
                csize[a] -= 1
                csize[b] += 1
                moved += 1
        idle = idle + 1 if moved == 0 else 0
        if idle >= 2:
            break
    return comm


def louvain(n, edges, gamma=1.0, max_sweeps=64, max_levels=32):
    """edges: iterable of (u, v), u != v, each undirected edge once.  Returns the membership list (community ids
    0..k-1 in order of their smallest level-wise id)."""
    adj = [dict() for _ in range(n)]
    for u, v in edges:
        adj[u][v] = adj[u].get(v, 0) + 1
        adj[v][u] = adj[v].get(u, 0) + 1
    deg = [sum(a.values()) for a in adj]
    two_m = sum(deg)
    member = list(range(n))
    if two_m == 0:
        return member
    cur_n = n
    for _ in range(max_levels):
        comm = _local_moving(cur_n, adj, deg, two_m, gamma, max_sweeps)
        ids = sorted(set(comm))
        relabel = {c: k for k, c in enumerate(ids)}
        comm = [relabel[c] for c in comm]
        member = [comm[m] for m in member]
        k = len(ids)
        if k == cur_n:
            break
        new_adj = [dict() for _ in range(k)]
        new_deg = [0] * k
        for i in range(cur_n):
            new_deg[comm[i]] += deg[i]
            for j, w in adj[i].items():
                if comm[i] != comm[j]:
                    new_adj[comm[i]][comm[j]] = new_adj[comm[i]].get(comm[j], 0) + w
        adj, deg, cur_n = new_adj, new_deg, k
    return member
