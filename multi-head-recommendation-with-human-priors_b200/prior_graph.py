"""Prior construction on the GPU: co-occurrence graph + community detection -> a prior dictionary in the
reference's `tag_to_general` format (consumed by `priors.PriorSpec.from_source`, i.e. by dataload.py:347-371).

    reference                                         here
    code/item-clustering.py:152-162                   item_graph_edges   (users' training windows -> item pairs)
    code/user-clustering.py:236-290                   user_graph_edges   (items' user lists, capped -> user pairs)
    code/item-clustering.py:227-250 (igraph Leiden)   louvain            (deterministic modularity local moving +
                                                                          contraction; igraph is not part of the
                                                                          reference tree -- see oracle/graph_oracle.py)
    REC/data/*_cluster_dict.py (hand-assembled)       cluster_prior_dict

torch is used for device memory and for the sorts / scans between our kernels (csrc/graph.cu: pair emission, move
evaluation, move application).  Every weight is an integer, so results are bit-reproducible.
"""
import torch

from . import _lib as L


def _group_offsets(group_sorted, n_groups):
    cnt = torch.bincount(group_sorted, minlength=n_groups)
    off = torch.zeros(n_groups + 1, dtype=torch.int64, device=group_sorted.device)
    off[1:] = torch.cumsum(cnt, 0)
    return off


def pairs_from_groups(group, member, n_groups, cap=0, pair_budget=1 << 26):
    """All unordered pairs of the DISTINCT members of every group, de-duplicated over groups.
    group, member: int64 [nnz] on the GPU (any order, duplicates allowed; member < 2^31).  cap > 0 keeps only the `cap`
    smallest distinct members of a group.  Returns int64 [E, 2] (a < b), sorted by (a, b).  Groups are processed in
    chunks of at most `pair_budget` pairs (8 B each) and merged into the running edge set."""
    dev = group.device
    if group.numel() == 0:
        return torch.zeros((0, 2), dtype=torch.int64, device=dev)
    M = int(member.max().item()) + 1
    key = torch.unique(group * M + member)                  # sorted (group, member), distinct
    g_sorted = torch.div(key, M, rounding_mode="floor")
    members = (key - g_sorted * M).to(torch.int32).contiguous()
    off = _group_offsets(g_sorted, n_groups)
    counts = torch.empty(n_groups, dtype=torch.int64, device=dev)
    L.call("b200rec_group_pairs_count", off.data_ptr(), n_groups, cap, counts.data_ptr(), L.stream())
    csum = torch.cumsum(counts, 0)
    total = int(csum[-1].item())
    if total == 0:
        return torch.zeros((0, 2), dtype=torch.int64, device=dev)
    csum_h = csum.cpu()
    edges = None
    g0 = 0
    while g0 < n_groups:
        base = int(csum_h[g0 - 1]) if g0 > 0 else 0
        g1 = int(torch.searchsorted(csum_h, torch.tensor(base + pair_budget), right=True))
        g1 = max(g1, g0 + 1)                                # a single group larger than the budget still goes through
        g1 = min(g1, n_groups)
        n_pairs = int(csum_h[g1 - 1]) - base
        if n_pairs > 0:
            pair_off = (csum[g0:g1] - counts[g0:g1] - base).contiguous()
            keys = torch.empty(n_pairs, dtype=torch.int64, device=dev)
            # group offsets stay absolute: members is indexed with them
            L.call("b200rec_group_pairs_emit", members.data_ptr(), off[g0:g1 + 1].contiguous().data_ptr(),
                   pair_off.data_ptr(), g1 - g0, cap, keys.data_ptr(), L.stream())
            keys = torch.unique(keys)
            edges = keys if edges is None else torch.unique(torch.cat([edges, keys]))
        g0 = g1
    return torch.stack([edges >> 32, edges & 0xFFFFFFFF], dim=1)


def _windows(seq_off, start, stop):
    """Flat positions of [start_u, stop_u) inside every user's segment -> (user index, flat position)."""
    n = (stop - start).clamp_min(0)
    u = torch.repeat_interleave(torch.arange(n.numel(), device=n.device), n)
    first = torch.zeros(n.numel() + 1, dtype=torch.int64, device=n.device)
    first[1:] = torch.cumsum(n, 0)
    pos = torch.arange(int(first[-1].item()), device=n.device) - first[u] + start[u] + seq_off[u]
    return u, pos


def item_graph_edges(items, seq_off, eval_pred_len, train_test_gap, max_user_seq_len, pair_budget=1 << 26):
    """code/item-clustering.py:152-157.  items: int64 [nnz] item ids, user u = items[seq_off[u] : seq_off[u+1]]."""
    length = seq_off[1:] - seq_off[:-1]
    train_len = length - eval_pred_len - train_test_gap
    start = (train_len - max_user_seq_len).clamp_min(0)
    stop = torch.where(train_len > 1, train_len, start)                    # users with train_len <= 1 contribute nothing
    u, pos = _windows(seq_off, start, stop)
    return pairs_from_groups(u, items[pos], length.numel(), pair_budget=pair_budget)


def user_graph_edges(items, seq_off, eval_pred_len, train_test_gap, context_len, max_users_per_item=2000,
                     pair_budget=1 << 26):
    """code/user-clustering.py:236-290 (user ids are 1-based, 0 = [PAD]; the window is `train_seq_len` items from
    `offset`, as the reference's `list.slice(offset, train_seq_len)` takes them)."""
    length = seq_off[1:] - seq_off[:-1]
    train_len = length - eval_pred_len - train_test_gap
    start = torch.where(train_len > context_len, train_len - context_len, torch.zeros_like(train_len))
    stop = torch.where(train_len > 0, torch.minimum(start + train_len, length), start)
    u, pos = _windows(seq_off, start, stop)
    n_items = int(items.max().item()) + 1 if items.numel() else 0
    return pairs_from_groups(items[pos], u + 1, n_items, cap=max_users_per_item, pair_budget=pair_budget)


def _local_moving(src, dst, w, deg, two_m, gamma, max_sweeps):
    """One level: directed entries (src, dst, w) both ways, no self loops; returns comm int32 [n]."""
    dev = deg.device
    n = deg.numel()
    comm = torch.arange(n, dtype=torch.int32, device=dev)
    tot = deg.clone()
    csize = torch.ones(n, dtype=torch.int32, device=dev)
    best = torch.empty(n, dtype=torch.int32, device=dev)
    moved = torch.zeros(1, dtype=torch.int64, device=dev)
    idle = 0
    for sweep in range(max_sweeps):
        # runs (node, neighbouring community) -> summed weight, ascending in both
        key = src * n + comm[dst].to(torch.int64)
        ukey, inv = torch.unique(key, return_inverse=True)
        run_w = torch.zeros(ukey.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, w)   # integer: order-free
        run_node = torch.div(ukey, n, rounding_mode="floor")
        run_comm = (ukey - run_node * n).to(torch.int32).contiguous()
        node_off = _group_offsets(run_node, n)
        L.call("b200rec_louvain_best_move", node_off.data_ptr(), run_comm.data_ptr(), run_w.data_ptr(), comm.data_ptr(),
               deg.data_ptr(), tot.data_ptr(), csize.data_ptr(), n, two_m, float(gamma), sweep & 1, best.data_ptr(),
               L.stream())
        moved.zero_()
        L.call("b200rec_louvain_apply", best.data_ptr(), comm.data_ptr(), deg.data_ptr(), tot.data_ptr(), csize.data_ptr(),
               n, moved.data_ptr(), L.stream())
        idle = idle + 1 if int(moved.item()) == 0 else 0
        if idle >= 2:
            break
    return comm


def louvain(n, edges, gamma=1.0, max_sweeps=64, max_levels=32):
    """edges: int64 [E, 2] on the GPU, each undirected edge once, no self loops.  Returns (membership int64 [n] with
    community ids 0..k-1, modularity at resolution gamma)."""
    dev = edges.device
    member = torch.arange(n, dtype=torch.int64, device=dev)
    if edges.numel() == 0:
        return member, 0.0
    src = torch.cat([edges[:, 0], edges[:, 1]])
    dst = torch.cat([edges[:, 1], edges[:, 0]])
    w = torch.ones(src.numel(), dtype=torch.int64, device=dev)
    deg = torch.zeros(n, dtype=torch.int64, device=dev).index_add_(0, src, w)
    two_m = int(deg.sum().item())
    src0, dst0, deg0 = src, dst, deg
    cur_n = n
    for _ in range(max_levels):
        comm = _local_moving(src, dst, w, deg, two_m, gamma, max_sweeps).to(torch.int64)
        ids, comm = torch.unique(comm, return_inverse=True)               # relabel 0..k-1 in id order
        member = comm[member]
        k = ids.numel()
        if k == cur_n:
            break
        new_deg = torch.zeros(k, dtype=torch.int64, device=dev).index_add_(0, comm, deg)
        cs, cd = comm[src], comm[dst]
        keep = cs != cd
        key, inv = torch.unique(cs[keep] * k + cd[keep], return_inverse=True)
        w = torch.zeros(key.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, w[keep])
        src = torch.div(key, k, rounding_mode="floor")
        dst = key - src * k
        deg, cur_n = new_deg, k
    inside = (member[src0] == member[dst0]).to(torch.float64)
    in_c = torch.zeros(int(member.max().item()) + 1, dtype=torch.float64, device=dev).index_add_(0, member[src0], inside)
    tot_c = torch.zeros_like(in_c).index_add_(0, member, deg0.to(torch.float64))
    q = float((in_c / two_m - gamma * (tot_c / two_m) ** 2).sum().item())
    return member, q


def cluster_prior_dict(membership, num_categories, version="v1", name="cluster"):
    """membership int64 [n] (node 0 = [PAD] is ignored) -> (tag_to_general dict in the format of
    REC/data/eb_nerd_512_cluster_dict.py, tag int64 [n]).  The `num_categories` largest communities become the
    categories `cluster_1 .. cluster_C` (by size, ties by community id); every other node gets tag 0, which the
    shipped dictionaries map to ALL categories."""
    m = membership[1:]
    sizes = torch.bincount(m)
    order = torch.argsort(sizes * (sizes.numel() + 1) + (sizes.numel() - torch.arange(sizes.numel(), device=m.device)),
                          descending=True)[:num_categories]
    new_id = torch.zeros(sizes.numel(), dtype=torch.int64, device=m.device)
    new_id[order] = torch.arange(1, order.numel() + 1, device=m.device)
    tag = torch.zeros_like(membership)
    tag[1:] = new_id[m]
    counts = {f"{name}_{i + 1}": int(sizes[order[i]].item()) for i in range(order.numel())}
    total = max(1, sum(counts.values()))
    cats = list(counts)
    spec = {"tag_to_category": {0: list(cats), **{i + 1: [cats[i]] for i in range(len(cats))}},
            "category_counts": counts,
            "category_percent": {c: counts[c] / total for c in cats}}
    return {version: spec}, tag


def build_prior_from_interactions(items, seq_off, graph="item", num_categories=8, gamma=1.0, eval_pred_len=1,
                                  train_test_gap=0, context_len=200, max_users_per_item=2000):
    """End to end: interactions -> co-occurrence graph -> communities -> prior dictionary.  Returns
    (tag_to_general, tag per node, modularity, number of edges)."""
    if graph == "item":
        edges = item_graph_edges(items, seq_off, eval_pred_len, train_test_gap, context_len)
        n = int(items.max().item()) + 1
    elif graph == "user":
        edges = user_graph_edges(items, seq_off, eval_pred_len, train_test_gap, context_len, max_users_per_item)
        n = seq_off.numel()                                  # users are 1-based
    else:
        raise ValueError(f"graph = {graph!r}: 'item' or 'user'")
    member, q = louvain(n, edges, gamma)
    ttg, tag = cluster_prior_dict(member, num_categories)
    return ttg, tag, q, edges.shape[0]
