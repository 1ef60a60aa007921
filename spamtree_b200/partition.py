"""Subtree partition of one SpamTrees problem over the GPUs of a node (BASELINE north_star).

Blocks of a tree level are conditionally independent and every dependency runs along an ancestor chain, so the tree is
cut at a level `gc`: the few blocks above the cut ("top") are replicated on every rank and computed redundantly, each
block at the cut level owns its whole subtree, and subtrees are dealt to ranks in contiguous runs.  Per iteration the
ranks exchange only (1) three scalars of the log-density, (2) the messages of the cut-level blocks to the top blocks,
(3) the sufficient statistics of the beta / tausq steps.  The reference has no counterpart (single process, OpenMP).

This module is pure host logic (numpy); it builds, for one rank, exactly the inputs `SpamTreeMV` takes plus the
bookkeeping needed to stitch results back together.
"""
import numpy as np


def _lists(ptr, idx):
    return [idx[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]


def plan(tree, y, nranks):
    """Choose the cut and the owner of every block.  Returns dict(gc, levels, owner[n_blocks] (-1 = top), top[n_blocks])."""
    nb = tree["n_blocks"]
    lev = tree["block_groups"].astype(np.int64)
    rows = _lists(tree["indexing_ptr"], tree["indexing_idx"])
    par = _lists(tree["parents_ptr"], tree["parents_idx"])
    obs = np.array([np.isfinite(y[r]).any() for r in rows])
    levels = sorted(set(lev[obs]))
    if nranks <= 1:
        return {"gc": 0, "levels": levels, "owner": np.zeros(nb, dtype=np.int64), "top": np.zeros(nb, dtype=bool)}
    gc = None
    for i, L in enumerate(levels):
        cnt = np.count_nonzero(obs & (lev == L))
        if cnt >= nranks and (cnt % nranks == 0 or cnt >= 4 * nranks):   # enough subtrees to balance the ranks
            gc = i
            break
    if gc is None:
        raise ValueError(f"no tree level has {nranks} blocks: too many ranks for this problem")
    if tree["res_is_ref"][gc] != 1:
        raise ValueError("the cut level must be a reference level")
    cutL = levels[gc]
    top = obs & (lev < cutL)
    # subtree root (block at the cut level) of every block at or below the cut
    root = np.full(nb, -1, dtype=np.int64)
    for u in range(nb):
        if top[u]:
            continue
        if obs[u] and lev[u] == cutL:
            root[u] = u
        elif len(par[u]) > gc:
            root[u] = par[u][gc]
    # cut-level blocks in the product's slot order: by (chain of ancestors, id) == DFS order
    cut_blocks = [u for u in range(nb) if obs[u] and lev[u] == cutL]
    cut_blocks.sort(key=lambda u: (tuple(par[u]), u))
    weight = np.zeros(nb)
    for u in range(nb):
        if root[u] >= 0:
            weight[root[u]] += len(rows[u]) * (1 + len(par[u]))  # ~ BUILD cost of the block's rows
    w = np.array([weight[u] for u in cut_blocks])
    cs = np.cumsum(w)
    owner_of_cut = np.zeros(len(cut_blocks), dtype=np.int64)
    prev = 0
    for r in range(nranks):
        if r == nranks - 1:
            b = len(cut_blocks)
        else:
            target = cs[-1] * (r + 1) / nranks
            i = int(np.searchsorted(cs, target, side="left"))            # cs[i] is the first prefix sum >= target
            below = cs[i - 1] if i > 0 else 0.0
            b = i if (i > 0 and target - below <= cs[min(i, len(cs) - 1)] - target) else i + 1   # nearer boundary
            b = min(max(b, prev + 1), len(cut_blocks) - (nranks - 1 - r))  # every rank gets at least one subtree
        owner_of_cut[prev:b] = r
        prev = b
    omap = {u: owner_of_cut[i] for i, u in enumerate(cut_blocks)}
    owner = np.full(nb, -1, dtype=np.int64)
    for u in range(nb):
        if top[u]:
            continue
        owner[u] = omap[root[u]] if root[u] >= 0 else 0  # prediction blocks hanging off a top block go to rank 0
    return {"gc": gc, "levels": levels, "owner": owner, "top": top}


def subproblem(d, tree, pl, rank, nranks, model_parents=None):
    """Inputs of SpamTreeMV for one rank: the replicated top blocks plus the blocks this rank owns.
    model_parents: (ptr, idx) of the parent lists the MODEL uses when they differ from the tree's ancestor chains —
    limited_tree = TRUE (make_edges_limited: the direct parent only); the plan is always made on the full chains."""
    nb = tree["n_blocks"]
    keep = pl["top"] | (pl["owner"] == rank)
    old = np.flatnonzero(keep)
    new_id = np.full(nb, -1, dtype=np.int64)
    new_id[old] = np.arange(old.size)
    rows = _lists(tree["indexing_ptr"], tree["indexing_idx"])
    par = _lists(*model_parents) if model_parents is not None else _lists(tree["parents_ptr"], tree["parents_idx"])
    grow = np.sort(np.concatenate([rows[u] for u in old]))           # global boundary rows, boundary order kept
    lrow = np.full(d["y"].size, -1, dtype=np.int64)
    lrow[grow] = np.arange(grow.size)
    idx_l = [lrow[rows[u]] for u in old]
    par_l = [new_id[par[u]] for u in old]
    assert all((p >= 0).all() for p in par_l), "a kept block has a parent outside the partition"
    chi_l = [[] for _ in old]
    for i, p in enumerate(par_l):
        if np.isfinite(d["y"][grow[idx_l[i]]]).any():               # children lists hold observed blocks only
            for a in p:
                chi_l[a].append(i)

    def csr(ls):
        ptr = np.zeros(len(ls) + 1, dtype=np.int64)
        for i, l in enumerate(ls):
            ptr[i + 1] = ptr[i] + len(l)
        idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in ls]) if ptr[-1] else np.zeros(0, np.int64)
        return ptr, idx

    ip, ii = csr(idx_l)
    pp, pi = csr(par_l)
    cp, ci = csr(chi_l)
    top_rows = np.sort(np.concatenate([rows[u] for u in np.flatnonzero(pl["top"])])) if pl["top"].any() else np.zeros(0, np.int64)
    n_own_rows = [int(sum(len(rows[u]) for u in np.flatnonzero(pl["owner"] == r))) for r in range(nranks)]
    return {
        "y": d["y"][grow], "X": d["X"][grow], "coords": d["coords"][grow], "mv_id": d["mv_id"][grow],
        "res_is_ref": tree["res_is_ref"], "csr": (ip, ii, pp, pi, cp, ci),
        "block_names": np.arange(1, old.size + 1, dtype=np.float64), "block_groups": tree["block_groups"][old],
        "global_rows": grow, "global_blocks": old, "n_global_rows": int(d["y"].size),
        "n_top_levels": int(pl["gc"]) if nranks > 1 else 0, "n_top_rows": int(top_rows.size),
        "rng_row_offset": int(top_rows.size * 0 + sum(n_own_rows[:rank])),
        "rank": rank, "nranks": nranks,
    }


def gather_rows(parts, values, n_global):
    """stitch per-rank row vectors (local boundary order) into the global boundary order; top rows are identical"""
    out = np.full(n_global, np.nan)
    for sp, v in zip(parts, values):
        out[sp["global_rows"]] = v
    return out
