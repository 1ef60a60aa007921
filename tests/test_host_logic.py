"""Host-side integer logic of the product (C++ behind the C ABI) against the oracle's literal restatement of
tree_dep.cpp and of the SpamTreeMV bookkeeping — bit-exact (north_star: "tree and partition indexing must match
bit-exactly").  Runs without a GPU (host-only handles, device = -1)."""
import numpy as np
import pytest

import common
from common import orc
import spamtree_b200 as sb


def test_kthresholds_bit_exact():
    rng = np.random.default_rng(0)
    for n, k in [(100, 4), (1000, 10), (997, 7), (5, 5), (12345, 40), (3, 1)]:
        x = rng.random(n)
        assert np.array_equal(sb.kthresholds(x, k), orc.kthresholds(x, k))
    x = np.round(rng.random(500), 1)  # many ties
    assert np.array_equal(sb.kthresholds(x, 8), orc.kthresholds(x, 8))
    assert list(sb.kthresholds(np.arange(1, 101.0), 4)) == [26, 51, 76]


def test_part_axis_parallel_lmt_bit_exact():
    rng = np.random.default_rng(1)
    c = rng.random((400, 2))
    thr = [np.sort(rng.random(9)), np.sort(rng.random(4))]
    thr[0][3] = c[7, 0]  # a point exactly on a threshold counts as above it (>=)
    a, b = sb.part_axis_parallel_lmt(c, thr), orc.part_axis_parallel_lmt(c, thr)
    assert np.array_equal(a, b) and a[7, 0] == 1 + (thr[0] <= c[7, 0]).sum()
    thr_u = [rng.random(5), np.zeros(0)]  # unsorted thresholds and an axis without thresholds
    assert np.array_equal(sb.part_axis_parallel_lmt(c, thr_u), orc.part_axis_parallel_lmt(c, thr_u))


def test_number_revalue_bit_exact():
    rng = np.random.default_rng(2)
    om = rng.integers(0, 50, size=(60, 5))
    fv = np.arange(1, 41)
    tv = rng.integers(0, 30, size=40)
    fv[5] = fv[4]  # duplicate key: first match wins
    assert np.array_equal(sb.number_revalue(om, fv, tv), orc.number_revalue(om, fv, tv))


def _same_edges(a, b):
    assert len(a["parents"]) == len(b["parents"])
    for x, y in zip(a["parents"], b["parents"]):
        assert np.array_equal(x, y)
    for x, y in zip(a["children"], b["children"]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("q,n,missing", [(1, 625, .1), (3, 3000, .1), (2, 1500, 0.0), (1, 40, .2)])
def test_make_edges_bit_exact_on_generated_trees(q, n, missing):
    pb = common.make_problem(q, n, missing=missing)
    t = pb["tree"]
    y = pb["d"]["y"]
    ne = np.unique(t["blocking"][np.isfinite(y)])
    for limited in (False, True):
        a = sb.make_edges_limited(t["parchi_map"], ne, t["res_is_ref"]) if limited else sb.make_edges(t["parchi_map"], ne, t["res_is_ref"])
        b = orc.make_edges(t["parchi_map"], ne, t["res_is_ref"], limited)
        _same_edges(a, b)
    # and the lists the tree builder itself returned are those of make_edges
    par = common.lists(t["parents_ptr"], t["parents_idx"])
    b = orc.make_edges(t["parchi_map"], ne, t["res_is_ref"], False)
    for x, yv in zip(par, b["parents"]):
        assert np.array_equal(x, yv)


def test_make_edges_random_parchimat_with_na():
    rng = np.random.default_rng(3)
    L, nr = 4, 40
    pm = np.full((nr, L), np.nan)
    nxt = 1
    ids = {}
    for i in range(nr):
        path = tuple(rng.integers(0, 2, size=L))
        for l in range(L):
            key = (l, path[:l + 1])
            if key not in ids:
                ids[key] = nxt
                nxt += 1
    # ids must grow with the level like make_tree's: renumber level by level
    order = sorted(ids, key=lambda k: (k[0], k[1]))
    ren = {k: i + 1 for i, k in enumerate(order)}
    rows = sorted({k[1] for k in ids if k[0] == L - 1})
    pm = np.array([[ren[(l, r[:l + 1])] for l in range(L)] for r in rows], dtype=float)
    pm[rng.random(pm.shape) < 0.1] = np.nan
    pm[:, L - 1] = np.where(np.isnan(pm[:, L - 1]), ren[(L - 1, rows[0])], pm[:, L - 1])
    pm[0, L - 1] = max(ren.values())
    ne = np.array(sorted(set(ren.values()) - {3, 5}))
    for rr in ([1, 1, 1, 0], [1, 0, 1, 0], [0, 0, 0, 0]):
        _same_edges(sb.make_edges(pm, ne, rr), orc.make_edges(pm, ne, rr, False))
        _same_edges(sb.make_edges_limited(pm, ne, rr), orc.make_edges(pm, ne, rr, True))


@pytest.mark.parametrize("q,n,missing", [(1, 625, .1), (3, 3000, .1), (2, 1500, 0.0), (5, 4000, .15), (1, 30, .1)])
def test_model_bookkeeping_bit_exact(q, n, missing):
    """init_indexing / na_study / make_gibbs_groups / init_finalize (spamtree_model.cpp:194-420)"""
    pb = common.make_problem(q, n, missing=missing)
    gm, om = common.product_model(pb, device=-1), common.oracle_model(pb)
    nb = pb["tree"]["n_blocks"]
    for name in ["blocks_not_empty", "blocks_predicting", "block_is_reference", "block_ct_obs", "n_actual_groups"]:
        assert np.array_equal(gm.index(name), om.geti(name)), name
    for g in range(int(om.geti("n_actual_groups")[0])):
        assert np.array_equal(gm.index("u_by_block_groups", g), om.geti("u_by_block_groups", g))
    chi = common.lists(pb["tree"]["children_ptr"], pb["tree"]["children_idx"])
    for u in range(nb):
        for name in ["parents_indexing", "children_indexing", "dim_by_parent", "this_is_jth_child"]:
            assert np.array_equal(gm.index(name, u), om.geti(name, u)), (name, u)
        for c in range(0, len(chi[u]), max(1, len(chi[u]) // 5)):
            first, last = gm.index("u_is_which_col", u, c)
            local, other = om.geti("u_is_which_col", u, 0, c), om.geti("u_is_which_col", u, 1, c)
            assert np.array_equal(local, np.arange(first, last))
            dimen = om.geti("parents_indexing", chi[u][c]).size
            assert np.array_equal(other, np.setdiff1d(np.arange(dimen), local))
    gm.close()
    om.close()


@pytest.mark.parametrize("q,n,missing", [(1, 625, .1), (3, 5000, .1), (2, 1500, 0.0), (1, 12, 0.0)])
def test_make_tree_structure(q, n, missing):
    """properties of R/make_tree.R that the deterministic stand-in must keep"""
    pb = common.make_problem(q, n, missing=missing)
    t, y = pb["tree"], pb["d"]["y"]
    nb = t["n_blocks"]
    rows = common.lists(t["indexing_ptr"], t["indexing_idx"])
    assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(n))          # a partition of the rows
    assert all(np.all(np.diff(r) > 0) for r in rows)                            # ascending (split(0:(n-1), blocking))
    assert np.array_equal(np.sort(t["block_names"]), np.arange(1, nb + 1))      # contiguous names
    lev = t["block_groups"].astype(int)
    par = common.lists(t["parents_ptr"], t["parents_idx"])
    obs = np.array([np.isfinite(y[r]).sum() for r in rows])
    assert all(np.all(np.isfinite(y[r])) or np.all(np.isnan(y[r])) for r in rows)  # blocks are all-observed or all-missing
    nlev_obs = len(set(lev[obs > 0]))
    rr = t["res_is_ref"]
    assert len(rr) == len(set(lev))
    if n > 100:
        assert rr[nlev_obs - 1] == 0 and all(rr[:nlev_obs - 1] == 1)            # last_not_reference (R/make_tree.R:162-165)
    for u in range(nb):
        # block ids grow with the level; the parent set is the chain of reference ancestors, root first
        assert all(lev[p] < lev[u] for p in par[u])
        assert all(np.diff(par[u]) > 0)
        if len(par[u]):
            lp = par[u][-1]
            assert np.array_equal(par[lp], par[u][:-1])
        if obs[u] > 0 and lev[u] <= nlev_obs - 1:
            assert rows[u].size <= 25                                             # one knot per cell of a 5 x 5 grid
    # same seed -> same tree; another seed -> same structure, other knots
    t2 = sb.make_tree(pb["d"]["coords"], y, pb["d"]["mv_id"], seed=0)
    assert np.array_equal(t2["blocking"], t["blocking"])


def test_colocated_outcomes_are_kept_together():
    """cherrypick_group_locations (R/make_tree.R:94-99): a picked knot drags every outcome at that location along"""
    rng = np.random.default_rng(4)
    n0 = 400
    c = rng.random((n0, 2))
    coords = np.vstack([c, c])
    mv = np.r_[np.ones(n0, int), 2 * np.ones(n0, int)]
    order = np.lexsort((np.arange(2 * n0), coords[:, 1], coords[:, 0]))
    coords, mv = coords[order], mv[order]
    y = rng.standard_normal(2 * n0)
    t = sb.make_tree(coords, y, mv, cherrypick_group_locations=True)
    b = t["blocking"]
    assert np.array_equal(b[0::2], b[1::2])  # sorted rows come in co-located pairs
    t2 = sb.make_tree(coords, y, mv, cherrypick_group_locations=False)
    assert not np.array_equal(t2["blocking"][0::2], t2["blocking"][1::2])


def test_malformed_edge_lists_are_rejected_with_a_status_code():
    """the ABI promises status codes: a parents / children list that names a block outside 0..n_blocks-1 must not be indexed"""
    import spamtree_b200 as sb
    pb = common.make_problem(2, 600)
    d, t = pb["d"], pb["tree"]
    for which in (3, 5):  # parents_idx, children_idx
        csr = [a.copy() for a in pb["csr"]]
        csr[which][0] = t["n_blocks"] + 7
        with pytest.raises(sb.SpamTreeError) as e:
            sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], t["res_is_ref"], None, None, False, t["block_names"], t["block_groups"],
                          None, pb["beta"], pb["theta"], pb["tausq"], csr=tuple(csr), device=-1)
        assert e.value.code == 1 and "block id" in str(e.value)
