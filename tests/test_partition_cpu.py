"""Host-side logic of the multi-GPU partition, run on the CPU with world_size-2 gloo process groups: ownership is a
partition of the blocks, every rank's sub-problem passes the product's own validation (host-only handle), and the
ownership rule + all-reduce reproduces the global sums (log-density pieces, beta/tausq statistics) without double
counting the replicated top levels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
import spamtree_b200 as sb
from spamtree_b200 import partition as part


def test_plan_is_a_partition_and_balanced():
    pb = common.make_problem(3, 20000)
    d, t = pb["d"], pb["tree"]
    nb = t["n_blocks"]
    for nr in (2, 4, 8):
        pl = part.plan(t, d["y"], nr)
        assert (pl["owner"][~pl["top"]] >= 0).all() and (pl["owner"][pl["top"]] == -1).all()
        sizes = []
        seen_blocks = np.zeros(nb, dtype=int)
        seen_rows = np.zeros(d["y"].size, dtype=int)
        for r in range(nr):
            sp = part.subproblem(d, t, pl, r, nr)
            own = sp["global_blocks"][~pl["top"][sp["global_blocks"]]]
            seen_blocks[own] += 1
            rows_own = np.concatenate([t["indexing_idx"][t["indexing_ptr"][u]:t["indexing_ptr"][u + 1]] for u in own])
            seen_rows[rows_own] += 1
            sizes.append(rows_own.size)
            # the product validates the sub-problem (chains closed under ancestry, children consistent, layout)
            m = sb.SpamTreeMV(sp["y"], sp["X"], sp["coords"], sp["mv_id"], sp["res_is_ref"], None, None, False, sp["block_names"],
                              sp["block_groups"], None, pb["beta"], pb["theta"], pb["tausq"], csr=sp["csr"], device=-1, q=3)
            m.close()
            assert sp["n_top_levels"] == pl["gc"]
            assert np.array_equal(d["y"][sp["global_rows"]], sp["y"], equal_nan=True)
        assert (seen_blocks[~pl["top"]] == 1).all() and (seen_blocks[pl["top"]] == 0).all()
        top_rows = np.concatenate([t["indexing_idx"][t["indexing_ptr"][u]:t["indexing_ptr"][u + 1]] for u in np.flatnonzero(pl["top"])])
        assert (np.delete(seen_rows, top_rows) == 1).all()
        assert max(sizes) <= 1.25 * min(sizes)
    with pytest.raises(ValueError):
        part.plan(t, d["y"], 100000)


@pytest.mark.parametrize("limited", [False, True])
def test_partitioned_layout_passes_the_products_validation(limited):
    """st_create with an st_partition on a host-only handle: the replicated frontier, its pseudo children and the layout are
    built and validated without a GPU — for full chains and for limited_tree = TRUE (direct parent only, make_edges_limited)"""
    pb = common.make_problem(3, 12000)
    d, t = pb["d"], pb["tree"]
    mp_ = None
    if limited:
        lp = sb.limited_edges_csr(t, d["y"])
        mp_ = (lp[0], lp[1])
    for nr in (2, 4):
        pl = part.plan(t, d["y"], nr)
        for r in range(nr):
            sp = part.subproblem(d, t, pl, r, nr, model_parents=mp_)
            sp["allreduce"] = lambda ptr, count: None
            m = sb.SpamTreeMV(sp["y"], sp["X"], sp["coords"], sp["mv_id"], sp["res_is_ref"], None, None, limited, sp["block_names"],
                              sp["block_groups"], None, pb["beta"], pb["theta"], pb["tausq"], csr=sp["csr"], device=-1, q=3, partition=sp)
            if limited:  # every block lists its direct parent only
                assert max(m.index("parents_indexing", u).size for u in range(m.n_blocks)) <= 60
            m.close()
    # a partition without the global row map is refused (the ranks' random streams would differ)
    sp = part.subproblem(d, t, part.plan(t, d["y"], 2), 0, 2)
    sp["allreduce"] = lambda ptr, count: None
    sp["global_rows"] = sp["global_rows"][:10]
    with pytest.raises((sb.SpamTreeError, ValueError, Exception)):
        bad = dict(sp, n_global_rows=5)
        sb.SpamTreeMV(sp["y"], sp["X"], sp["coords"], sp["mv_id"], sp["res_is_ref"], None, None, False, sp["block_names"],
                      sp["block_groups"], None, pb["beta"], pb["theta"], pb["tausq"], csr=sp["csr"], device=-1, q=3, partition=bad)


def _worker(rank, world, port, q, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pb = common.make_problem(q, n)
        d, t = pb["d"], pb["tree"]
        pl = part.plan(t, d["y"], world)
        sp = part.subproblem(d, t, pl, rank, world)
        om = common.oracle_model(pb)            # every rank holds the global oracle (CPU, small)
        w = np.random.default_rng(1).standard_normal(n)
        om.w = w
        ok, ll, ld = om.get_loglik_comps_w(0)
        comps = om.get("logdetCi_comps") + om.get("loglik_w_comps")
        own = sp["global_blocks"]
        mine = own[~pl["top"][own]]
        # rule of the library: replicated blocks are added once (not all-reduced), owned blocks are all-reduced
        local = torch.tensor([comps[mine].sum()], dtype=torch.float64)
        dist.all_reduce(local)
        total = comps[pl["top"]].sum() + local.item()
        # beta / tausq statistics: replicated rows are counted by rank 0 only
        rows = sp["global_rows"]
        top_rows = np.zeros(n, dtype=bool)
        for u in np.flatnonzero(pl["top"]):
            top_rows[t["indexing_idx"][t["indexing_ptr"][u]:t["indexing_ptr"][u + 1]]] = True
        use = rows[np.isfinite(d["y"][rows]) & ~(top_rows[rows] & (rank > 0))]
        stat = torch.tensor(d["X"][use].T @ (d["y"][use] - w[use]), dtype=torch.float64)
        cnt = torch.tensor([float(use.size)], dtype=torch.float64)
        dist.all_reduce(stat)
        dist.all_reduce(cnt)
        obs = np.isfinite(d["y"])
        want = d["X"][obs].T @ (d["y"][obs] - w[obs])
        if rank == 0:
            out.put((ok, abs(total - ll) / abs(ll), float(np.max(np.abs(stat.numpy() - want)) / np.max(np.abs(want))), cnt.item() == obs.sum()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("q,n", [(2, 3000)])
def test_ownership_rule_under_gloo_allreduce(q, n):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, n, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ok, e_ll, e_stat, cnt_ok = res
    assert ok and e_ll < 1e-12 and e_stat < 1e-12 and cnt_ok
