"""Shared helpers of the test-suite: synthetic problems at oracle-friendly sizes, product + oracle model pairs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402  (test infrastructure)
from spamtree_b200 import synth  # noqa: E402


def make_problem(q, n, seed=2021, missing=0.1, proportions=None, cell_size=25, tree_seed=0, theta=None, limited=False):
    """data + tree (product tree builder) + start values, everything the C++ boundary takes"""
    import spamtree_b200 as sb
    d = synth.make_data(q, n, proportions, missing=missing, seed=seed)
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"], cell_size=cell_size, seed=tree_seed)
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"],
           tree["children_idx"])
    if limited:  # limited_tree = TRUE: every block conditions on its direct parent only (tree_dep.cpp:133-186)
        csr = csr[:2] + sb.limited_edges_csr(tree, d["y"])
        tree = dict(tree, parents_ptr=csr[2], parents_idx=csr[3], children_ptr=csr[4], children_idx=csr[5])
    th = synth.theta_for(q) if theta is None else np.asarray(theta, dtype=np.float64)
    return {"d": d, "tree": tree, "csr": csr, "theta": th, "beta": np.zeros(3), "tausq": 0.1, "q": q, "n": n, "limited": bool(limited)}


def product_model(pb, **kw):
    import spamtree_b200 as sb
    d, t = pb["d"], pb["tree"]
    return sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], t["res_is_ref"], None, None, pb.get("limited", False), t["block_names"],
                         t["block_groups"], None, pb["beta"], pb["theta"], pb["tausq"], csr=pb["csr"], **kw)


def oracle_model(pb, flags=orc.FLAG_PROBES | orc.FLAG_CORRECT_PREDICT_CACHE):
    """the oracle as the product's checker: the product never predicts from never-computed weights (a
    reference quirk the oracle reproduces by default and tests/test_reference_driver.py pins), hence the flag"""
    d, t = pb["d"], pb["tree"]
    return orc.OracleModel(d["y"], d["X"], d["coords"], d["mv_id"], t["res_is_ref"], pb["csr"], pb.get("limited", False), t["block_names"],
                           t["block_groups"], pb["beta"], pb["theta"], pb["tausq"], flags=flags)


def relerr(a, b):
    """max-norm relative error (the contract of SURVEY §8c: per block, max-norm)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return np.inf
    den = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / den) if a.size else 0.0


def lists(csr_ptr, csr_idx):
    return [csr_idx[csr_ptr[i]:csr_ptr[i + 1]] for i in range(len(csr_ptr) - 1)]
